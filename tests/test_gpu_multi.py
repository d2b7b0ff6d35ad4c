"""Multi-GPU GPU tests (need >= 2 CUDA devices; skipped on a single-GPU box): the sharded engine over
NCCL — replicated, halo and all-gather modes, fused scan + hit broadcast over peer memory — must be
bit-identical to the single-GPU engine (tools/dist_check.py under torchrun)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_engine_bit_identical_to_single_gpu(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29700 + world), os.path.join(ROOT, "tools", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if "bitwise" in ln]
    assert len(lines) == 4 and all("crd True vel True energies True" in ln for ln in lines), r.stdout[-2000:]


@pytest.mark.parametrize("world", [2, 8])
def test_single_process_group_bit_identical_to_single_gpu(world):
    """edmd_group_*: one process driving several GPUs (what `edmddna --gpus N` uses) == one GPU, bit for bit."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    import importlib
    import numpy as np
    sys.path.insert(0, ROOT)
    pkg = importlib.import_module("msc-project_b200")
    s = pkg.make_ring(200, kind="coil", laps=4, lap_gap=9.0)
    grp = pkg.EngineGroup(s, world)
    one = pkg.Engine(s, device=0)
    for eng in (grp, one):
        eng.set_rng(1000000000, 1, 0)
    assert grp.build_pairlist() == one.build_pairlist()
    for eng in (grp, one):
        eng.step(5, 1); eng.merge()
    assert grp.build_pairlist() == one.build_pairlist()
    for eng in (grp, one):
        eng.step(4, 1)
    (cg, vg), (c1, v1) = grp.get_state(), one.get_state()
    assert np.array_equal(cg, c1) and np.array_equal(vg, v1)
    assert np.array_equal(grp.energies(), one.energies())
    grp.close(); one.close()


def test_edmddna_driver_on_two_gpus_writes_the_same_files(tmp_path):
    """`edmddna --gpus 2` (single process, edmd_group_*, side-stream output) must write byte-identical energies,
    trajectory and restart files to the one-GPU run of the same deck."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import importlib
    sys.path.insert(0, ROOT)
    pkg = importlib.import_module("msc-project_b200")
    fm = importlib.import_module("msc-project_b200.formats")
    s = pkg.make_ring(60, kind="coil", laps=2, lap_gap=9.0, nev=6)
    exe = os.path.join(ROOT, "msc-project_b200", "edmddna")
    outs = []
    for gpus in (1, 2):
        d = str(tmp_path / f"g{gpus}")
        os.makedirs(d)
        fm.write_prm(os.path.join(d, "t.prm"), s); fm.write_crd(os.path.join(d, "t.crd"), s)
        fm.write_config(os.path.join(d, "config.txt"), s, "./t.prm", "./t.crd", "./e.eng", "./t.trj", "./new.crd",
                        nsteps=200, ntsync=50, ntwt=50, ntpr=100)
        r = subprocess.run([exe, "--time", "1000000000", "--gpus", str(gpus)], cwd=d, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append([open(os.path.join(d, f)).read() for f in ("e.eng", "t.trj", "new.crd")])
    assert outs[0] == outs[1]
    assert len([ln for ln in outs[0][0].splitlines() if ln]) == 4
