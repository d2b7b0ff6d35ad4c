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
    assert len(lines) == 2 and all("crd True vel True" in ln for ln in lines), r.stdout[-2000:]
