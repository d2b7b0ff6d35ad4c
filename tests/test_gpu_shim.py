"""GPU tests of the reference-facing C++ surface: class EDMD method by method (batch-of-one calls,
msc-project_b200/host/edmd.cpp) and the config-driven `edmddna` driver, against the CPU oracle."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np
import pytest

from test_host_io import hostlib, ROOT, PD  # noqa: F401  (fixture re-use)

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _p(a):
    return np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(PD)


def _static(ps):
    return (ps.num_atoms, ps.num_evecs, _p(ps.avg), _p(ps.masses3), _p(ps.abq), _p(ps.evals), _p(ps.evecs))


@pytest.fixture(scope="module")
def coil(edmd):
    return edmd.make_ring(40, kind="coil", laps=2, lap_gap=9.0)


def test_class_edmd_ed_forces(edmd, po, hostlib, coil):
    o = po.Oracle(coil, "port")
    off = coil.off3
    for t in (0, 7, 23):
        ps = coil.psets[int(coil.param_id[t])]
        crd = coil.crd[off[t]:off[t + 1]].copy()
        ed = np.zeros(crd.size + 1)
        hostlib.shim_ed(_p(coil.params), *_static(ps), _p(crd), _p(ed))
        o.ed_forces(t)
        f = o.get_forces(); c, _ = o.get_state()
        assert rel(ed[:-1], f["ed"][off[t]:off[t + 1]]) < TOL
        assert abs(ed[-1] - f["ed_energy"][t]) <= TOL * abs(f["ed_energy"][t])
        assert rel(crd, c[off[t]:off[t + 1]]) < TOL


def test_class_edmd_nb_forces_unclipped(edmd, po, hostlib, coil):
    o = po.Oracle(coil, "port"); o.build_pairs()
    off = coil.off3
    ps = coil.psets[0]
    hostlib.shim_nb.restype = None
    done = 0
    for t1, t2 in o.get_pairs():
        if int(coil.param_id[t1]) != 0 or int(coil.param_id[t2]) != 0:
            continue
        o.nb_forces_pair(int(t1), int(t2))
        f = o.get_forces()
        w1 = np.concatenate([f["nb"][off[t1]:off[t1 + 1]], f["nb_energy"][t1]])
        w2 = np.concatenate([f["nb"][off[t2]:off[t2 + 1]], f["nb_energy"][t2]])
        n1 = np.zeros(w1.size); n2 = np.zeros(w2.size)
        hostlib.shim_nb(_p(coil.params), *_static(ps), _p(coil.crd[off[t1]:off[t1 + 1]]), _p(coil.crd[off[t2]:off[t2 + 1]]), _p(n1), _p(n2))
        assert np.array_equal(n1[:-2], w1[:-2]) and np.array_equal(n2[:-2], w2[:-2])   # forces bit-identical, not clipped
        assert np.allclose(n1[-2:], w1[-2:], rtol=1e-12) and np.array_equal(n1[-2:], n2[-2:])
        done += 1
    assert done >= 3


def test_class_edmd_random_terms(edmd, po, hostlib, coil):
    hostlib.shim_random_calls.restype = C.c_long
    o = po.Oracle(coil, "port"); o.set_time(1234567)
    ps = coil.psets[0]
    for rank in (1, 3):
        o.set_rng_calls(hostlib.shim_random_calls())
        out = np.zeros(3 * ps.num_atoms)
        hostlib.shim_random(_p(coil.params), *_static(ps), rank, C.c_long(1234567), _p(out))
        o.random_terms(0, rank=rank)
        assert np.array_equal(out, o.get_forces()["rnd"][:out.size])


def test_class_edmd_update_velocities_coordinates(edmd, po, hostlib, coil):
    rng = np.random.default_rng(8)
    ps = coil.psets[0]
    n3 = 3 * ps.num_atoms
    crd = coil.crd[:n3].copy(); vel = rng.normal(size=n3) * 0.05
    ed, rnd, nb = rng.normal(size=n3), rng.normal(size=n3), rng.uniform(-1, 1, n3)
    s1 = coil.subset([0])
    o = po.Oracle(s1, "port"); o.set_state(crd, vel); o.set_forces(ed, rnd, nb)
    o.update_velocities(); o.update_coordinates()
    temp = C.c_double()
    hostlib.shim_integrate(_p(coil.params), *_static(ps), _p(crd), _p(vel), _p(ed), _p(rnd), _p(nb), C.byref(temp))
    co, vo = o.get_state()
    assert rel(crd, co) < TOL and rel(vel, vo) < TOL
    assert abs(temp.value - o.get_forces()["temperature"][0]) < TOL * temp.value


def test_edmddna_driver_end_to_end(edmd, po, hostlib, tmp_path):
    """The reference's input deck in, the reference's output files out: 150 steps, ntsync 50, random on."""
    fm = importlib.import_module("msc-project_b200.formats")
    s = edmd.make_ring(30, kind="coil", laps=2, lap_gap=9.0, nev=6)
    d = str(tmp_path)
    fm.write_prm(os.path.join(d, "t.prm"), s); fm.write_crd(os.path.join(d, "t.crd"), s)
    fm.write_config(os.path.join(d, "config.txt"), s, "./t.prm", "./t.crd", "./e.eng", "./t.trj", "./new.crd",
                    nsteps=150, ntsync=50, ntwt=50, ntpr=100)
    exe = os.path.join(ROOT, "msc-project_b200", "edmddna")
    r = subprocess.run([exe, "--time", "1000000000"], cwd=d, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    # oracle run of the same protocol (simulation.cpp:34-55)
    s2, cfg = fm.read_deck(os.path.join(d, "config.txt"))
    o = po.Oracle(s2, "port"); o.set_time(1000000000); o.set_rng_calls(0)
    lines = []
    for istep in range(0, 150, 50):
        o.build_pairs(); o.step(50, random_on=True); o.merge()
        lines.append((istep + 50, o.energies()))
    got = [ln for ln in open(os.path.join(d, "e.eng")).read().splitlines() if ln]
    assert len(got) == 3
    for ln, (it, e) in zip(got, lines):
        assert ln.startswith("Iteration, ED Energy, NB_Energy, ELE_Energy, Temperature: %d, " % it)
        vals = [float(x) for x in ln.split(":")[1].split(",")[1:]]
        assert np.allclose(vals, e, rtol=2e-7)          # the file carries 8 significant digits
    # restart file written at istep 0 and 100 -> holds the state after 150 steps? no: after istep=100 block
    tok = open(os.path.join(d, "new.crd")).read().split()
    nbp = int(tok[0]); vals = np.array(tok[1 + nbp:], dtype=float).reshape(nbp, 2, -1)
    co, vo = o.get_state()
    want = co.reshape(s.n, 4, -1)[:, 0]
    assert np.allclose(vals[:s.n, 0], want, atol=6e-5)   # %10.4f
    first = open(os.path.join(d, "t.trj")).read().split()[0]
    assert int(first) == (s.n + 3) * 63
