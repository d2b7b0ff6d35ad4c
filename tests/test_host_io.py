"""CPU tests of the C++ host layer's file formats (class IO in msc-project_b200/host) against the
reference's own, unmodified IO class (oracle/_ref/libedmd_refio.so) and against committed golden
output files.  No GPU needed: nothing here calls a compute entry point."""
import ctypes as C
import filecmp
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "deck")
PD = C.POINTER(C.c_double)


@pytest.fixture(scope="session")
def hostlib():
    subprocess.run(["make", "-C", os.path.join(ROOT, "msc-project_b200", "host")], check=True, stdout=subprocess.DEVNULL)
    out = os.path.join(ROOT, "tests", "cpp", "libhost_capi.so")
    src = os.path.join(ROOT, "tests", "cpp", "host_capi.cpp")
    lib_dir = os.path.join(ROOT, "msc-project_b200")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out, src, "-I" + os.path.join(lib_dir, "host"),
                    "-L" + lib_dir, "-ledmd_host", "-ledmd_b200", "-Wl,-rpath," + lib_dir], check=True)
    return C.CDLL(out)


def _load(lib, prefix, deck_dir):
    rc = getattr(lib, prefix + "_load")(deck_dir.encode())
    assert rc == 0
    counts = (C.c_int * 8)(); getattr(lib, prefix + "_counts")(counts)
    params = (C.c_double * 8)(); getattr(lib, prefix + "_params")(params)
    n = counts[0]
    tets = []
    for t in range(n):
        na, nev = C.c_int(), C.c_int()
        getattr(lib, prefix + "_tetrad_sizes")(t, C.byref(na), C.byref(nev))
        n3 = 3 * na.value
        arrs = [np.zeros(n3), np.zeros(n3), np.zeros(n3), np.zeros(nev.value), np.zeros(nev.value * n3), np.zeros(n3), np.zeros(n3)]
        getattr(lib, prefix + "_tetrad")(t, *[a.ctypes.data_as(PD) for a in arrs])
        tets.append((na.value, nev.value, arrs))
    displ = getattr(lib, prefix + "_displ")
    return list(counts), np.array(params), tets, [displ(b) for b in range(counts[1] + 1)]


def _make_deck(edmd, d, irest=0):
    fm = __import__("importlib").import_module("msc-project_b200.formats")
    s = edmd.make_ring(20, kind="coil", laps=2, lap_gap=9.0, nev=5)
    rng = np.random.default_rng(4)
    s.vel[:] = rng.normal(size=s.vel.size)
    os.makedirs(d, exist_ok=True)
    fm.write_prm(os.path.join(d, "t.prm"), s)
    fm.write_crd(os.path.join(d, "t.crd"), s)
    fm.write_crd(os.path.join(d, "restart.crd"), s, with_velocities=True)
    fm.write_config(os.path.join(d, "config.txt"), s, "./t.prm", "./t.crd", "./e.eng", "./t.trj",
                    "./restart.crd" if irest else "./new.crd", irest=irest, nsteps=300, ntsync=50, ntwt=70, ntpr=120)
    return s


def test_deck_reader_matches_reference_io(edmd, hostlib, have_ref, tmp_path, irest=0):
    """irest = 1 cannot be compared: the reference's read_Cofig deletes new_Crd_File (io.cpp:114)
    before read_Crd tries to open it (io.cpp:193) and exits — see test_restart_deck below."""
    if not have_ref:
        pytest.skip("oracle/_ref/libedmd_refio.so not built (needs /root/reference)")
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libedmd_refio.so"))
    for lib in (ref, hostlib):
        for f in ("refio_displ", "hostio_displ"):
            if hasattr(lib, f):
                getattr(lib, f).restype = C.c_int
    d1, d2 = str(tmp_path / "a"), str(tmp_path / "b")
    s = _make_deck(edmd, d1, irest)
    shutil.copytree(d1, d2)   # read_Cofig deletes old outputs: give each reader its own copy
    a = _load(ref, "refio", d1)
    b = _load(hostlib, "hostio", d2)
    assert a[0] == b[0] and a[3] == b[3]
    assert np.array_equal(a[1], b[1])
    for (na1, ne1, x1), (na2, ne2, x2) in zip(a[2], b[2]):
        assert (na1, ne1) == (na2, ne2)
        for u, v in zip(x1, x2):
            assert np.array_equal(u, v)
    # the deck round-trips bit-exactly through the text format
    assert np.array_equal(np.concatenate([t[2][5] for t in b[2]]), s.crd)
    if irest:
        assert np.array_equal(np.concatenate([t[2][6] for t in b[2]]), s.vel)
    # writers: same bytes
    rng = np.random.default_rng(9)
    tot = a[0][2] * 3
    crd = rng.normal(scale=40, size=tot); vel = rng.normal(size=tot)
    e = np.array([655.53997, 632866.04, 0.0, 19.1986562])
    idx = a[3][a[0][1] - 3]
    for lib, pre, d in ((ref, "refio", d1), (hostlib, "hostio", d2)):
        for istep in (0, 50):
            getattr(lib, pre + "_write")(d.encode(), istep, e.ctypes.data_as(PD), idx, crd.ctypes.data_as(PD), vel.ctypes.data_as(PD))
    out = "restart.crd" if irest else "new.crd"
    for f in ("e.eng", "t.trj", out):
        assert filecmp.cmp(os.path.join(d1, f), os.path.join(d2, f), shallow=False), f


def test_restart_deck(edmd, hostlib, tmp_path):
    """irest != 0: coordinates and velocities interleaved per base pair (io.cpp:218-222)."""
    d = str(tmp_path / "deck")
    s = _make_deck(edmd, d, irest=1)
    b = _load(hostlib, "hostio", d)
    assert b[0][3] == 1
    assert np.array_equal(np.concatenate([t[2][5] for t in b[2]]), s.crd)
    # velocities of overlapping copies come from one per-base-pair record: compare per base pair
    got = np.concatenate([t[2][6] for t in b[2]]).reshape(s.n, 4, -1)
    want = s.vel.reshape(s.n, 4, -1)
    assert np.array_equal(got[:, 0], want[:, 0])
    assert os.path.exists(os.path.join(d, "restart.crd"))     # kept (deliberate deviation, see io.cpp)


def test_python_deck_reader_matches_cpp(edmd, hostlib, tmp_path):
    fm = __import__("importlib").import_module("msc-project_b200.formats")
    d = str(tmp_path / "deck")
    s = _make_deck(edmd, d, 0)
    s2, cfg = fm.read_deck(os.path.join(d, "config.txt"))
    b = _load(hostlib, "hostio", d)
    assert np.array_equal(s2.params, b[1]) and np.allclose(s2.params, s.params, rtol=1e-15)
    assert np.array_equal(s2.crd, s.crd)
    assert cfg["nsteps"] == 300 and cfg["ntsync"] == 50


def test_writers_match_golden_files(edmd, hostlib, tmp_path):
    """Golden outputs written by the reference's IO class (tests/golden/make_golden_deck.py)."""
    d = str(tmp_path / "deck")
    shutil.copytree(os.path.join(GOLD, "in"), d)
    b = _load(hostlib, "hostio", d)
    z = np.load(os.path.join(GOLD, "arrays.npz"))
    idx = b[3][b[0][1] - 3]
    for istep in (0, 50):
        hostlib.hostio_write(d.encode(), istep, z["e"].ctypes.data_as(PD), idx, z["crd"].ctypes.data_as(PD), z["vel"].ctypes.data_as(PD))
    for f in ("e.eng", "t.trj", "new.crd"):
        assert filecmp.cmp(os.path.join(d, f), os.path.join(GOLD, "out", f), shallow=False), f
