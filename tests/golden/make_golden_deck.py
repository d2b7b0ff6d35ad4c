"""Golden deck + outputs written by the reference's UNMODIFIED IO class (oracle/_ref/libedmd_refio.so).
Run in the build container:  python tests/golden/make_golden_deck.py"""
import ctypes as C, importlib, os, shutil, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("msc-project_b200"); fm = importlib.import_module("msc-project_b200.formats")
PD = C.POINTER(C.c_double)
gold = os.path.join(ROOT, "tests", "golden", "deck")
shutil.rmtree(gold, ignore_errors=True)
din, dout = os.path.join(gold, "in"), os.path.join(gold, "out")
os.makedirs(din); os.makedirs(dout)
s = pkg.make_ring(20, kind="coil", laps=2, lap_gap=9.0, nev=3)
fm.write_prm(os.path.join(din, "t.prm"), s); fm.write_crd(os.path.join(din, "t.crd"), s)
fm.write_config(os.path.join(din, "config.txt"), s, "./t.prm", "./t.crd", "./e.eng", "./t.trj", "./new.crd",
                nsteps=300, ntsync=50, ntwt=70, ntpr=120)
work = os.path.join(gold, "_work"); shutil.copytree(din, work)
ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libedmd_refio.so"))
assert ref.refio_load(work.encode()) == 0
counts = (C.c_int * 8)(); ref.refio_counts(counts)
rng = np.random.default_rng(9)
tot = counts[2] * 3
crd = rng.normal(scale=40, size=tot); vel = rng.normal(size=tot); e = np.array([655.53997, 632866.04, 0.0, 19.1986562])
ref.refio_displ.restype = C.c_int
idx = ref.refio_displ(counts[1] - 3)
for istep in (0, 50):
    ref.refio_write(work.encode(), istep, e.ctypes.data_as(PD), idx, crd.ctypes.data_as(PD), vel.ctypes.data_as(PD))
for f in ("e.eng", "t.trj", "new.crd"):
    shutil.copy(os.path.join(work, f), os.path.join(dout, f))
shutil.rmtree(work)
np.savez_compressed(os.path.join(gold, "arrays.npz"), crd=crd, vel=vel, e=e)
print("golden deck written to", gold)
