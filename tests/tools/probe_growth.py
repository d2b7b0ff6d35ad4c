"""Error growth of the GPU trajectory against the CPU oracle, step by step (test tooling: the only
places that may use oracle/ are tests/, smoke() and bench.py's CPU baseline)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import pyoracle as po
m = importlib.import_module("msc-project_b200")

def rel(a, b): return np.abs(a - b).max() / np.abs(b).max()
which = sys.argv[1] if len(sys.argv) > 1 else "coil"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
s = m.make_gc90() if which == "gc90" else m.make_ring(60, kind="coil", laps=2, lap_gap=9.0)
o = po.Oracle(s, "port"); o.build_pairs()
g = m.Engine(s); g.build_pairlist()
for k in range(nsteps):
    o.step(1); g.step(1)
    (co, vo), (cg, vg) = o.get_state(), g.get_state()
    fo, fg = o.get_forces(), g.get_forces()
    print(k, "crd %.2e vel %.2e ed %.2e nb %.2e T %.2e | vmax %.3g Tmean %.4g nbdiff %d" % (
        rel(cg, co), rel(vg, vo), rel(fg["ed"], fo["ed"]), rel(fg["nb"], fo["nb"]) if np.abs(fo["nb"]).max() > 0 else 0,
        rel(fg["temperature"], fo["temperature"]), np.abs(vo).max(), fo["temperature"].mean(),
        np.count_nonzero(fg["nb"] != fo["nb"])))
