"""Quick on-GPU probe: per-kernel device times on the C1 workload (1000 tetrads, all pairs)."""
import importlib, os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
m = importlib.import_module("msc-project_b200")
from importlib import import_module
eng = import_module("msc-project_b200.engine")

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
cut = float(sys.argv[2]) if len(sys.argv) > 2 else 1e9
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 1
tf, mhz = eng.measure_fp64_peak(0)
print(f"fp64 peak {tf:.2f} TFLOP/s at ~{mhz:.0f} MHz")
s = m.make_ring(n, kind="ring", mole_cutoff=cut)
t0 = time.time(); g = m.Engine(s); print("upload", time.time() - t0)
t0 = time.time(); npairs = g.build_pairlist(); print("pairs", npairs, time.time() - t0)
g.set_nb_scan_mode(mode)
g.step(2); g.sync()
g.timing(True); g.timing_reset()
t0 = time.time(); g.step(steps); g.sync(); wall = time.time() - t0
names = ["ed", "scan", "acc", "rng", "merge", "pairs", "int", "layout"]
out = {}
for i, nm in enumerate(names):
    ms, k = g.timing_get(i)
    if k: out[nm] = ms / k
print("per-launch ms:", json.dumps(out))
scan = out["scan"]
flops = npairs * 63504 * 9
print(f"mode {mode}: wall/step {wall/steps*1e3:.3f} ms; scan {scan:.3f} ms -> {flops/scan/1e9:.2f} TFLOP/s algorithmic = {flops/scan/1e9/tf*100:.1f}% of measured FP64 peak; pairs/s {npairs/scan*1e3:.3e}")
