"""python tools/prof_step.py N kind laps lap_gap qfac random_mode steps : a few plain (un-graphed) steps for ncu."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("msc-project_b200")
a = sys.argv[1:]
n, kind, laps, gap, qfac, rmode, steps = int(a[0]), a[1], int(a[2]), float(a[3]), float(a[4]), int(a[5]), int(a[6])
s = pkg.make_ring(n, kind=kind, laps=laps, lap_gap=gap)
g = pkg.Engine(s, device=0, pinned=False)
g.set_nb_consts(100.0, qfac)
g.set_graph(False)
print("pairs", g.build_pairlist(), flush=True)
g.set_rng(1000000000, 1, 0)
g.step(steps, rmode)
g.sync()
print("stats", g.nb_stats(), "launches", g.launch_count, flush=True)
g.close()
