"""GC90 (BASELINE configs[0]) step rate on one GPU: 100-step ntsync blocks incl. pair list + merge."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("msc-project_b200")
s = m.make_gc90()
g = m.Engine(s)
for random_mode in (0, 1):
    g.set_state(s.crd, s.vel); g.set_rng(1000000000, 1, 0)
    g.build_pairlist(); g.step(100, random_mode); g.merge(); g.sync()
    t0 = time.perf_counter()
    blocks = 10
    for _ in range(blocks):
        g.build_pairlist(); g.step(100, random_mode); g.merge()
    g.sync()
    dt = time.perf_counter() - t0
    print(f"GC90 random={random_mode}: {blocks*100/dt:.0f} steps/s ({dt/blocks/100*1e6:.1f} us/step incl. pair list + merge every 100 steps), pairs {g.num_pairs}")
