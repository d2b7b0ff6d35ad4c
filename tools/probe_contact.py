"""Per-phase step times on contact-bearing systems (compact coils), with and without electrostatics.
    python tools/probe_contact.py [n_tetrads laps ...]"""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("msc-project_b200")

def run(n, kind, laps, qfac, steps=20, random_mode=0, lap_gap=40.0):
    s = pkg.make_ring(n, kind=kind, laps=laps, lap_gap=lap_gap)
    g = pkg.Engine(s)
    g.set_nb_consts(100.0, qfac)
    npairs = g.build_pairlist()
    g.set_rng(1, 1, 0)
    crd0, vel0 = s.crd.copy(), s.vel.copy()
    g.step(3, random_mode); g.sync()
    g.set_state(crd0, vel0)
    g.mark(0); g.step(steps, random_mode); g.mark(1)
    ms = g.mark_elapsed(0, 1) / steps
    st = g.nb_stats()
    g.set_state(crd0, vel0)
    g.timing(True); g.timing_reset()
    g.step(steps, random_mode); g.sync()
    ph = {}
    for nm, idx in (("ed", g.T_ED), ("rng", g.T_RNG), ("scan", g.T_SCAN), ("finish", g.T_ACC)):
        t, k = g.timing_get(idx)
        ph[nm] = round(t / steps, 4)
    g.timing(False)
    print(json.dumps(dict(n=n, kind=kind, laps=laps, qfac=qfac, pairs=npairs, ms_per_step=round(ms, 4), phases=ph, **st)), flush=True)
    g.close()

if __name__ == "__main__":
    args = [int(x) for x in sys.argv[1:]]
    cases = list(zip(args[0::2], args[1::2])) or [(10000, 2), (100000, 2)]   # (many laps in one annulus = millions of pairs)
    for n, laps in cases:
        run(n, "ring", 1, 0.0)
        for qfac in (0.0, 332.064 / 4.0):
            run(n, "coil", laps, qfac)
