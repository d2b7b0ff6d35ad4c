"""tools/ed_mma_check.py [N_TETRADS] : the tensor-core ED projection (edmd_set_ed_kernel(3), csrc/ed_mma.cu) against the
warp-per-tetrad projection (2) and the CPU oracle — parity, independence of the tetrad grouping, and time.
No torch: numpy + ctypes only, so it starts in a second on a fresh box."""
import importlib, json, os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
pkg = importlib.import_module("msc-project_b200")


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    d = np.abs(b).max()
    return float(np.abs(a - b).max() / (d if d > 0 else 1.0))


def out(**kw):
    print(json.dumps(kw), flush=True)


def ed_pass(s, which, shard=None):
    g = pkg.Engine(s)
    g.set_ed_kernel(which)
    np_ = g.build_pairlist()
    if shard is not None:
        g.set_shard(shard[0], shard[1], 0, np_)
    g.ed_forces()
    f = g.get_forces(); c, _ = g.get_state()
    g.close()
    return f["ed"], f["ed_energy"], c


def small_cases():
    import pyoracle
    cases = [("coil120", pkg.make_ring(120, kind="coil", laps=3, lap_gap=9.0)), ("gc90", pkg.make_gc90()),
             ("ragged_s224", pkg.make_ragged(40, nevs=(3, 7, 12), bp_lo=50, bp_hi=57, full_tetrad=False)),
             ("ragged_s256", pkg.make_ragged(40, nevs=(3, 7, 12))),
             ("ring1010", pkg.make_ring(1010, kind="ring"))]
    for name, s in cases:
        try:
            o = pyoracle.Oracle(s, "port"); o.build_pairs(); o.forces(random_on=False)
            fo = o.get_forces(); co, _ = o.get_state(); o.close()
            r = {}
            for which in (2, 3):
                fe, ee, c = ed_pass(s, which)
                r[which] = (fe, ee, c)
                out(case=name, kernel=which, n=s.n, nev=int(np.max(s.num_evecs)), atoms=int(np.max(s.num_atoms)),
                    vs_oracle=dict(ed=rel(fe, fo["ed"]), e_ed=rel(ee, fo["ed_energy"]), crd=rel(c, co)),
                    finite=bool(np.isfinite(fe).all() and np.isfinite(c).all()))
            out(case=name, k3_vs_k2=dict(ed=rel(r[3][0], r[2][0]), e_ed=rel(r[3][1], r[2][1]), crd=rel(r[3][2], r[2][2])))
        except Exception as ex:
            out(case=name, error=repr(ex), tb=traceback.format_exc()[-600:])


def big_case(n):
    s = pkg.make_ring(n, kind="ring")
    res = {}
    variants = [(2, None)] + [(3, m) for m in os.environ.get("CHECK_PREFETCH_MODES", "1").split(",")]
    for which, pf in variants:
        if pf is not None:
            os.environ["EDMD_ED_MMA_PREFETCH"] = pf
        g = pkg.Engine(s); g.set_ed_kernel(which); g.set_rng(1000000000, 1, 0)
        npairs = g.build_pairlist()
        g.ed_forces()
        f = g.get_forces(); c, _ = g.get_state()
        res[which] = (f["ed"], f["ed_energy"], c)
        g.set_state(s.crd, s.vel)
        g.timing(True); g.timing_reset()
        for _ in range(10):
            g.ed_forces()
        g.sync()
        ed_ms = g.timing_get(g.T_ED)[0] / 10; sup_ms = g.timing_get(g.T_SUP)[0] / 10
        g.timing(False)
        g.set_state(s.crd, s.vel)
        g.step(5, 1); g.sync()
        g.mark(0); g.step(50, 1); g.mark(1)
        step_ms = g.mark_elapsed(0, 1) / 50
        out(case=f"ring{n}", kernel=which, prefetch=pf, pairs=npairs, ed_ms=ed_ms, superpose_ms=sup_ms, project_ms=ed_ms - sup_ms,
            step_ms_random_on=step_ms)
        g.close()
    out(case=f"ring{n}", k3_vs_k2=dict(ed=rel(res[3][0], res[2][0]), e_ed=rel(res[3][1], res[2][1]), crd=rel(res[3][2], res[2][2])),
        finite=bool(np.isfinite(res[3][0]).all() and np.isfinite(res[3][2]).all()))
    # a tetrad's result must not depend on which tetrads share its MMA group: a shifted shard regroups everything
    t0, t1 = 3, n - 5
    fe, ee, c = ed_pass(s, 3, shard=(t0, t1))
    w = 3 * 252
    a, b = res[3][0].reshape(n, w), fe.reshape(n, w)
    ca, cb = res[3][2].reshape(n, w), c.reshape(n, w)
    out(case=f"ring{n}", regrouped_bitwise=dict(ed=bool(np.array_equal(a[t0:t1], b[t0:t1])), crd=bool(np.array_equal(ca[t0:t1], cb[t0:t1])),
                                                e_ed=bool(np.array_equal(res[3][1][t0:t1], ee[t0:t1]))))


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    t0 = time.time()
    small_cases()
    try:
        big_case(n)
    except Exception as ex:
        out(case="big", error=repr(ex), tb=traceback.format_exc()[-800:])
    out(done=True, wall_s=round(time.time() - t0, 1))
