"""GPU edge cases and size-independent properties.

* ragged system: base pairs of 58..64 atoms (one tetrad at the 256-atom engine maximum), every tetrad
  with its own parameter set, 3 / 12 / 20 / 29 eigenvectors (below, at and above the ED kernel's
  12-per-chunk register tile);
* empty pair list, single flagged pair, engine limits;
* properties at BASELINE sizes (C1: 1,000 tetrads all-pairs; 10^4 tetrads, random on): scan flags are a
  superset of the true hits and both arithmetic forms agree, determinism (bit-stable reruns),
  zero NB force on the clash-free ring, merge idempotence, sampled-tetrad parity with the oracle.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def ragged(edmd):
    return edmd.make_ragged(40)


def test_ragged_per_function_parity(edmd, po, ragged):
    s = ragged
    assert s.num_atoms.max() == 256 and s.num_atoms.min() < 250
    o = po.Oracle(s, "port"); g = edmd.Engine(s)
    assert g.build_pairlist() == o.build_pairs()
    assert np.array_equal(g.get_pairlist(), o.get_pairs())
    o.forces(random_on=False); g.ed_forces(); g.nb_forces()
    fo, fg = o.get_forces(), g.get_forces()
    assert rel(fg["ed"], fo["ed"]) < TOL and rel(fg["ed_energy"], fo["ed_energy"]) < TOL
    assert np.count_nonzero(fo["nb"]) > 100
    assert rel(fg["nb"], fo["nb"]) < TOL and rel(fg["nb_energy"], fo["nb_energy"]) < TOL
    o.update_velocities(); o.update_coordinates(); g.update_velocities(); g.update_coordinates()
    (co, vo), (cg, vg) = o.get_state(), g.get_state()
    assert rel(cg, co) < TOL and rel(vg, vo) < TOL
    o.merge(); g.merge()
    (co, vo), (cg, vg) = o.get_state(), g.get_state()
    assert rel(cg, co) < TOL and rel(vg, vo) < TOL


def test_ragged_stochastic_trajectory(edmd, po, ragged):
    s = ragged
    o = po.Oracle(s, "port"); o.set_time(77); o.set_rng_calls(5); o.build_pairs()
    g = edmd.Engine(s); g.build_pairlist(); g.set_rng(77, 1, 5)
    o.step(6, random_on=True); g.step(6, random_mode=1)
    assert np.array_equal(g.get_forces()["rnd"], o.get_forces()["rnd"])
    (co, vo), (cg, vg) = o.get_state(), g.get_state()
    assert rel(cg, co) < TOL and rel(vg, vo) < TOL


def test_both_scan_modes_give_identical_results(edmd, ragged):
    out = []
    for mode in (0, 1, 2, 4):
        g = edmd.Engine(ragged); g.set_nb_scan_mode(mode); g.build_pairlist(); g.step(4)
        out.append(g.get_state() + (g.flagged_pairs(),))
    for k in (1, 2, 3):
        assert np.array_equal(out[0][0], out[k][0]) and np.array_equal(out[0][1], out[k][1])
        assert out[k][2] >= out[0][2] >= 1           # filters are supersets of the direct FP64 test


def test_empty_and_tiny_pair_lists(edmd, po, ragged):
    s = ragged
    g = edmd.Engine(s); g.set_pairlist(np.zeros((0, 2), dtype=np.int32))
    g.step(2)
    o = po.Oracle(s, "port"); o.set_pairs(np.zeros((0, 2), dtype=np.int32)); o.step(2)
    assert rel(g.get_state()[0], o.get_state()[0]) < TOL
    assert not np.any(g.get_forces()["nb"])
    one = np.array([[9, 3]], dtype=np.int32)            # a single pair, "odd" orientation
    g.set_pairlist(one); o.set_pairs(one)
    g.ed_forces(); g.nb_forces(); o.forces()
    assert rel(g.get_forces()["nb"], o.get_forces()["nb"]) < TOL or not np.any(o.get_forces()["nb"])


def test_engine_limits_are_reported(edmd):
    s = edmd.make_ring(20)
    big = edmd.ParamSet(300, 2, np.zeros(900), np.ones(900), np.zeros(900), np.ones(2), np.zeros((2, 900)))
    s2 = edmd.EdmdSystem([big], np.zeros(4, dtype=np.int32), None, np.zeros(4 * 900), np.zeros(4 * 900))
    with pytest.raises(edmd.EngineError, match="limit"):
        edmd.Engine(s2)
    g = edmd.Engine(s)
    with pytest.raises(edmd.EngineError, match="pair list"):
        g.step(1)                                       # no pair list yet
    with pytest.raises(edmd.EngineError):
        g.set_pairlist(np.array([[0, 99]], dtype=np.int32))


def test_bit_stable_reruns(edmd):
    s = edmd.make_ring(200, kind="coil", laps=4, lap_gap=9.0)
    res = []
    for _ in range(3):
        g = edmd.Engine(s); g.build_pairlist(); g.set_rng(1, 1, 0); g.step(8, random_mode=1); g.merge()
        res.append(g.get_state())
    for c, v in res[1:]:
        assert np.array_equal(c, res[0][0]) and np.array_equal(v, res[0][1])


def test_c1_full_size_properties(edmd, po):
    """BASELINE configs[1]: 1,000 tetrads, all pairs (494,500), random off."""
    s = edmd.make_ring(1000, kind="ring", mole_cutoff=1e9)
    g = edmd.Engine(s)
    assert g.build_pairlist() == 494500
    p = g.get_pairlist()
    sep = np.abs(p[:, 0] - p[:, 1])
    assert sep.min() == 6 and sep.max() == 994                      # master.cpp:194-195
    g.step(3)
    f = g.get_forces()
    assert not np.any(f["nb"]) and not np.any(f["nb_energy"])      # clash-free ring: exactly zero
    assert g.flagged_pairs() == 0
    # sampled tetrads against the oracle: the ED / integrate path of those tetrads does not depend on
    # the others when NB forces vanish
    pick = [0, 1, 499, 998, 999]
    sub = s.subset(pick)
    o = po.Oracle(sub, "port"); o.set_pairs(np.zeros((0, 2), dtype=np.int32)); o.step(3)
    co, vo = o.get_state(); cg, vg = g.get_state()
    off = s.off3
    idx = np.concatenate([np.arange(off[t], off[t + 1]) for t in pick])
    assert rel(cg[idx], co) < TOL and rel(vg[idx], vo) < TOL
    # merge: idempotent at full size
    g.merge(); c1, v1 = g.get_state(); g.merge(); c2, v2 = g.get_state()
    assert rel(c2, c1) < 1e-15 and rel(v2, v1) < 1e-14


def test_c2_size_coil_random_on(edmd, po):
    """10^4-tetrad coil, Langevin random forces on: hit flags vs exact evaluation on a pair sample,
    stochastic stream parity on sampled tetrads."""
    s = edmd.make_ring(10000, kind="coil", laps=50, lap_gap=9.0)
    g = edmd.Engine(s)
    npairs = g.build_pairlist()
    assert npairs > 40000
    g.ed_forces(); g.random_terms(mode=1, time0=5, rank=1, calls_before=0); g.nb_forces()
    crd, _ = g.get_state()
    f = g.get_forces()
    pairs = g.get_pairlist()
    flagged = g.flagged_pairs()
    assert 0 < flagged < npairs
    # exact check of a sample of pairs with the oracle: force on tetrad t from its listed partners
    o = po.Oracle(s.subset([0]), "port")                 # only for its libc-exact RNG below
    picks = [0, 17, 5003, 9999]
    for t in picks:
        sub_ids = sorted(set([t] + [int(b if a == t else a) for a, b in pairs[(pairs[:, 0] == t) | (pairs[:, 1] == t)]]))
        sub = s.subset(sub_ids)
        remap = {u: i for i, u in enumerate(sub_ids)}
        off = s.off3
        sub.crd[:] = np.concatenate([crd[off[u]:off[u + 1]] for u in sub_ids])
        oo = po.Oracle(sub, "port")
        mine = pairs[(pairs[:, 0] == t) | (pairs[:, 1] == t)]
        acc = np.zeros(756)
        for a, b in mine:                                  # pair-list order, worker.cpp:166-179
            oo.nb_forces_pair(remap[int(a)], remap[int(b)])
            ff = oo.get_forces()["nb"]
            k = remap[t]
            acc += ff[756 * k:756 * (k + 1)]
        assert np.array_equal(f["nb"][off[t]:off[t + 1]], np.clip(acc, -1, 1))
    # random terms of tetrad t = call number t of the stream
    for t in picks:
        o.set_time(5); o.set_rng_calls(t)
        o.random_terms(0, rank=1)
        # masses differ per parameter set: compare the normal deviates (rnd / noise)
        ps_o = s.psets[int(s.param_id[0])]; ps_t = s.psets[int(s.param_id[t])]
        z_o = o.get_forces()["rnd"] / np.sqrt(ps_o.masses3)
        z_g = f["rnd"][s.off3[t]:s.off3[t + 1]] / np.sqrt(ps_t.masses3)
        assert np.allclose(z_g, z_o, rtol=1e-14)


@pytest.mark.parametrize("kind,n", [("coil", 600), ("ring", 3000)])
def test_cell_list_builder_equals_sweep_and_oracle(edmd, po, kind, n):
    """The uniform-grid builder must return the reference's list: same pairs, same orientation, same order."""
    s = edmd.make_ring(n, kind=kind, laps=6, lap_gap=9.0)
    g = edmd.Engine(s)
    g.set_pair_builder(1); n1 = g.build_pairlist(); p1 = g.get_pairlist()
    g.set_pair_builder(2); n2 = g.build_pairlist(); p2 = g.get_pairlist()
    assert n1 == n2 and np.array_equal(p1, p2)
    o = po.Oracle(s, "port")
    assert o.build_pairs() == n2 and np.array_equal(o.get_pairs(), p2)


def test_cell_list_builder_large_ring(edmd):
    """2e5 tetrads: auto-selected cell list; every tetrad of the planar ring has the same partner count."""
    s = edmd.make_ring(200000, kind="ring")
    g = edmd.Engine(s)
    npairs = g.build_pairlist()
    p = g.get_pairlist()
    deg = np.bincount(p.reshape(-1), minlength=s.n)
    assert 4 * s.n <= npairs <= 5 * s.n + 5                # separations 6..9 (or ..10) either side
    assert deg.max() - deg.min() <= 2
    sep = np.abs(p[:, 0] - p[:, 1]); sep = np.minimum(sep, s.n - sep)
    assert sep.min() == 6
    assert np.all(np.diff(np.minimum(p[:, 0], p[:, 1])) >= 0)     # rows ascending


def test_step_host_matches_the_three_separate_calls(edmd, ragged):
    """edmd_step_host (pipelined uploads) == edmd_set_state + edmd_step + edmd_get_state, bit for bit,
    starting from the same host Tetrad arrays, over several calls, with and without random forces."""
    for random_mode in (0, 1):
        a = edmd.Engine(ragged); a.build_pairlist(); a.set_rng(123, 1, 0)
        b = edmd.Engine(ragged); b.build_pairlist(); b.set_rng(123, 1, 0)
        for nsteps in (1, 3, 1):
            a.step_host(nsteps, random_mode)                       # host block in -> stepped -> host block
            b.push_state(); b.step(nsteps, random_mode); b.get_state(copy=False)
            assert np.array_equal(a.block.crd, b.block.crd) and np.array_equal(a.block.vel, b.block.vel)
        a.merge(); b.merge()                                        # the device state is consistent afterwards
        (ca, va), (cb, vb) = a.get_state(), b.get_state()
        assert np.array_equal(ca, cb) and np.array_equal(va, vb)


def test_both_projection_kernels_match_the_oracle(edmd, po):
    """ED pass with the thread-per-atom projection (edmd_set_ed_kernel(1)), the warp-per-tetrad projection (2)
    and the tensor-core projection (3, DMMA over groups of eight tetrads) on a system whose tetrads share
    parameter sets and on one where every tetrad has its own (groups of one): all within the parity bound
    of the CPU oracle, and equal to each other up to the order of the 12 projection sums."""
    for s in (edmd.make_ring(120, kind="coil", laps=3, lap_gap=9.0), edmd.make_gc90()):
        o = po.Oracle(s, "port"); o.build_pairs(); o.forces(random_on=False)
        fo = o.get_forces(); co, _ = o.get_state()
        out = []
        for which in (1, 2, 3, 0):
            g = edmd.Engine(s); g.set_ed_kernel(which); g.build_pairlist(); g.ed_forces()
            fg = g.get_forces(); cg, _ = g.get_state()
            assert rel(fg["ed"], fo["ed"]) < TOL and rel(fg["ed_energy"], fo["ed_energy"]) < TOL
            assert rel(cg, co) < TOL
            out.append((fg["ed"], cg))
        assert rel(out[0][0], out[1][0]) < 1e-12 and rel(out[0][1], out[1][1]) < 1e-13
        assert rel(out[0][0], out[2][0]) < 1e-12 and rel(out[0][1], out[2][1]) < 1e-13


def test_tensor_core_projection_is_independent_of_the_grouping(edmd):
    """The tensor-core projection puts eight tetrads of one parameter set into the columns of an MMA tile.  A
    tetrad's result must not depend on which tetrads share its tile (that is what keeps multi-GPU runs
    bit-identical to one GPU): a shard that starts three tetrads later regroups every tile.  Also: 4,100 tetrads
    with shared sets is a system for which 'automatic' picks this kernel."""
    s = edmd.make_ring(4100, kind="ring")
    n, w = s.n, 3 * 252
    res = []
    for which, shard in ((3, None), (3, (3, n - 5)), (0, None), (2, None)):
        g = edmd.Engine(s); g.set_ed_kernel(which)
        npairs = g.build_pairlist()
        if shard is not None:
            g.set_shard(shard[0], shard[1], 0, npairs)
        g.ed_forces()
        f = g.get_forces(); c, _ = g.get_state()
        res.append((f["ed"].reshape(n, w), f["ed_energy"].copy(), c.reshape(n, w)))
        g.close()
    full, part, auto, warp = res
    assert np.array_equal(full[0][3:n - 5], part[0][3:n - 5]) and np.array_equal(full[2][3:n - 5], part[2][3:n - 5])
    assert np.array_equal(full[1][3:n - 5], part[1][3:n - 5])
    assert np.array_equal(full[0], auto[0]) and np.array_equal(full[2], auto[2])
    assert rel(full[0], warp[0]) < 1e-12 and rel(full[2], warp[2]) < 1e-13 and rel(full[1], warp[1]) < 1e-12


def test_small_atom_stride_system(edmd, po):
    """Every tetrad has at most 224 atoms, so the atom stride S is below the engine maximum of 256 while the
    culling groups still address 8 x 32 slots: the slot tables must not depend on S (a round-1 overflow)."""
    s = edmd.make_ragged(40, bp_lo=50, bp_hi=57, full_tetrad=False)
    assert s.num_atoms.max() <= 224
    o = po.Oracle(s, "port"); g = edmd.Engine(s)
    assert g.atom_stride < 256
    assert g.build_pairlist() == o.build_pairs()
    o.forces(random_on=False); g.ed_forces(); g.nb_forces()
    fo, fg = o.get_forces(), g.get_forces()
    assert np.count_nonzero(fo["nb"]) > 50
    assert rel(fg["ed"], fo["ed"]) < TOL and rel(fg["nb"], fo["nb"]) < TOL and rel(fg["nb_energy"], fo["nb_energy"]) < TOL
    out = []
    for mode in (0, 1, 2, 4):
        h = edmd.Engine(s); h.set_nb_scan_mode(mode); h.build_pairlist(); h.set_rng(9, 1, 0); h.step(4, random_mode=1)
        out.append(h.get_state())
    o2 = po.Oracle(s, "port"); o2.set_time(9); o2.set_rng_calls(0); o2.build_pairs(); o2.step(4, random_on=True)
    co, vo = o2.get_state()
    for c, v in out:
        assert np.array_equal(c, out[0][0]) and np.array_equal(v, out[0][1])
    assert rel(out[0][0], co) < TOL and rel(out[0][1], vo) < TOL


@pytest.mark.parametrize("qfac", [0.0, 332.064 / 4.0])
def test_block_masked_force_pass_equals_all_blocks(edmd, po, qfac):
    """The culled distance pass hands the force pass a refined 8 x 8 block mask per pair; the brute-force
    passes hand it all ones.  Forces, energies and trajectories must be bit-identical either way, with and
    without electrostatics (every atom pair inside atom_Cutoff counts when qfac != 0)."""
    s = edmd.make_ring(120, kind="coil", laps=3, lap_gap=9.0)
    res = []
    for mode in (0, 4):
        g = edmd.Engine(s); g.set_nb_consts(100.0, qfac); g.set_nb_scan_mode(mode); g.build_pairlist()
        g.ed_forces(); g.nb_forces()
        f = g.get_forces()
        g.step(3)
        res.append((f["nb"], f["nb_energy"], g.get_state(), g.flagged_pairs()))
    assert np.count_nonzero(res[0][0]) > 100
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert np.array_equal(res[0][2][0], res[1][2][0]) and np.array_equal(res[0][2][1], res[1][2][1])
    assert res[1][3] >= res[0][3] >= 1
    o = po.Oracle(s, "port"); o.set_nb_consts(100.0, qfac); o.build_pairs(); o.forces(random_on=False)
    assert rel(res[1][0], o.get_forces()["nb"]) < TOL


def _adjoint_column_norms(ref, mov):
    """Squared norms of the four quaternion candidates in the order FastCalcRMSDAndRotation tries them
    (qcprot.c:252-309): columns 1, 2, 4, 3 of adj(K - lambda_max I)."""
    a = ref - ref.mean(axis=0); x = mov - mov.mean(axis=0)
    m = a.T @ x
    (xx, xy, xz), (yx, yy, yz), (zx, zy, zz) = m
    k = np.array([[xx + yy + zz, yz - zy, zx - xz, xy - yx],
                  [yz - zy, xx - yy - zz, xy + yx, zx + xz],
                  [zx - xz, xy + yx, -xx + yy - zz, yz + zy],
                  [xy - yx, zx + xz, yz + zy, -xx - yy + zz]])
    lam = np.linalg.eigvalsh(k).max()
    b = k - lam * np.eye(4)
    adj = np.empty((4, 4))
    for i in range(4):
        for j in range(4):
            adj[j, i] = (-1) ** (i + j) * np.linalg.det(np.delete(np.delete(b, i, 0), j, 1))
    return [float(adj[:, c] @ adj[:, c]) for c in (0, 1, 3, 2)]


def test_qcp_fallback_branches(edmd, po):
    """Force every exit of the quaternion fallback chain (qcprot.c:252-309): a structure that is EXACTLY its
    average turned by 180 degrees about x / z / y makes the first one / two / three adjoint columns vanish,
    and a structure collapsed to a point makes all four vanish (identity rotation).  The structures are
    0.05 x DNA-sized so that a vanishing column is far below, and a good one far above, evecprec = 1e-6."""
    base = edmd.make_gc90()
    flips = [np.array([1.0, 1.0, 1.0]), np.array([1.0, -1.0, -1.0]), np.array([-1.0, -1.0, 1.0]),
             np.array([-1.0, 1.0, -1.0]), None]
    psets, crd = [], []
    for k, d in enumerate(flips):
        p = base.psets[k]
        avg = (p.avg.reshape(-1, 3) - p.avg.reshape(-1, 3).mean(axis=0)) * 0.05
        psets.append(edmd.ParamSet(p.num_atoms, p.num_evecs, np.ascontiguousarray(avg.reshape(-1)), p.masses3, p.abq,
                                   p.evals, p.evecs))
        x = avg * d if d is not None else np.tile(np.array([0.3, -0.2, 0.1]), (p.num_atoms, 1))
        crd.append(x.reshape(-1))
        norms = _adjoint_column_norms(avg, x)
        want = {0: 0, 1: 1, 2: 2, 3: 3, 4: 4}[k]       # how many candidates must be rejected
        assert all(v < 1e-12 for v in norms[:want]) and (want == 4 or norms[want] > 1e-3), (k, norms)
    crd = np.concatenate(crd)
    s = edmd.EdmdSystem(psets, np.arange(len(flips), dtype=np.int32), None, crd, np.zeros_like(crd))
    o = po.Oracle(s, "port"); o.ed_forces()
    g = edmd.Engine(s); g.ed_forces()
    fo, fg = o.get_forces(), g.get_forces()
    co, _ = o.get_state(); cg, _ = g.get_state()
    assert np.all(np.isfinite(cg)) and np.all(np.isfinite(fg["ed"]))
    assert np.abs(cg - co).max() < 1e-12 and np.abs(fg["ed"] - fo["ed"]).max() < 1e-12
    assert np.abs(fg["ed_energy"] - fo["ed_energy"]).max() < 1e-12


def test_c3_full_size_sampled_parity(edmd, po):
    """BASELINE configs[3]: 100,000 tetrads, cutoff pair list (cell-list builder), Langevin random forces on.
    The ring is clash-free, so every tetrad evolves independently of the others and sampled tetrads can be
    checked against the oracle stepping those tetrads alone with the same random-stream call numbers
    (call number = step * n + tetrad, edmd.cpp:173,179)."""
    n = 100000
    s = edmd.make_ring(n, kind="ring")
    g = edmd.Engine(s)
    npairs = g.build_pairlist()
    assert 4 * n <= npairs <= 5 * n + 5
    g.set_rng(424242, 1, 0)
    steps = 3
    g.step(steps, random_mode=1)
    assert g.flagged_pairs() == 0
    f = g.get_forces()
    assert not np.any(f["nb"])
    cg, vg = g.get_state()
    pick = [0, 1, 31337, 50000, 99998, 99999]
    sub = s.subset(pick)
    o = po.Oracle(sub, "port"); o.set_time(424242)
    for k in range(steps):
        for i, t in enumerate(pick):
            o.set_rng_calls(k * n + t)
            o.ed_forces(i); o.random_terms(i, rank=1)
            o.update_velocities(i); o.update_coordinates(i)
    co, vo = o.get_state()
    off = s.off3
    idx = np.concatenate([np.arange(off[t], off[t + 1]) for t in pick])
    assert np.array_equal(f["rnd"][idx], o.get_forces()["rnd"])
    assert rel(cg[idx], co) < TOL and rel(vg[idx], vo) < TOL
    g.merge(); c1, v1 = g.get_state(); g.merge(); c2, v2 = g.get_state()
    assert rel(c2, c1) < 1e-15 and rel(v2, v1) < 1e-14
