"""GPU parity: the CUDA path through the C ABI vs the CPU oracle on identical inputs.

Tolerance (BASELINE.json north_star): 1e-10 relative in double precision on per-tetrad forces
and on short trajectories.  "Relative" is taken per array against its largest magnitude
(max |a-b| / max |b|), the only meaningful reading for force components that pass through 0.
NB forces and random terms are additionally required to be BIT-IDENTICAL: the accumulate kernel
and the RNG kernel reproduce the reference's expression trees and summation order.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-10


def rel(a, b):
    d = np.max(np.abs(a - b)) if a.size else 0.0
    s = np.max(np.abs(b)) if b.size else 0.0
    return d / s if s > 0 else d


@pytest.fixture(scope="module")
def coil(edmd):
    return edmd.make_ring(60, kind="coil", laps=2, lap_gap=9.0)


@pytest.fixture(scope="module")
def gc90(edmd):
    return edmd.make_gc90()


def test_ed_forces_gc90(edmd, po, gc90):
    o = po.Oracle(gc90, "port"); o.ed_forces()
    g = edmd.Engine(gc90); g.ed_forces()
    fo, fg = o.get_forces(), g.get_forces()
    assert rel(fg["ed"], fo["ed"]) < TOL
    assert rel(fg["ed_energy"], fo["ed_energy"]) < TOL
    co, _ = o.get_state(); cg, _ = g.get_state()
    assert rel(cg, co) < TOL          # shaken coordinates overwrite the input, edmd.cpp:134-136


def test_ed_forces_rotated_input(edmd, po, gc90):
    """QCP must find large rotations: rotate + translate every tetrad rigidly first."""
    rng = np.random.default_rng(5)
    s = gc90.copy()
    x = s.crd.reshape(s.n, 252, 3)
    for t in range(s.n):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        x[t] = x[t] @ q.T + rng.normal(scale=20.0, size=3)
    o = po.Oracle(s, "port"); o.ed_forces()
    g = edmd.Engine(s); g.ed_forces()
    assert rel(g.get_forces()["ed"], o.get_forces()["ed"]) < TOL
    assert rel(g.get_state()[0], o.get_state()[0]) < TOL


def test_pairlist_matches_reference_order(edmd, po, gc90, coil):
    for s, expect in ((gc90, 450), (coil, None)):
        o = po.Oracle(s, "port"); n_o = o.build_pairs()
        g = edmd.Engine(s); n_g = g.build_pairlist()
        assert n_g == n_o
        if expect is not None:
            assert n_g == expect          # SURVEY.md §8a a-PL: 450 pairs for GC90
        assert np.array_equal(g.get_pairlist(), o.get_pairs())


@pytest.mark.parametrize("krep,qfac", [(100.0, 0.0), (0.01, 0.0), (100.0, 332.064 / 4.0)])
def test_nb_forces_bitwise(edmd, po, coil, krep, qfac):
    """krep=100,qfac=0: the reference's literals (edmd.cpp:237,242) — overlaps are mostly clipped.
    krep=0.01: same geometry, forces stay inside the clip so the raw sums are compared.
    qfac=332.064/4: electrostatics on (edmd.cpp:240), every atom pair inside atom_Cutoff counts."""
    o = po.Oracle(coil, "port"); o.set_nb_consts(krep, qfac); o.build_pairs()
    o.forces(random_on=False)                       # ED shake first, then NB on shaken coordinates
    g = edmd.Engine(coil); g.set_nb_consts(krep, qfac); g.build_pairlist()
    g.ed_forces(); g.nb_forces()
    fo, fg = o.get_forces(), g.get_forces()
    assert np.count_nonzero(fo["nb"]) > 100          # the case exercises real overlaps
    if krep == 100.0:
        assert np.any(np.abs(fo["nb"]) == 1.0)       # ... and the clip
    else:
        assert np.count_nonzero(np.abs(fo["nb"]) < 1.0) > 1000
    # shaken coordinates agree to 1e-10, so NB forces agree to a looser tolerance ...
    assert rel(fg["nb"], fo["nb"]) < TOL
    assert rel(fg["nb_energy"], fo["nb_energy"]) < TOL
    # ... and bit-for-bit when both start from the SAME coordinates:
    crd, _ = g.get_state()
    pairs = g.get_pairlist()
    o2 = po.Oracle(coil, "port"); o2.set_nb_consts(krep, qfac); o2.set_state(crd, None)
    acc = np.zeros_like(crd)
    off = coil.off3
    en = np.zeros((coil.n, 2))
    for t1, t2 in pairs:
        o2.nb_forces_pair(int(t1), int(t2))          # EDMD::calculate_NB_Forces on one pair
        f = o2.get_forces()
        for t in (t1, t2):                           # worker.cpp:173-178
            acc[off[t]:off[t + 1]] += f["nb"][off[t]:off[t + 1]]
            en[t] += f["nb_energy"][t]
    ref = np.clip(acc, -1.0, 1.0)                    # master.cpp:304-305
    assert np.array_equal(fg["nb"], ref)
    assert rel(fg["nb_energy"], en) < 1e-12


def test_nb_scan_flags_are_superset(edmd, po, coil):
    g = edmd.Engine(coil); g.build_pairlist(); g.nb_scan(); g.sync()
    flagged = g.flagged_pairs()
    assert 0 < flagged < g.num_pairs


def test_update_velocities_and_coordinates(edmd, po, coil):
    rng = np.random.default_rng(11)
    tot = coil.crd.size
    ed, rnd, nb = (rng.normal(size=tot), rng.normal(size=tot) * 3, rng.uniform(-1, 1, tot))
    vel = rng.normal(size=tot) * 0.05
    o = po.Oracle(coil, "port"); o.set_state(None, vel); o.set_forces(ed, rnd, nb)
    o.update_velocities(); o.update_coordinates()
    g = edmd.Engine(coil); g.set_state(None, vel); g.set_forces(ed, rnd, nb)
    g.update_velocities(); g.update_coordinates()
    (co, vo), (cg, vg) = o.get_state(), g.get_state()
    assert rel(vg, vo) < TOL and rel(cg, co) < TOL
    assert rel(g.get_forces()["temperature"], o.get_forces()["temperature"]) < TOL


def test_random_terms_bitwise(edmd, po, gc90):
    o = po.Oracle(gc90, "port"); o.set_time(1000000000); o.set_rng_calls(0)
    o.random_terms(rank=1)
    g = edmd.Engine(gc90); g.random_terms(mode=1, time0=1000000000, rank=1, calls_before=0)
    ro, rg = o.get_forces()["rnd"], g.get_forces()["rnd"]
    assert np.array_equal(rg, ro)
    # second call continues the seed counter (edmd.cpp:173,179)
    o.random_terms(rank=1)
    g.random_terms(mode=1, time0=1000000000, rank=1, calls_before=gc90.n)
    assert np.array_equal(g.get_forces()["rnd"], o.get_forces()["rnd"])


def test_merge_bitwise(edmd, po, coil):
    rng = np.random.default_rng(3)
    crd = coil.crd + rng.normal(scale=0.1, size=coil.crd.size)
    vel = rng.normal(size=coil.crd.size)
    o = po.Oracle(coil, "port"); o.set_state(crd, vel); o.merge()
    g = edmd.Engine(coil); g.set_state(crd, vel); g.merge()
    (co, vo), (cg, vg) = o.get_state(), g.get_state()
    assert np.array_equal(cg, co) and np.array_equal(vg, vo)
    g.merge()                                    # idempotent: averaging equal copies
    cg2, vg2 = g.get_state()
    assert rel(cg2, cg) < 1e-15 and rel(vg2, vg) < 1e-15


@pytest.mark.parametrize("random_on", [False, True])
def test_trajectory(edmd, po, coil, random_on):
    """N steps of the full protocol, then merge; error growth is reported, bound at 10 steps."""
    o = po.Oracle(coil, "port"); o.set_time(1000000000); o.set_rng_calls(0); o.build_pairs()
    g = edmd.Engine(coil); g.build_pairlist(); g.set_rng(1000000000, 1, 0)
    worst = 0.0
    for k in range(10):
        o.step(1, random_on=random_on)
        g.step(1, random_mode=1 if random_on else 0)
        (co, vo), (cg, vg) = o.get_state(), g.get_state()
        worst = max(worst, rel(cg, co), rel(vg, vo))
    # The 60-tetrad double coil is deliberately violent (interpenetrating helices, forces at the
    # clip, stiff soft-core); measured growth: 1e-12 after 10 steps (tests/tools/probe_growth.py).
    assert worst < TOL
    o.merge(); g.merge()
    assert rel(g.get_state()[0], o.get_state()[0]) < TOL
    eo, eg = o.energies(), g.energies()
    assert np.allclose(eg, eo, rtol=1e-9, atol=1e-12)


def test_gc90_shipped_config_block(edmd, po, gc90):
    """config[0]: GC90, shipped cutoffs, 20 steps + merge (random off)."""
    o = po.Oracle(gc90, "port"); assert o.build_pairs() == 450
    g = edmd.Engine(gc90); assert g.build_pairlist() == 450
    o.step(20); g.step(20)
    o.merge(); g.merge()
    (co, vo), (cg, vg) = o.get_state(), g.get_state()
    assert rel(cg, co) < TOL and rel(vg, vo) < TOL
    assert np.allclose(g.energies(), o.energies(), rtol=1e-10)


@pytest.mark.parametrize("target", [0, 1, 2 ** 31 - 1, 2 ** 31, 2 ** 31 + 5, 2 ** 32 - 1])
def test_random_terms_edge_seeds(edmd, po, target):
    """srand() edge cases: seed 0 (glibc substitutes 1), 2^31-1 (every seed word 0), and seeds whose int32 view is
    negative (C's truncating division in the seeding recurrence) — the kernel's closed-form seeding must give the
    reference stream for all of them."""
    s = edmd.make_gc90().subset(range(4))
    time0 = target - 13580                      # first call: seed = (unsigned)(13579 + rank^3 + time0), rank 1
    o = po.Oracle(s, "port"); o.set_time(time0); o.set_rng_calls(0); o.random_terms(rank=1)
    g = edmd.Engine(s); g.random_terms(mode=1, time0=time0, rank=1, calls_before=0)
    assert np.array_equal(g.get_forces()["rnd"], o.get_forces()["rnd"])
