"""CPU tests that pin the oracle (no GPU needed).

1. port (oracle/edmd_oracle.c) == reference build (unmodified reference sources), bit for bit,
   when oracle/_ref is present (it is built wherever /root/reference exists);
2. port == tests/golden/*.npz (generated from the reference build by tests/golden/make_golden.py);
3. the reference's own known answers: gamfac = 0.9960 (edmd.cpp:298) and
   sum(noise_Factor) = 3594.75 at gamma = 2.0 (edmd.cpp:182) within the 0.1 % the missing
   .prm's masses allow;
4. glibc rand()/srand() restatement == this machine's libc;
5. mathematical invariants: QCP rotations, Newton's third law, merge idempotence.
"""
import ctypes
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STRIDE = 7
T0 = 1000000000


def _run_like_golden(po, system, kind, rng_calls_before):
    out = {}
    o = po.Oracle(system, kind)
    out["npairs"] = np.array([o.build_pairs()]); out["pairs"] = o.get_pairs()
    o.forces(random_on=False)
    f = o.get_forces(); c, v = o.get_state()
    out.update(f1_ed=f["ed"], f1_ed_energy=f["ed_energy"], f1_nb=f["nb"], f1_nb_energy=f["nb_energy"], f1_crd=c)
    o.update_velocities(); o.update_coordinates()
    c, v = o.get_state()
    out.update(s1_crd=c, s1_vel=v, s1_temp=o.get_forces()["temperature"])
    o.step(4); o.merge()
    c, v = o.get_state()
    out.update(s5m_crd=c, s5m_vel=v, s5m_energies=o.energies())
    o2 = po.Oracle(system, kind); o2.set_time(T0)
    if kind == "port":
        o2.set_rng_calls(rng_calls_before)
    o2.build_pairs(); o2.step(3, random_on=True)
    c, v = o2.get_state()
    out.update(r3_crd=c, r3_vel=v, r3_rnd=o2.get_forces()["rnd"])
    return out


@pytest.mark.parametrize("name", ["gc90", "coil40"])
def test_port_matches_golden_vectors(edmd, po, name):
    system = edmd.make_gc90() if name == "gc90" else edmd.make_ring(40, kind="coil", laps=2, lap_gap=9.0)
    gold = np.load(os.path.join(GOLD, f"golden_{name}.npz"))
    got = _run_like_golden(po, system, "port", int(gold["rng_calls_before"][0]))
    for k, v in got.items():
        v = np.asarray(v)
        if v.ndim == 1 and v.size > 1000:
            assert np.array_equal(v[::STRIDE], gold[k]), k
            assert v.sum() == gold[k + "_sum"][0], k
        else:
            assert np.array_equal(v, gold[k]), k


def test_port_is_bit_identical_to_reference_build(edmd, po, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    system = edmd.make_ring(40, kind="coil", laps=2, lap_gap=12.0)
    res = {}
    for kind in ("port", "reference"):
        o = po.Oracle(system, kind); o.set_time(T0)
        calls = o.rng_calls()
        if kind == "reference":
            ref_calls = calls
        o.build_pairs()
        res[kind] = (o, calls)
    # align the port's RNG_Seed counter with however many calls the reference library has made
    res["port"][0].set_rng_calls(ref_calls)
    outs = {}
    for kind in ("port", "reference"):
        o = res[kind][0]
        o.step(3, random_on=True); o.merge()
        c, v = o.get_state(); f = o.get_forces()
        outs[kind] = dict(crd=c, vel=v, pairs=o.get_pairs(), energies=o.energies(), **f)
    for k in outs["port"]:
        assert np.array_equal(outs["port"][k], outs["reference"][k]), k


def test_reference_comment_known_answers(edmd, po):
    s = edmd.make_gc90()
    dt, gamma = s.params[0], s.params[1]
    assert round(1.0 / (1.0 + gamma * dt), 4) == 0.9960                 # edmd.cpp:298
    m3 = s.psets[0].masses3
    noise = np.sqrt(2.0 * gamma * s.params[4] * m3 / dt)                # edmd.cpp:184
    assert abs(noise.sum() - 3594.75) / 3594.75 < 1e-3                  # edmd.cpp:182


def test_glibc_rand_restatement_matches_libc(po):
    libc = ctypes.CDLL("libc.so.6")
    libc.rand.restype = ctypes.c_int
    for seed in (1, 13581, 1000013580, 4000000000, 0):
        libc.srand(ctypes.c_uint(seed))
        want = np.array([libc.rand() for _ in range(500)], dtype=np.int32)
        assert np.array_equal(po.glibc_rand_stream(seed, 500), want), seed


def test_qcp_invariants(edmd, po):
    s = edmd.make_gc90()
    o = po.Oracle(s, "port")
    rng = np.random.default_rng(1)
    ref = s.psets[0].avg.reshape(-1, 3)
    rot, rms = o.qcp(ref, ref)
    assert np.allclose(rot, np.eye(3), atol=1e-12) and rms < 1e-6
    for _ in range(5):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        moved = ref @ q.T + rng.normal(size=3) * 10
        rot, rms = o.qcp(ref, moved)
        assert np.allclose(rot @ rot.T, np.eye(3), atol=1e-12)
        assert abs(np.linalg.det(rot) - 1.0) < 1e-12
        assert np.allclose(rot, q.T, atol=1e-9)        # rot rotates `moved` back onto `ref`
        assert rms < 1e-5


def test_nb_pair_newton_third_law_and_double_counted_energy(edmd, po):
    s = edmd.make_ring(40, kind="coil", laps=2, lap_gap=9.0)
    o = po.Oracle(s, "port"); o.build_pairs()
    off = s.off3
    hits = 0
    for t1, t2 in o.get_pairs()[:200]:
        o.nb_forces_pair(int(t1), int(t2))
        f = o.get_forces()
        f1 = f["nb"][off[t1]:off[t1 + 1]].reshape(-1, 3); f2 = f["nb"][off[t2]:off[t2 + 1]].reshape(-1, 3)
        assert np.allclose(f1.sum(0), -f2.sum(0), rtol=1e-9, atol=1e-9)
        assert np.array_equal(f["nb_energy"][t1], f["nb_energy"][t2])   # edmd.cpp:282-283
        hits += int(np.any(f1 != 0))
    assert hits > 0


def test_merge_averages_and_is_idempotent(edmd, po):
    s = edmd.make_gc90()
    rng = np.random.default_rng(2)
    o = po.Oracle(s, "port")
    o.set_state(s.crd + rng.normal(scale=0.1, size=s.crd.size), rng.normal(size=s.crd.size))
    o.merge()
    c1, v1 = o.get_state()
    x = c1.reshape(s.n, 4, 63, 3)
    for t in range(s.n):                      # overlapping copies agree after the merge
        u = (t + 1) % s.n
        assert np.array_equal(x[t, 1:], x[u, :3]) or t == s.n - 1
    assert np.array_equal(x[s.n - 1, 1:], x[0, :3])     # ring closure, master.cpp:365-369
    o.merge()
    c2, v2 = o.get_state()
    assert np.allclose(c2, c1, rtol=0, atol=1e-13) and np.allclose(v2, v1, rtol=0, atol=1e-13)


def test_pair_list_predicate_gc90(edmd, po):
    s = edmd.make_gc90()
    o = po.Oracle(s, "port")
    assert o.build_pairs() == 450                       # SURVEY.md §8a a-PL
    p = o.get_pairs()
    sep = np.abs(p[:, 0] - p[:, 1]); sep = np.minimum(sep, s.n - sep)
    assert sep.min() == 6 and sep.max() == 10
    odd = (np.abs(p[:, 0] - p[:, 1]) % 2) == 1
    assert np.all(p[odd, 0] > p[odd, 1]) and np.all(p[~odd, 0] < p[~odd, 1])   # master.cpp:197-201
    assert np.all(np.bincount(p.reshape(-1), minlength=s.n) == 10)


def test_multithreaded_baseline_matches_serial(edmd, po):
    s = edmd.make_ring(40, kind="coil", laps=2, lap_gap=9.0)
    a = po.Oracle(s, "port"); a.build_pairs(); a.step(2, nthreads=1)
    b = po.Oracle(s, "port"); b.build_pairs(); b.step(2, nthreads=4)
    assert np.allclose(a.get_state()[0], b.get_state()[0], rtol=1e-12, atol=1e-12)


def test_division_free_rand_max_quotient_is_exact(tmp_path):
    """rng.cu replaces rand()/RAND_MAX's division by an FMA-corrected product; it must give the
    correctly rounded quotient for every possible rand() value (all 2^31 of them, ~2 s threaded)."""
    import os, subprocess
    if " fma " not in open("/proc/cpuinfo").read():
        pytest.skip("host CPU has no FMA unit: the exhaustive loop would run on libm's software fma")
    src = os.path.join(os.path.dirname(__file__), "cpp", "div_rand_max.c")
    exe = str(tmp_path / "div_rand_max")
    # -ffp-contract=off: the products and sums in the checker must round like the CUDA intrinsics
    subprocess.run(["/usr/bin/gcc", "-O2", "-fopenmp", "-mfma", "-ffp-contract=off", "-o", exe, src, "-lm"], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    assert int(out[0]) == 0          # corrected form: no mismatch
    assert int(out[1]) > 0           # (the plain product alone is NOT exact - the correction matters)


def test_glibc_seed_closed_form(po):
    """rng.cu seeds every lane with 16807^i * (int32)seed mod (2^31-1) instead of iterating srand()'s Schrage
    recurrence: the kernel's table and the closed form must reproduce the words stdlib/random_r.c's seeding loop
    (restated here with C's truncating division and int32 range checks) produces, for edge and random seeds."""
    M = 2147483647
    table = [pow(16807, i, M) for i in range(31)]
    src = open(os.path.join(ROOT, "msc-project_b200", "csrc", "rng.cu")).read()
    body = src[src.index("kSeedPow[31] = {") + len("kSeedPow[31] = {"):]
    body = body[:body.index("}")]
    assert [int(x.strip().rstrip("u")) for x in body.split(",")] == table

    def cdiv(a, b):
        q = abs(a) // abs(b)
        return q if (a >= 0) == (b >= 0) else -q

    def iterate(seed):
        seed = seed or 1
        w = seed if seed < 2 ** 31 else seed - 2 ** 32
        out = [seed]
        for _ in range(30):
            hi = cdiv(w, 127773); lo = w - hi * 127773
            w = 16807 * lo - 2836 * hi
            assert -2 ** 31 <= w < 2 ** 31
            if w < 0:
                w += M
            out.append(w)
        return out

    def closed(seed):
        seed = seed or 1
        s = seed if seed < 2 ** 31 else seed - 2 ** 32
        out = [seed]
        for i in range(1, 31):
            p = table[i] * s
            rem = abs(p) % M
            rem = -rem if p < 0 else rem
            out.append(rem + M if rem < 0 else rem)
        return out

    rng = np.random.default_rng(3)
    seeds = [0, 1, 2, M - 1, M, M + 1, 2 ** 31, 2 ** 31 + 5, 2 ** 32 - 2, 2 ** 32 - 1, 1000013580] + [int(x) for x in rng.integers(0, 2 ** 32, 20000)]
    for sd in seeds:
        assert iterate(sd) == closed(sd), sd
