"""World-size-2 test of the multi-GPU step protocol on CPU (gloo): the sharding plan, the two
all-gathers and the owner-computes accumulation of msc-project_b200/dist.py, with the compute
done by the CPU oracle standing in for the engine.  Result must equal the single-rank oracle
bit for bit (no force reduction exists in the protocol, so there is no reordering)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    """Implements the backend surface of dist.ShardedEngine on top of the CPU oracle."""

    def __init__(self, po, system):
        self.po, self.sys = po, system
        self.o = po.Oracle(system, "port")
        self.n = system.n
        self.off = system.off3
        self.S = 252
        self.crd = torch.zeros(self.n * 3 * self.S, dtype=torch.float64)
        self.vel = torch.zeros(self.n * 3 * self.S, dtype=torch.float64)
        self.hit = torch.zeros(0, dtype=torch.uint8)
        self.t0, self.t1, self.p0, self.p1 = 0, self.n, 0, 0
        self.fnb = np.zeros(system.crd.size)
        self._push()

    # the dist layer sees one flat tensor with tetrad-major blocks; layout inside a block is free
    def _push(self):
        c, v = self.o.get_state()
        self.crd.view(self.n, -1)[:, :756] = torch.from_numpy(c.reshape(self.n, 756))
        self.vel.view(self.n, -1)[:, :756] = torch.from_numpy(v.reshape(self.n, 756))

    def _pull(self):
        c = self.crd.view(self.n, -1)[:, :756].numpy().reshape(-1).copy()
        v = self.vel.view(self.n, -1)[:, :756].numpy().reshape(-1).copy()
        self.o.set_state(c, v)

    plane_elems = 3 * 252

    def crd_tensor(self): return self.crd
    def vel_tensor(self): return self.vel
    def hit_tensor(self): return self.hit

    def stream_ctx(self):
        import contextlib
        return contextlib.nullcontext()

    def set_shard(self, t0, t1, p0, p1): self.t0, self.t1, self.p0, self.p1 = t0, t1, p0, p1

    def build_pairlist(self):
        self._pull()
        npairs = self.o.build_pairs()
        self.pairs = self.o.get_pairs()
        self.hit = torch.zeros(npairs, dtype=torch.uint8)
        return npairs

    def get_pairlist(self):
        return self.pairs

    def ed_forces(self):
        self._pull()
        for t in range(self.t0, self.t1):
            self.o.ed_forces(t)
        self._push()

    def random_terms(self, mode, time0, rank, calls):
        assert mode == 0

    def nb_scan(self):
        self._pull()
        for p in range(self.p0, self.p1):
            t1, t2 = map(int, self.pairs[p])
            self.o.nb_forces_pair(t1, t2)
            f = self.o.get_forces()
            any_hit = np.any(f["nb"][self.off[t1]:self.off[t1 + 1]] != 0) or f["nb_energy"][t1].any()
            self.hit[p] = 1 if any_hit else 0

    def nb_accumulate(self):
        acc = np.zeros(self.sys.crd.size)
        for p, (t1, t2) in enumerate(self.pairs):
            if not self.hit[p]:
                continue
            own1, own2 = self.t0 <= t1 < self.t1, self.t0 <= t2 < self.t1
            if not (own1 or own2):
                continue
            self.o.nb_forces_pair(int(t1), int(t2))
            f = self.o.get_forces()["nb"]
            for t, own in ((t1, own1), (t2, own2)):
                if own:
                    acc[self.off[t]:self.off[t + 1]] += f[self.off[t]:self.off[t + 1]]
        self.fnb = np.clip(acc, -1.0, 1.0)

    def integrate(self):
        f = self.o.get_forces()
        self.o.set_forces(None, np.zeros_like(self.fnb), self.fnb)
        for t in range(self.t0, self.t1):
            self.o.update_velocities(t); self.o.update_coordinates(t)
        self._push()

    def merge(self):
        self._pull(); self.o.merge(); self._push()

    def sync(self): pass


def _skewed_weights(pairs):
    """Pairs that touch the first ten tetrads count fifty-fold (a stand-in for pairs in contact)."""
    return 1.0 + 49.0 * ((pairs[:, 0] < 10) | (pairs[:, 1] < 10))


def _worker(rank, world, port, q, replicate, halo=False, n_tet=30, weighted=False):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    pkg = importlib.import_module("msc-project_b200")
    dmod = importlib.import_module("msc-project_b200.dist")
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    system = pkg.make_ring(n_tet, kind="coil", laps=2, lap_gap=9.0)
    b = OracleBackend(po, system)
    sh = dmod.ShardedEngine(b, dist, system.n, rank, world, replicate=replicate, halo=halo,
                            pair_weights=_skewed_weights if weighted else None)
    npairs = sh.build_pairlist()
    sh.step(2, 0)
    sh.merge()
    sh.finish()
    b._pull()
    c, v = b.o.get_state()
    if rank == 0:
        halo_rows = int(sh.halo_plan.recv_idx.size) if sh.halo_plan is not None else -1
        q.put((npairs, c, v, (sh.plan.t0, sh.plan.t1, sh.plan.p0, sh.plan.p1), halo_rows))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("replicate", [False, True])
def test_sharded_protocol_world2_equals_single_rank(edmd, po, replicate):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (7 if replicate else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, replicate)) for r in range(2)]
    for p in procs:
        p.start()
    npairs, c, v, plan, _ = q.get(timeout=600)
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    system = edmd.make_ring(30, kind="coil", laps=2, lap_gap=9.0)
    o = po.Oracle(system, "port")
    assert o.build_pairs() == npairs
    o.step(2); o.merge()
    co, vo = o.get_state()
    assert np.array_equal(c, co) and np.array_equal(v, vo)
    assert plan == ((0, 30) if replicate else (0, 15)) + (0, (npairs + 1) // 2)


def test_shard_plan_covers_everything_once(edmd):
    dmod = importlib.import_module("msc-project_b200.dist")
    for n, npairs, world in ((1000, 494500, 8), (90, 450, 4), (1000, 494500, 3), (7, 5, 8), (10, 0, 2)):
        seen_t = np.zeros(n, int); seen_p = np.zeros(npairs, int)
        for r in range(world):
            pl = dmod.ShardPlan(n, npairs, world, r)
            seen_t[pl.t0:pl.t1] += 1; seen_p[pl.p0:pl.p1] += 1
            assert pl.t1 - pl.t0 <= pl.t_chunk and pl.p1 - pl.p0 <= pl.p_chunk
        assert np.all(seen_t == 1) and np.all(seen_p == 1)


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_equals_single_rank(edmd, po, world):
    """Sparse per-step exchange (variable-size all-to-all of the tetrad blocks a rank reads) instead
    of the coordinate all-gather; uneven blocks at world 3."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 20 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, False, True, 40)) for r in range(world)]
    for p in procs:
        p.start()
    npairs, c, v, plan, halo_rows = q.get(timeout=600)
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    system = edmd.make_ring(40, kind="coil", laps=2, lap_gap=9.0)
    o = po.Oracle(system, "port")
    assert o.build_pairs() == npairs
    o.step(2); o.merge()
    co, vo = o.get_state()
    assert np.array_equal(c, co) and np.array_equal(v, vo)
    assert 0 < halo_rows < system.n          # a real, partial halo


def test_weighted_pair_slices_equal_single_rank(edmd, po):
    """Pair slices of equal WEIGHT instead of equal count (world 3, skewed weights): the slices differ from the
    default plan's, the hit flags travel as uneven blocks, and the result is still the single-rank result bit for bit."""
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 40
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, False, True, 40, True)) for r in range(world)]
    for p in procs:
        p.start()
    npairs, c, v, plan, halo_rows = q.get(timeout=600)
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    system = edmd.make_ring(40, kind="coil", laps=2, lap_gap=9.0)
    o = po.Oracle(system, "port")
    assert o.build_pairs() == npairs
    dmod = importlib.import_module("msc-project_b200.dist")
    pairs = o.get_pairs()
    want = dmod.ShardPlan(system.n, npairs, world, 0, pair_weights=_skewed_weights(pairs))
    assert plan == (want.t0, want.t1, want.p0, want.p1)
    assert (want.p0, want.p1) != dmod.ShardPlan(system.n, npairs, world, 0).pair_range(0)
    o.step(2); o.merge()
    co, vo = o.get_state()
    assert np.array_equal(c, co) and np.array_equal(v, vo)


def test_weighted_shard_plan_balances_weight():
    dmod = importlib.import_module("msc-project_b200.dist")
    rng = np.random.default_rng(5)
    for npairs, world in ((1000, 8), (37, 3), (5, 8), (400005, 8)):
        w = rng.uniform(0.0, 1.0, npairs) ** 8 * 64 + 1.0          # a few heavy pairs among many light ones
        seen = np.zeros(npairs, int)
        sums = []
        for r in range(world):
            pl = dmod.ShardPlan(100, npairs, world, r, pair_weights=w)
            assert (pl.p0, pl.p1) == pl.pair_range(r) and not pl.even_pairs
            seen[pl.p0:pl.p1] += 1
            sums.append(w[pl.p0:pl.p1].sum())
        assert np.all(seen == 1)
        assert max(sums) <= w.sum() / world + w.max() + 1e-9       # no slice exceeds its share by more than one pair
    # all-zero weights, no weights, one rank: the equal-count plan
    for kw in (dict(pair_weights=np.zeros(10)), dict()):
        pl = dmod.ShardPlan(100, 10, 2, 1, **kw)
        assert (pl.p0, pl.p1) == (5, 10) and pl.p_bounds is None
    with pytest.raises(ValueError):
        dmod.ShardPlan(100, 10, 2, 0, pair_weights=np.ones(9))
    with pytest.raises(ValueError):
        dmod.ShardPlan(100, 10, 2, 0, pair_weights=-np.ones(10))


def test_halo_plan_lists(edmd):
    dmod = importlib.import_module("msc-project_b200.dist")
    n, world = 100, 4
    pairs = np.array([(i, (i + d) % n) for i in range(n) for d in (6, 7, 8) if (i + d) < n], dtype=np.int32)
    for r in range(world):
        pl = dmod.ShardPlan(n, len(pairs), world, r)
        hp = dmod.HaloPlan(pairs, n, world, r, (pl.t0, pl.t1), (pl.p0, pl.p1))
        used = set(pairs[pl.p0:pl.p1].ravel().tolist())
        for a, b in pairs:
            if pl.t0 <= a < pl.t1: used.add(int(b))
            if pl.t0 <= b < pl.t1: used.add(int(a))
        want = sorted(t for t in used if not (pl.t0 <= t < pl.t1))
        assert hp.recv_idx.tolist() == want
        assert hp.recv_counts.sum() == len(want)


class _FakeDist:
    """Single-process stand-in for torch.distributed (world 2 seen from rank 0): records the collectives."""

    def __init__(self, log, world):
        self.log, self.world = log, world

    def all_gather_object(self, out, obj, group=None):
        self.log.append("all_gather_object")
        for i in range(len(out)):
            out[i] = obj

    def barrier(self, group=None):
        self.log.append("barrier")


class _FakeNativeBackend:
    """Records what ShardedEngine asks of a backend that runs the sharded step itself (the GPU engine's peer API)."""

    def __init__(self, log, npairs_seq, outgrow_at=None):
        self.log, self.npairs_seq, self.builds, self.outgrow_at = log, list(npairs_seq), 0, outgrow_at
        self.attached = False

    def peer_export(self):
        self.log.append("export"); return b"x" * 320

    def peer_attach(self, blobs, world, rank):
        assert len(blobs) == 320 * world
        self.log.append("attach"); self.attached = True

    def peer_detach(self):
        self.log.append("detach"); self.attached = False

    def peer_gather(self, what):
        self.log.append(f"gather{what}")

    def build_pairlist(self):
        self.builds += 1
        if self.outgrow_at == self.builds and self.attached:
            self.log.append("build!outgrew")
            raise RuntimeError("edmd_b200 error -3: pair list of 99 pairs outgrew the peer-visible mask buffer (64)")
        self.log.append("build")
        return self.npairs_seq[min(self.builds, len(self.npairs_seq)) - 1]

    def set_shard(self, t0, t1, p0, p1): self.log.append(("shard", t0, t1, p0, p1))
    def engine_step(self, nsteps, mode): self.log.append(("step", nsteps, mode))
    def engine_set_rng(self, time0, rank, calls): self.log.append("set_rng")
    def merge(self): self.log.append("merge")
    def sync(self): self.log.append("sync")


def test_native_protocol_sequence():
    """The one-process-per-GPU glue around the in-engine multi-GPU step (dist.py, native path): handles are
    exchanged once, the shard follows the pair count, a rebuild right after a merge does not gather twice, and a
    pair list that outgrew the mapped mask buffer is answered by detach -> rebuild -> export -> attach."""
    dmod = importlib.import_module("msc-project_b200.dist")
    log = []
    b = _FakeNativeBackend(log, [101, 103, 99], outgrow_at=3)
    sh = dmod.ShardedEngine(b, _FakeDist(log, 2), 40, 0, 2, native=True)
    assert sh.native and not sh.replicate
    assert sh.build_pairlist() == 101
    assert log == ["build", "export", "all_gather_object", "attach", "barrier", ("shard", 0, 20, 0, 51)]
    del log[:]
    sh.step(5, 1)
    assert log == [("step", 5, 1)] and sh.calls == 5 * 40
    del log[:]
    sh.merge()
    assert sh.build_pairlist() == 103                     # coordinates are complete after the merge: no second gather
    assert log == ["gather3", "merge", "build", ("shard", 0, 20, 0, 52)]
    del log[:]
    sh.step(1, 0)
    assert sh.build_pairlist() == 99                      # stepped since: gather first; then the outgrow path
    assert log == [("step", 1, 0), "gather1", "build!outgrew", "barrier", "detach", "barrier", "build", "export",
                   "all_gather_object", "attach", "barrier", ("shard", 0, 20, 0, 50)]
    del log[:]
    sh.finish(); sh.close()
    assert log == ["gather7", "sync", "barrier", "detach", "barrier"]
