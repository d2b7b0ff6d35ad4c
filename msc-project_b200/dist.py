"""Multi-GPU driver: one process per GPU, torch.distributed for the plumbing.

With the GPU engine the per-step exchanges do NOT go through this file: ``ShardedEngine`` exchanges the
engines' CUDA IPC handles once (``edmd_peer_export`` / ``edmd_peer_attach``), sets the shard, and from then on
``step`` is one ``edmd_step`` call per rank — barriers, halo pull and mask broadcast are kernels inside the
engine's CUDA graph (csrc/peer.cu, csrc/engine.cu).  The torch-collective protocol below is the same
decomposition spelled out with all-gathers; it is what runs on CPU (gloo, world_size 2) with the oracle
standing in for the engine, and what any backend without peer memory would use.

Replaces the reference's MPI master/worker decomposition
(/root/reference/src/master.cpp:214-292 generate_Indexes / send_Workload_Indexes /
calculate_Forces, src/worker.cpp:109-187, src/mpilib.cpp): there is no master.  Every rank
holds the whole system on its GPU (as every MPI worker did, worker.cpp:67-87) but computes
only its shard:

  * tetrads are split into contiguous equal blocks (owner of a tetrad runs its ED force, random
    term, NB accumulation + clip, velocity and coordinate update);
  * the flat tetrad-pair list is split into contiguous slices for the FP64 distance pass (the reference's
    NB_Index split, master.cpp:225-239): equal counts, or equal total weight when the caller passes per-pair
    work estimates (``pair_weights``).

Small systems (n <= REPLICATE_MAX_TETRADS) skip the first exchange: the per-tetrad kernels cost
~50 ns per tetrad while an all-gather costs ~50 us of latency, so every rank runs ED / accumulate /
integrate for ALL tetrads (identical, deterministic results on every rank) and only the pair
scan is sharded — one all-gather of hit bytes per step.

Per step the path has two real exchange points, both all-gathers over NCCL/NVLink issued on
the engine's own CUDA stream:
    ED (owned) -> all-gather post-shake coordinates (replaces MPI_Bcast(MPI_Crds), master.cpp:279)
    NB scan (own pair slice) -> all-gather per-pair hit flags (replaces the dense
        MPI_Reduce of N x 758 doubles, master.cpp:288: one byte per pair instead)
    NB accumulate + clip + integrate (owned), deterministic: each tetrad is summed by one CTA.
No force reduction is needed, so results are bit-identical for every world size.

Large systems (the default above REPLICATE_MAX_TETRADS) do not all-gather coordinates either:
``HaloPlan`` lists, from the pair list, exactly the non-owned tetrads a rank reads (both ends of the
pairs in its scan slice, and the partners of its owned tetrads for the accumulate pass) and the
per-step exchange becomes one variable-size all-to-all of those 6 KB tetrad blocks — a few
hundred KB for chain-like DNA instead of 6 KB x N (605 MB at 1e5 tetrads).

``ShardPlan`` is pure host logic (tested on CPU with gloo, world_size 2).  ``ShardedEngine``
works with any backend object exposing the small method set of ``EngineBackend`` below.
"""
from __future__ import annotations

import numpy as np


def block_range(total: int, world: int, rank: int):
    """Equal contiguous blocks of ceil(total/world); the tail block may be short or empty."""
    chunk = -(-total // world) if total > 0 else 0
    lo = min(total, rank * chunk)
    hi = min(total, lo + chunk)
    return lo, hi, chunk


class ShardPlan:
    """Contiguous tetrad blocks and contiguous slices of the pair list (replaces generate_Indexes, master.cpp:214-242).
    Pair slices hold equal COUNTS by default; with ``pair_weights`` (one non-negative work estimate per listed
    pair, the same array on every rank) they hold equal total WEIGHT instead — the distance pass of a culled pair
    costs a sphere test, that of a pair in contact up to 64 atom blocks.  Any slicing gives the same result bits:
    a pair's block mask does not depend on which rank computes it."""

    def __init__(self, n_tetrads: int, n_pairs: int, world: int, rank: int, pair_weights=None):
        self.n, self.np, self.world, self.rank = n_tetrads, n_pairs, world, rank
        self.t0, self.t1, self.t_chunk = block_range(n_tetrads, world, rank)
        self.p0, self.p1, self.p_chunk = block_range(n_pairs, world, rank)
        self.p_bounds = None
        if pair_weights is not None and n_pairs > 0 and world > 1:
            w = np.asarray(pair_weights, dtype=np.float64).reshape(-1)
            if w.shape[0] != n_pairs or not np.all(np.isfinite(w)) or np.any(w < 0):
                raise ValueError("pair_weights: one finite non-negative value per listed pair")
            c = np.cumsum(w)
            if c[-1] > 0:
                # slice r ends with the first pair at which the running weight reaches (r + 1) / world of the total
                cuts = np.searchsorted(c, c[-1] * np.arange(1, world) / world, side="left") + 1
                b = np.concatenate([[0], np.minimum(cuts, n_pairs), [n_pairs]]).astype(np.int64)
                self.p_bounds = np.maximum.accumulate(b)
                self.p0, self.p1 = int(self.p_bounds[rank]), int(self.p_bounds[rank + 1])

    def tetrad_range(self, r):
        return block_range(self.n, self.world, r)[:2]

    def pair_range(self, r):
        if self.p_bounds is not None:
            return int(self.p_bounds[r]), int(self.p_bounds[r + 1])
        return block_range(self.np, self.world, r)[:2]

    @property
    def even_tetrads(self):
        return self.n == self.t_chunk * self.world

    @property
    def even_pairs(self):
        return self.p_bounds is None and self.np == self.p_chunk * self.world


class HaloPlan:
    """Which tetrad blocks this rank must receive / send every step (host logic, numpy)."""

    def __init__(self, pairs: np.ndarray, n: int, world: int, rank: int, t_range, p_range):
        t0, t1 = t_range
        p0, p1 = p_range
        chunk = block_range(n, world, 0)[2]
        a, b = pairs[:, 0].astype(np.int64), pairs[:, 1].astype(np.int64)
        own_a = (a >= t0) & (a < t1)
        own_b = (b >= t0) & (b < t1)
        need = np.unique(np.concatenate([a[p0:p1], b[p0:p1], b[own_a], a[own_b]])) if pairs.size else np.zeros(0, np.int64)
        need = need[(need < t0) | (need >= t1)]
        self.recv_idx = need                                   # sorted -> grouped by owner block
        owner = need // chunk if chunk else need
        self.recv_counts = np.bincount(owner, minlength=world).astype(np.int64)[:world]
        self.send_idx = None                                   # filled by exchange_lists
        self.send_counts = None

    def exchange_lists(self, dist, torch, device, group=None):
        world = len(self.recv_counts)
        rc = torch.tensor(self.recv_counts, dtype=torch.int64, device=device)
        sc = torch.zeros(world, dtype=torch.int64, device=device)
        dist.all_to_all_single(sc, rc, group=group)
        self.send_counts = [int(x) for x in sc.cpu()]
        want = torch.tensor(self.recv_idx, dtype=torch.int64, device=device)
        give = torch.zeros(sum(self.send_counts), dtype=torch.int64, device=device)
        dist.all_to_all_single(give, want, output_split_sizes=self.send_counts,
                               input_split_sizes=[int(x) for x in self.recv_counts], group=group)
        self.send_idx_t = give
        self.recv_idx_t = want
        self.recv_counts_l = [int(x) for x in self.recv_counts]
        return self


class _CudaView:
    """__cuda_array_interface__ shim so torch can wrap an engine-owned device buffer."""

    def __init__(self, ptr, nbytes, typestr, itemsize):
        self.__cuda_array_interface__ = {"shape": (nbytes // itemsize,), "typestr": typestr,
                                         "data": (ptr, False), "version": 2}


class EngineBackend:
    """Adapter: msc-project_b200 Engine + torch CUDA tensors over its buffers."""

    def __init__(self, engine, device):
        import torch
        self.torch = torch
        self.e = engine
        self.device = device
        self.stream = torch.cuda.ExternalStream(engine.stream, device=device)
        self.S = engine.atom_stride
        self._crd = None
        self._vel = None
        self._hit = None

    def _wrap(self, which, typestr, itemsize, dtype):
        ptr, nbytes = self.e.device_buffer(which)
        return self.torch.as_tensor(_CudaView(ptr, nbytes, typestr, itemsize), device=self.device).view(dtype)

    def crd_tensor(self):
        # edmd_device_buffer also makes the primary coordinate buffer current (the fused finish
        # kernel leaves new coordinates in a second buffer), so call it every time
        ptr, nbytes = self.e.device_buffer(self.e.BUF_CRD)
        if self._crd is None or self._crd[0] != ptr:
            self._crd = (ptr, self._wrap(self.e.BUF_CRD, "<f8", 8, self.torch.float64))
        return self._crd[1]

    def vel_tensor(self):
        if self._vel is None:
            self._vel = self._wrap(self.e.BUF_VEL, "<f8", 8, self.torch.float64)
        return self._vel

    def hit_tensor(self):
        ptr, nbytes = self.e.device_buffer(self.e.BUF_HIT)
        if self._hit is None or self._hit[0] != (ptr, nbytes):
            # one 64-bit block mask per pair (0 = not flagged), see nb.cu
            self._hit = ((ptr, nbytes), self._wrap(self.e.BUF_HIT, "<i8", 8, self.torch.int64))
        return self._hit[1]

    def invalidate(self):
        self._hit = None

    @property
    def plane_elems(self):   # doubles per tetrad in a state buffer
        return 3 * self.S

    def stream_ctx(self):
        return self.torch.cuda.stream(self.stream)

    # compute
    def set_shard(self, t0, t1, p0, p1): self.e.set_shard(t0, t1, p0, p1)
    def ed_forces(self): self.e.ed_forces()
    def random_terms(self, mode, time0, rank, calls): self.e.random_terms(mode, time0, rank, calls)
    def nb_scan(self): self.e.nb_scan()
    def nb_accumulate(self): self.e.nb_accumulate()
    def integrate(self): self.e.integrate()
    def nb_finish(self): self.e.nb_finish()
    def merge(self): self.e.merge()
    def build_pairlist(self): return self.e.build_pairlist()
    def get_pairlist(self): return self.e.get_pairlist()
    def sync(self): self.e.sync()

    # in-engine multi-GPU (peer memory)
    def peer_export(self): return self.e.peer_export()
    def peer_attach(self, blobs, world, rank): self.e.peer_attach(blobs, world, rank)
    def peer_detach(self): self.e.peer_detach()
    def peer_gather(self, what): self.e.peer_gather(what)
    def engine_step(self, nsteps, random_mode): self.e.step(nsteps, random_mode)
    def engine_set_rng(self, time0, rank, calls): self.e.set_rng(time0, rank, calls)


# Below this size the per-tetrad kernels are replicated on every rank instead of exchanging
# coordinates (cost model in the module docstring: ~47 ns of ED+finish per tetrad vs ~50 us + 18 ns per
# tetrad for the all-gather -> break-even near 2,000 tetrads).
REPLICATE_MAX_TETRADS = 2048


class ShardedEngine:
    """Runs the reference step protocol (simulation.cpp:34-55) over world_size ranks."""

    def __init__(self, backend, dist, n_tetrads: int, rank: int, world: int, group=None, replicate=None,
                 halo=None, native=None, pair_weights=None):
        self.b, self.dist, self.n, self.rank, self.world, self.group = backend, dist, n_tetrads, rank, world, group
        # per-pair work estimates for ShardPlan: None (equal counts), an array, or a function of the pair list
        # ([np, 2] int array) evaluated after every rebuild — it must return the same values on every rank
        self.pair_weights = pair_weights
        # native: the engine itself runs the sharded step over peer memory (GPU engine, world > 1)
        self.native = (hasattr(backend, "peer_export") and world > 1) if native is None else bool(native)
        if replicate is None:
            replicate = False if self.native else n_tetrads <= REPLICATE_MAX_TETRADS
        self.replicate = bool(replicate)
        # halo: sparse exchange of the tetrad blocks actually read (default for large systems)
        self.halo = (not self.replicate) if halo is None else (bool(halo) and not self.replicate)
        self.halo_plan = None
        self._attached = False
        self._complete = False                   # every rank's copy of the coordinates is complete (after merge / finish)
        self.plan = None
        self.time0, self.rng_rank, self.calls = 1000000000, 1, 0

    # -- collectives -------------------------------------------------------------------
    def _all_gather_blocks(self, flat, elems_per_unit, total_units, chunk_units, ranges=None):
        """In-place all-gather of contiguous per-rank blocks of `flat` (a 1-D tensor): equal blocks of
        `chunk_units`, or the explicit per-rank `ranges` [(lo, hi)] of a weighted plan."""
        if self.world == 1:
            return
        even = ranges is None and total_units == chunk_units * self.world
        lo, hi, _ = block_range(total_units, self.world, self.rank)
        if even:
            mine = flat[lo * elems_per_unit: hi * elems_per_unit]
            self.dist.all_gather_into_tensor(flat[: total_units * elems_per_unit], mine, group=self.group)
        else:
            parts = []
            for r in range(self.world):
                a, b = ranges[r] if ranges is not None else block_range(total_units, self.world, r)[:2]
                parts.append(flat[a * elems_per_unit: b * elems_per_unit])
            # uneven blocks: one broadcast per non-empty block
            for r, part in enumerate(parts):
                if part.numel():
                    self.dist.broadcast(part, src=r if self.group is None else self.dist.get_global_rank(self.group, r),
                                        group=self.group)

    def gather_coordinates(self, full=False):
        if self.replicate or self.world == 1:
            return          # every rank integrates every tetrad: nothing to exchange
        if self.halo and not full and self.halo_plan is not None:
            hp = self.halo_plan
            crd = self.b.crd_tensor().view(self.n, self.b.plane_elems)
            send = crd.index_select(0, hp.send_idx_t)
            recv = crd.new_empty((int(hp.recv_idx_t.numel()), self.b.plane_elems))
            self.dist.all_to_all_single(recv, send, output_split_sizes=hp.recv_counts_l,
                                        input_split_sizes=hp.send_counts, group=self.group)
            crd.index_copy_(0, hp.recv_idx_t, recv)
            return
        self._all_gather_blocks(self.b.crd_tensor(), self.b.plane_elems, self.n, self.plan_t_chunk)

    def gather_velocities(self):
        if self.replicate:
            return
        self._all_gather_blocks(self.b.vel_tensor(), self.b.plane_elems, self.n, self.plan_t_chunk)

    def gather_hits(self):
        ranges = None
        if self.plan.p_bounds is not None:
            ranges = [self.plan.pair_range(r) for r in range(self.world)]
        self._all_gather_blocks(self.b.hit_tensor(), 1, self.plan.np, self.plan.p_chunk, ranges)

    def _make_plan(self, npairs):
        w = self.pair_weights
        if callable(w):
            w = w(self.b.get_pairlist()) if npairs > 0 else None
        return ShardPlan(self.n, npairs, self.world, self.rank, pair_weights=w)

    @property
    def plan_t_chunk(self):
        return block_range(self.n, self.world, 0)[2]

    # -- protocol ----------------------------------------------------------------------
    def build_pairlist(self) -> int:
        """Every rank builds the same list from the same (gathered) coordinates — the list is
        a pure function of the centroids, so no exchange is needed (master.cpp:180-210)."""
        if self.native:
            return self._build_native()
        self.b.set_shard(0, self.n, 0, 0)
        npairs = self.b.build_pairlist()
        self.plan = self._make_plan(npairs)
        if self.replicate:
            self.plan.t0, self.plan.t1 = 0, self.n
        self.b.set_shard(self.plan.t0, self.plan.t1, self.plan.p0, self.plan.p1)
        if hasattr(self.b, "invalidate"):
            self.b.invalidate()
        if self.halo and self.world > 1:
            pairs = self.b.get_pairlist()
            hp = HaloPlan(pairs, self.n, self.world, self.rank, (self.plan.t0, self.plan.t1), (self.plan.p0, self.plan.p1))
            dev = self.b.crd_tensor().device
            import torch
            with self.b.stream_ctx():
                self.halo_plan = hp.exchange_lists(self.dist, torch, dev, self.group)
        return npairs

    # -- in-engine path ------------------------------------------------------------------
    def _attach(self):
        blob = self.b.peer_export()
        every = [None] * self.world
        self.dist.all_gather_object(every, blob, group=self.group)
        self.b.peer_attach(b"".join(every), self.world, self.rank)
        self.dist.barrier(group=self.group)      # every rank has mapped every buffer
        self._attached = True

    def _detach(self):
        if self._attached:
            self.dist.barrier(group=self.group)  # nobody is still stepping into a buffer about to be unmapped
            self.b.peer_detach()
            self.dist.barrier(group=self.group)
            self._attached = False

    def _build_native(self) -> int:
        if self._attached and not self._complete:
            self.b.peer_gather(1)                # owners' coordinates -> every rank
        try:
            npairs = self.b.build_pairlist()
        except RuntimeError as err:              # the list outgrew the peer-visible mask buffer (same on every rank)
            if not (self._attached and "outgrew" in str(err)):
                raise
            self._detach()
            npairs = self.b.build_pairlist()
        if not self._attached:
            self._attach()
        self.plan = self._make_plan(npairs)
        if self.replicate:
            self.plan.t0, self.plan.t1 = 0, self.n
        self.b.set_shard(self.plan.t0, self.plan.t1, self.plan.p0, self.plan.p1)
        return npairs

    def close(self):
        """Unmap the peers (collective); call before destroying the engines."""
        if self.native:
            self._detach()

    def set_rng(self, time0, rank, calls):
        self.time0, self.rng_rank, self.calls = time0, rank, calls
        if self.native:
            self.b.engine_set_rng(time0, rank, calls)

    def step(self, nsteps=1, random_mode=0):
        if self.native:                          # one call: the sharded step is inside the engine
            self._complete = False
            self.b.engine_step(nsteps, random_mode)
            if random_mode == 1:
                self.calls += nsteps * self.n
            return
        with self.b.stream_ctx():
            for _ in range(nsteps):
                self.b.ed_forces()
                self.b.random_terms(random_mode, self.time0, self.rng_rank, self.calls)
                if random_mode == 1:
                    self.calls += self.n
                self.gather_coordinates()
                self.b.nb_scan()
                self.gather_hits()
                if hasattr(self.b, "nb_finish"):
                    self.b.nb_finish()          # accumulate + clip + integrate in one kernel
                else:
                    self.b.nb_accumulate()
                    self.b.integrate()

    def merge(self):
        """4-fold overlap average needs neighbours' copies: gather state, merge everywhere."""
        if self.native:
            self.b.peer_gather(3)
            self.b.merge()
            self._complete = True
            return
        with self.b.stream_ctx():
            self.gather_coordinates(full=True)
            self.gather_velocities()
            self.b.merge()

    def finish(self):
        """Make every rank's copy of coordinates and velocities complete (for output)."""
        if self.native:
            self.b.peer_gather(7)
            self.b.sync()
            self._complete = True
            return
        with self.b.stream_ctx():
            self.gather_coordinates(full=True)
            self.gather_velocities()
        self.b.sync()
