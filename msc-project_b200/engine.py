"""ctypes binding of include/edmd_b200.h.

``Tetrad`` mirrors the reference's ``class Tetrad`` data members
(/root/reference/src/tetrad.hpp:36-60) byte for byte; ``build_tetrads`` lays an
:class:`EdmdSystem` out the way ``IO::read_Prm`` + ``initialise_Tetrad_Crds`` do (io.cpp:120-288),
so the engine is driven through exactly the structs reference code would hand it.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .system import EdmdSystem

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

PD = C.POINTER(C.c_double)


class EngineError(RuntimeError):
    pass


class Tetrad(C.Structure):
    _fields_ = [("num_Atoms", C.c_int), ("num_Evecs", C.c_int),
                ("avg", PD), ("masses", PD), ("abq", PD), ("eigenvalues", PD),
                ("eigenvectors", C.POINTER(PD)),
                ("ED_Forces", PD), ("random_Terms", PD), ("NB_Forces", PD),
                ("temperature", C.c_double), ("velocities", PD), ("coordinates", PD)]


assert C.sizeof(Tetrad) == 96

# numpy view of the same struct, for vectorised pointer filling
TETRAD_DTYPE = np.dtype([("num_Atoms", "<i4"), ("num_Evecs", "<i4"), ("avg", "<u8"), ("masses", "<u8"),
                         ("abq", "<u8"), ("eigenvalues", "<u8"), ("eigenvectors", "<u8"),
                         ("ED_Forces", "<u8"), ("random_Terms", "<u8"), ("NB_Forces", "<u8"),
                         ("temperature", "<f8"), ("velocities", "<u8"), ("coordinates", "<u8")])
assert TETRAD_DTYPE.itemsize == 96


class Params(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("dt", "gamma", "tautp", "temperature", "scaled",
                                          "mole_Cutoff", "atom_Cutoff", "mole_Least")]


def library_path() -> str:
    return os.path.join(_HERE, "libedmd_b200.so")


def load_library():
    """Load libedmd_b200.so (built by ``__graft_entry__.build()`` / ``make -C msc-project_b200/csrc``)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise EngineError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`."
                          " There is no CPU fallback.")
    lib = C.CDLL(path)
    E = C.c_void_p
    LL = C.c_longlong
    sig = {
        "edmd_last_error": (C.c_char_p, []),
        "edmd_abi_version": (C.c_int, []),
        "edmd_create": (C.c_int, [C.POINTER(Params), C.c_int, C.POINTER(E)]),
        "edmd_destroy": (None, [E]),
        "edmd_set_params": (C.c_int, [E, C.POINTER(Params)]),
        "edmd_set_nb_consts": (C.c_int, [E, C.c_double, C.c_double]),
        "edmd_set_nb_clip": (C.c_int, [E, C.c_int]),
        "edmd_set_nb_scan_mode": (C.c_int, [E, C.c_int]),
        "edmd_set_ed_kernel": (C.c_int, [E, C.c_int]),
        "edmd_step_host": (C.c_int, [E, C.c_void_p, C.c_int, C.c_int, C.c_int]),
        "edmd_upload_tetrads": (C.c_int, [E, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
        "edmd_set_state": (C.c_int, [E, C.c_void_p, C.c_int]),
        "edmd_get_state": (C.c_int, [E, C.c_void_p, C.c_int]),
        "edmd_set_state_owned": (C.c_int, [E, C.c_void_p, C.c_int]),
        "edmd_get_state_owned": (C.c_int, [E, C.c_void_p, C.c_int]),
        "edmd_get_forces": (C.c_int, [E, C.c_void_p, C.c_int]),
        "edmd_set_forces": (C.c_int, [E, C.c_void_p, C.c_int]),
        "edmd_set_state_flat": (C.c_int, [E, PD, PD]),
        "edmd_get_state_flat": (C.c_int, [E, PD, PD]),
        "edmd_get_forces_flat": (C.c_int, [E, PD, PD, PD, PD, PD, PD]),
        "edmd_set_pair_builder": (C.c_int, [E, C.c_int]),
        "edmd_build_pairlist": (C.c_int, [E, C.POINTER(LL)]),
        "edmd_get_pairlist": (C.c_int, [E, C.POINTER(C.c_int), LL]),
        "edmd_set_pairlist": (C.c_int, [E, C.POINTER(C.c_int), LL]),
        "edmd_set_shard": (C.c_int, [E, C.c_int, C.c_int, LL, LL]),
        "edmd_peer_export": (C.c_int, [E, C.c_char_p]),
        "edmd_peer_attach": (C.c_int, [E, C.c_char_p, C.c_int, C.c_int]),
        "edmd_peer_detach": (C.c_int, [E]),
        "edmd_peer_gather": (C.c_int, [E, C.c_int]),
        "edmd_world": (C.c_int, [E, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "edmd_group_create": (C.c_int, [C.POINTER(Params), C.c_int, C.POINTER(C.c_int), C.POINTER(E)]),
        "edmd_group_destroy": (None, [E]),
        "edmd_group_size": (C.c_int, [E]),
        "edmd_group_engine": (E, [E, C.c_int]),
        "edmd_group_upload_tetrads": (C.c_int, [E, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
        "edmd_group_set_state": (C.c_int, [E, C.c_void_p, C.c_int]),
        "edmd_group_get_state": (C.c_int, [E, C.c_void_p, C.c_int]),
        "edmd_group_set_rng": (C.c_int, [E, C.c_long, C.c_int, C.c_long]),
        "edmd_group_build_pairlist": (C.c_int, [E, C.POINTER(LL)]),
        "edmd_group_step": (C.c_int, [E, C.c_int, C.c_int]),
        "edmd_group_merge": (C.c_int, [E]),
        "edmd_group_get_energies": (C.c_int, [E, PD]),
        "edmd_group_sync": (C.c_int, [E]),
        "edmd_group_snapshot_begin": (C.c_int, [E, C.c_int, PD, PD, PD]),
        "edmd_group_snapshot_wait": (C.c_int, [E, C.c_int]),
        "edmd_snapshot_begin": (C.c_int, [E, C.c_int, PD, PD, PD]),
        "edmd_snapshot_wait": (C.c_int, [E, C.c_int]),
        "edmd_state_doubles": (LL, [E]),
        "edmd_ed_forces": (C.c_int, [E]),
        "edmd_random_terms": (C.c_int, [E, C.c_int, C.c_long, C.c_int, C.c_long]),
        "edmd_nb_forces": (C.c_int, [E]),
        "edmd_nb_scan": (C.c_int, [E]),
        "edmd_nb_accumulate": (C.c_int, [E]),
        "edmd_update_velocities": (C.c_int, [E]),
        "edmd_update_coordinates": (C.c_int, [E]),
        "edmd_nb_finish": (C.c_int, [E]),
        "edmd_integrate": (C.c_int, [E]),
        "edmd_merge": (C.c_int, [E]),
        "edmd_get_energies": (C.c_int, [E, PD]),
        "edmd_set_rng": (C.c_int, [E, C.c_long, C.c_int, C.c_long]),
        "edmd_step": (C.c_int, [E, C.c_int, C.c_int]),
        "edmd_set_graph": (C.c_int, [E, C.c_int]),
        "edmd_set_rng_branch": (C.c_int, [E, C.c_int]),
        "edmd_sync": (C.c_int, [E]),
        "edmd_device_buffer": (C.c_int, [E, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
        "edmd_atom_stride": (C.c_int, [E]),
        "edmd_num_tetrads": (C.c_int, [E]),
        "edmd_num_pairs": (LL, [E]),
        "edmd_stream": (C.c_void_p, [E]),
        "edmd_timer_start": (C.c_int, [E]),
        "edmd_timer_stop": (C.c_int, [E, PD]),
        "edmd_mark": (C.c_int, [E, C.c_int]),
        "edmd_mark_elapsed": (C.c_int, [E, C.c_int, C.c_int, PD]),
        "edmd_timing_enable": (C.c_int, [E, C.c_int]),
        "edmd_timing_get": (C.c_int, [E, C.c_int, PD, C.POINTER(LL)]),
        "edmd_timing_reset": (C.c_int, [E]),
        "edmd_launch_count": (LL, [E]),
        "edmd_nb_stats": (C.c_int, [E, C.POINTER(LL)]),
        "edmd_nb_stats_ex": (C.c_int, [E, C.POINTER(LL)]),
        "edmd_alloc_pinned": (C.c_void_p, [C.c_size_t]),
        "edmd_free_pinned": (None, [C.c_void_p]),
        "edmd_measure_fp64_peak": (C.c_int, [C.c_int, PD, PD]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib._edmd_symbols = tuple(sig)
    _LIB = lib
    return lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(PD)


class TetradBlock:
    """Host memory for n reference-layout Tetrad structs plus the arrays they point into."""

    def __init__(self, sys: EdmdSystem, pinned: bool = False):
        n = sys.n
        self.sys = sys
        na = sys.num_atoms.astype(np.int64)
        self.off3 = sys.off3
        n3 = 3 * na
        tot = int(self.off3[-1])
        self._pinned = []
        if pinned:      # coordinates / velocities in page-locked memory (edmd_alloc_pinned)
            lib = load_library()
            bufs = []
            for _ in range(2):
                p = lib.edmd_alloc_pinned(8 * tot)
                if not p:
                    raise EngineError(lib.edmd_last_error().decode())
                self._pinned.append((lib, p))
                bufs.append(np.ctypeslib.as_array(C.cast(p, PD), shape=(tot,)))
            self.crd, self.vel = bufs
            self.crd[:] = sys.crd; self.vel[:] = sys.vel
        else:
            self.crd = np.ascontiguousarray(sys.crd, dtype=np.float64).copy()
            self.vel = np.ascontiguousarray(sys.vel, dtype=np.float64).copy()
        self.off_ed = self.off3 + np.arange(n + 1)          # 3N+1 per tetrad
        self.off_nb = self.off3 + 2 * np.arange(n + 1)      # 3N+2 per tetrad
        self.ed = np.zeros(tot + n)
        self.rnd = np.zeros(tot)
        self.nb = np.zeros(tot + 2 * n)
        # row-pointer tables for eigenvectors (Array::allocate_2D_Double_Array, array.cpp:24-34)
        self._rows = []
        for ps in sys.psets:
            base = ps.evecs.ctypes.data
            self._rows.append(np.array([base + 8 * 3 * ps.num_atoms * j for j in range(max(ps.num_evecs, 1))],
                                       dtype=np.uint64))
        self.structs = np.zeros(n, dtype=TETRAD_DTYPE)
        pid = sys.param_id
        s = self.structs
        s["num_Atoms"] = na
        s["num_Evecs"] = sys.num_evecs
        for fld, attr in (("avg", "avg"), ("masses", "masses3"), ("abq", "abq"), ("eigenvalues", "evals")):
            table = np.array([getattr(ps, attr).ctypes.data for ps in sys.psets], dtype=np.uint64)
            s[fld] = table[pid]
        s["eigenvectors"] = np.array([r.ctypes.data for r in self._rows], dtype=np.uint64)[pid]
        s["coordinates"] = self.crd.ctypes.data + 8 * self.off3[:-1].astype(np.uint64)
        s["velocities"] = self.vel.ctypes.data + 8 * self.off3[:-1].astype(np.uint64)
        s["ED_Forces"] = self.ed.ctypes.data + 8 * self.off_ed[:-1].astype(np.uint64)
        s["random_Terms"] = self.rnd.ctypes.data + 8 * self.off3[:-1].astype(np.uint64)
        s["NB_Forces"] = self.nb.ctypes.data + 8 * self.off_nb[:-1].astype(np.uint64)
        for ps in sys.psets:
            for a in (ps.avg, ps.masses3, ps.abq, ps.evals, ps.evecs):
                assert a.dtype == np.float64 and a.flags.c_contiguous

    @property
    def ptr(self):
        return C.c_void_p(self.structs.ctypes.data)

    def __del__(self):
        for lib, p in getattr(self, "_pinned", []):
            try:
                lib.edmd_free_pinned(p)
            except Exception:
                pass
        self._pinned = []

    # views of the reference-layout force arrays
    def ed_forces(self):
        n = self.sys.n
        idx = np.ones(self.ed.shape[0], dtype=bool); idx[self.off_ed[1:] - 1] = False
        return self.ed[idx], self.ed[self.off_ed[1:] - 1]

    def nb_forces(self):
        idx = np.ones(self.nb.shape[0], dtype=bool)
        idx[self.off_nb[1:] - 1] = False; idx[self.off_nb[1:] - 2] = False
        e = np.stack([self.nb[self.off_nb[1:] - 2], self.nb[self.off_nb[1:] - 1]], axis=1)
        return self.nb[idx], e

    def temperature(self):
        return self.structs["temperature"].copy()


def build_tetrads(sys: EdmdSystem, pinned: bool = False) -> TetradBlock:
    return TetradBlock(sys, pinned)


class Engine:
    """One engine = one GPU.  Mirrors the reference call surface: the five EDMD methods
    (edmd.hpp:54-170) batched over tetrads, plus the Master bookkeeping on the hot path."""

    T_ED, T_SCAN, T_ACC, T_RNG, T_MERGE, T_PAIRS, T_INT, T_LAYOUT, T_SUP = range(9)
    BUF_CRD, BUF_VEL, BUF_FED, BUF_RND, BUF_FNB, BUF_HIT, BUF_CENTROID = range(7)

    def __init__(self, sys: EdmdSystem, device: int = 0, pinned: bool = True):
        self.lib = load_library()
        self.h = C.c_void_p()
        p = Params(*[float(x) for x in sys.params])
        self._ck(self.lib.edmd_create(C.byref(p), device, C.byref(self.h)))
        self.sys = sys
        self.block = build_tetrads(sys, pinned)
        bp = None
        if sys.bp_atoms is not None:
            self._bp = np.ascontiguousarray(sys.bp_atoms, dtype=np.int32)
            bp = self._bp.ctypes.data_as(C.POINTER(C.c_int))
        self._ck(self.lib.edmd_upload_tetrads(self.h, self.block.ptr, sys.n, bp))

    def _ck(self, rc):
        if rc != 0:
            raise EngineError(f"edmd_b200 error {rc}: {self.lib.edmd_last_error().decode()}")

    def close(self):
        if self.h:
            self.lib.edmd_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- data -------------------------------------------------------------------------
    def set_state(self, crd=None, vel=None):
        if crd is not None:
            self.block.crd[:] = crd
        if vel is not None:
            self.block.vel[:] = vel
        self._ck(self.lib.edmd_set_state(self.h, self.block.ptr, self.sys.n))

    def get_state(self, copy: bool = True):
        """Device -> the host Tetrad arrays (self.block.crd / .vel); returns copies unless copy=False."""
        self._ck(self.lib.edmd_get_state(self.h, self.block.ptr, self.sys.n))
        if not copy:
            return self.block.crd, self.block.vel
        return self.block.crd.copy(), self.block.vel.copy()

    def push_state(self):
        """Host Tetrad arrays (self.block.crd / .vel, edited in place) -> device."""
        self._ck(self.lib.edmd_set_state(self.h, self.block.ptr, self.sys.n))

    def push_state_owned(self):
        """Owned tetrads of the host Tetrad arrays -> device (asynchronous)."""
        self._ck(self.lib.edmd_set_state_owned(self.h, self.block.ptr, self.sys.n))

    def get_state_owned(self):
        """Device -> the owned tetrads' part of the host Tetrad arrays (blocks until it is there)."""
        self._ck(self.lib.edmd_get_state_owned(self.h, self.block.ptr, self.sys.n))
        return self.block.crd, self.block.vel

    def get_forces(self):
        """dict(ed, ed_energy, rnd, nb, nb_energy[n,2], temperature) via the Tetrad structs."""
        self._ck(self.lib.edmd_get_forces(self.h, self.block.ptr, self.sys.n))
        ed, eed = self.block.ed_forces()
        nb, enb = self.block.nb_forces()
        return dict(ed=ed.copy(), ed_energy=eed.copy(), rnd=self.block.rnd.copy(), nb=nb.copy(),
                    nb_energy=enb.copy(), temperature=self.block.temperature())

    def set_forces(self, ed=None, rnd=None, nb=None):
        b = self.block
        if ed is not None:
            idx = np.ones(b.ed.shape[0], dtype=bool); idx[b.off_ed[1:] - 1] = False
            b.ed[idx] = ed
        if rnd is not None:
            b.rnd[:] = rnd
        if nb is not None:
            idx = np.ones(b.nb.shape[0], dtype=bool)
            idx[b.off_nb[1:] - 1] = False; idx[b.off_nb[1:] - 2] = False
            b.nb[idx] = nb
        self._ck(self.lib.edmd_set_forces(self.h, b.ptr, self.sys.n))

    def set_state_flat(self, crd=None, vel=None):
        c = None if crd is None else _ptr(np.ascontiguousarray(crd, dtype=np.float64))
        v = None if vel is None else _ptr(np.ascontiguousarray(vel, dtype=np.float64))
        self._ck(self.lib.edmd_set_state_flat(self.h, c, v))

    def get_state_flat(self, crd: np.ndarray, vel: np.ndarray):
        self._ck(self.lib.edmd_get_state_flat(self.h, _ptr(crd), _ptr(vel)))

    # ---- pair list --------------------------------------------------------------------
    def set_pair_builder(self, which: int):
        self._ck(self.lib.edmd_set_pair_builder(self.h, which))

    def build_pairlist(self) -> int:
        npairs = C.c_longlong()
        self._ck(self.lib.edmd_build_pairlist(self.h, C.byref(npairs)))
        return npairs.value

    def get_pairlist(self) -> np.ndarray:
        npairs = self.lib.edmd_num_pairs(self.h)
        out = np.zeros((npairs, 2), dtype=np.int32)
        if npairs:
            self._ck(self.lib.edmd_get_pairlist(self.h, out.ctypes.data_as(C.POINTER(C.c_int)), npairs))
        return out

    def set_pairlist(self, pairs):
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        self._ck(self.lib.edmd_set_pairlist(self.h, pairs.ctypes.data_as(C.POINTER(C.c_int)), pairs.shape[0]))

    def set_shard(self, t0, t1, p0, p1):
        self._ck(self.lib.edmd_set_shard(self.h, t0, t1, p0, p1))

    @property
    def num_pairs(self) -> int:
        return self.lib.edmd_num_pairs(self.h)

    # ---- the hot path -----------------------------------------------------------------
    def set_nb_consts(self, krep=100.0, qfac=0.0):
        self._ck(self.lib.edmd_set_nb_consts(self.h, krep, qfac))

    def set_ed_kernel(self, which: int):
        self._ck(self.lib.edmd_set_ed_kernel(self.h, which))

    def set_nb_scan_mode(self, mode: int):
        self._ck(self.lib.edmd_set_nb_scan_mode(self.h, mode))

    PEER_BLOB_BYTES = 320
    GATHER_CRD, GATHER_VEL, GATHER_OBS = 1, 2, 4

    def peer_export(self) -> bytes:
        """IPC handles of this engine's peer-visible buffers (edmd_peer_export)."""
        buf = C.create_string_buffer(self.PEER_BLOB_BYTES)
        self._ck(self.lib.edmd_peer_export(self.h, buf))
        return buf.raw

    def peer_attach(self, all_blobs: bytes, world: int, rank: int):
        self._ck(self.lib.edmd_peer_attach(self.h, all_blobs, world, rank))

    def peer_detach(self):
        self._ck(self.lib.edmd_peer_detach(self.h))

    def peer_gather(self, what: int = 3):
        """Collective (every rank calls it): complete this rank's copy of coordinates (1), velocities (2),
        per-tetrad energies and temperatures (4)."""
        self._ck(self.lib.edmd_peer_gather(self.h, what))

    def ed_forces(self):
        self._ck(self.lib.edmd_ed_forces(self.h))

    def random_terms(self, mode=1, time0=1000000000, rank=1, calls_before=0):
        self._ck(self.lib.edmd_random_terms(self.h, mode, time0, rank, calls_before))

    def nb_forces(self):
        self._ck(self.lib.edmd_nb_forces(self.h))

    def nb_scan(self):
        self._ck(self.lib.edmd_nb_scan(self.h))

    def nb_accumulate(self):
        self._ck(self.lib.edmd_nb_accumulate(self.h))

    def update_velocities(self):
        self._ck(self.lib.edmd_update_velocities(self.h))

    def update_coordinates(self):
        self._ck(self.lib.edmd_update_coordinates(self.h))

    def nb_finish(self):
        self._ck(self.lib.edmd_nb_finish(self.h))

    def integrate(self):
        self._ck(self.lib.edmd_integrate(self.h))

    def merge(self):
        self._ck(self.lib.edmd_merge(self.h))

    def energies(self) -> np.ndarray:
        out = np.zeros(4)
        self._ck(self.lib.edmd_get_energies(self.h, _ptr(out)))
        return out

    def set_rng(self, time0=1000000000, rank=1, calls_before=0):
        self._ck(self.lib.edmd_set_rng(self.h, time0, rank, calls_before))

    def step(self, nsteps=1, random_mode=0):
        self._ck(self.lib.edmd_step(self.h, nsteps, random_mode))

    def step_host(self, nsteps=1, random_mode=0):
        """edmd_step_host: the host Tetrad block (self.block.crd / .vel) in, stepped, and back in place."""
        self._ck(self.lib.edmd_step_host(self.h, self.block.ptr, self.sys.n, nsteps, random_mode))
        return self.block.crd, self.block.vel

    def set_rng_branch(self, on: bool):
        self._ck(self.lib.edmd_set_rng_branch(self.h, int(on)))

    def set_graph(self, on: bool):
        self._ck(self.lib.edmd_set_graph(self.h, int(on)))

    def sync(self):
        self._ck(self.lib.edmd_sync(self.h))

    # ---- introspection ----------------------------------------------------------------
    def device_buffer(self, which):
        p = C.c_void_p(); b = C.c_size_t()
        self._ck(self.lib.edmd_device_buffer(self.h, which, C.byref(p), C.byref(b)))
        return p.value, b.value

    @property
    def stream(self):
        return self.lib.edmd_stream(self.h)

    @property
    def atom_stride(self):
        return self.lib.edmd_atom_stride(self.h)

    def timer_start(self):
        self._ck(self.lib.edmd_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_double()
        self._ck(self.lib.edmd_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def mark(self, slot: int):
        self._ck(self.lib.edmd_mark(self.h, slot))

    def mark_elapsed(self, a: int, b: int) -> float:
        ms = C.c_double()
        self._ck(self.lib.edmd_mark_elapsed(self.h, a, b, C.byref(ms)))
        return ms.value

    def timing(self, on=True):
        self._ck(self.lib.edmd_timing_enable(self.h, int(on)))

    def timing_get(self, which):
        ms = C.c_double(); n = C.c_longlong()
        self._ck(self.lib.edmd_timing_get(self.h, which, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def timing_reset(self):
        self._ck(self.lib.edmd_timing_reset(self.h))

    @property
    def launch_count(self):
        return self.lib.edmd_launch_count(self.h)

    def flagged_pairs(self) -> int:
        c = C.c_longlong()
        self._ck(self.lib.edmd_nb_stats(self.h, C.byref(c)))
        return c.value


class EngineGroup:
    """edmd_group_*: one process driving several GPUs (what the C++ driver `edmddna --gpus N` uses)."""

    def __init__(self, sys: EdmdSystem, n_gpus: int, devices=None):
        self.lib = load_library()
        self.h = C.c_void_p()
        p = Params(*[float(x) for x in sys.params])
        dev = None
        if devices is not None:
            self._dev = (C.c_int * n_gpus)(*devices)
            dev = self._dev
        self._ck(self.lib.edmd_group_create(C.byref(p), n_gpus, dev, C.byref(self.h)))
        self.sys = sys
        self.block = build_tetrads(sys, pinned=False)
        bp = None
        if sys.bp_atoms is not None:
            self._bp = np.ascontiguousarray(sys.bp_atoms, dtype=np.int32)
            bp = self._bp.ctypes.data_as(C.POINTER(C.c_int))
        self._ck(self.lib.edmd_group_upload_tetrads(self.h, self.block.ptr, sys.n, bp))

    def _ck(self, rc):
        if rc != 0:
            raise EngineError(f"edmd_b200 error {rc}: {self.lib.edmd_last_error().decode()}")

    def close(self):
        if self.h:
            self.lib.edmd_group_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def build_pairlist(self) -> int:
        npairs = C.c_longlong()
        self._ck(self.lib.edmd_group_build_pairlist(self.h, C.byref(npairs)))
        return npairs.value

    def set_rng(self, time0=1000000000, rank=1, calls_before=0):
        self._ck(self.lib.edmd_group_set_rng(self.h, time0, rank, calls_before))

    def step(self, nsteps=1, random_mode=0):
        self._ck(self.lib.edmd_group_step(self.h, nsteps, random_mode))

    def merge(self):
        self._ck(self.lib.edmd_group_merge(self.h))

    def get_state(self):
        self._ck(self.lib.edmd_group_get_state(self.h, self.block.ptr, self.sys.n))
        return self.block.crd.copy(), self.block.vel.copy()

    def energies(self) -> np.ndarray:
        out = np.zeros(4)
        self._ck(self.lib.edmd_group_get_energies(self.h, _ptr(out)))
        return out

    def sync(self):
        self._ck(self.lib.edmd_group_sync(self.h))


def _nb_stats(self) -> dict:
    """Counts from the last NB pass: flagged tetrad pairs, marked (owned) tetrads, 32 x 32 atom blocks the
    force pass evaluates, tetrad pairs selected by the bounding-sphere tests."""
    out = (C.c_longlong * 5)()
    self._ck(self.lib.edmd_nb_stats_ex(self.h, out))
    return dict(flagged_pairs=out[0], marked_tetrads=out[1], force_blocks=out[2], selected_pairs=out[3],
                sphere_blocks=out[4])


Engine.nb_stats = _nb_stats


def measure_fp64_peak(device=0):
    lib = load_library()
    tf = C.c_double(); mhz = C.c_double()
    rc = lib.edmd_measure_fp64_peak(device, C.byref(tf), C.byref(mhz))
    if rc != 0:
        raise EngineError(lib.edmd_last_error().decode())
    return tf.value, mhz.value
