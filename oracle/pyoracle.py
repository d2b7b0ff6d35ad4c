"""ctypes wrapper over the two CPU checkers (oracle/oracle_api.h).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (msc-project_b200/) never does.

    Oracle("port")       oracle/_build/libedmd_oracle.so — plain-C restatement (edmd_oracle.c)
    Oracle("reference")  oracle/_ref/libedmd_ref.so      — the unmodified reference sources

Both are built by ``make -C oracle port ref`` (``__graft_entry__.build()`` does it); the
reference build needs /root/reference and therefore only ever happens in the build container —
the resulting .so travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PD = C.POINTER(C.c_double)
PI = C.POINTER(C.c_int)
LL = C.c_longlong

PATHS = {"port": os.path.join(_HERE, "_build", "libedmd_oracle.so"),
         "reference": os.path.join(_HERE, "_ref", "libedmd_ref.so"),
         "reference_O0": os.path.join(_HERE, "_ref", "libedmd_ref_O0.so")}   # same sources at the shipped -O0 (timing only)
_LIBS = {}


def available(kind: str) -> bool:
    return os.path.exists(PATHS[kind])


def build(kind: str = "port", ref_root: str = "/root/reference") -> bool:
    target = "port" if kind == "port" else "ref"
    if kind != "port" and not os.path.isdir(ref_root):
        return available(kind)
    subprocess.run(["make", "-C", _HERE, target, f"REF={ref_root}"], check=True,
                   stdout=subprocess.DEVNULL)
    return True


def _load(kind: str):
    if kind in _LIBS:
        return _LIBS[kind]
    if not available(kind):
        raise FileNotFoundError(f"{PATHS[kind]} not built (make -C oracle {'port' if kind == 'port' else 'ref'})")
    lib = C.CDLL(PATHS[kind], mode=C.RTLD_LOCAL)
    S = C.c_void_p
    sig = {
        "orc_kind": (C.c_char_p, []),
        "orc_create": (S, [C.c_int, PI, PI]),
        "orc_destroy": (None, [S]),
        "orc_set_params": (None, [S, PD]),
        "orc_set_nb_consts": (C.c_int, [S, C.c_double, C.c_double]),
        "orc_set_static": (None, [S, C.c_int, PD, PD, PD, PD, PD]),
        "orc_set_state": (None, [S, PD, PD]),
        "orc_get_state": (None, [S, PD, PD]),
        "orc_get_forces": (None, [S, PD, PD, PD, PD, PD, PD]),
        "orc_set_forces": (None, [S, PD, PD, PD]),
        "orc_ed_forces": (None, [S, C.c_int]),
        "orc_random_terms": (None, [S, C.c_int, C.c_int]),
        "orc_nb_forces": (None, [S, C.c_int, C.c_int]),
        "orc_update_velocities": (None, [S, C.c_int]),
        "orc_update_coordinates": (None, [S, C.c_int]),
        "orc_qcp": (C.c_double, [PD, PD, C.c_int, PD]),
        "orc_set_time": (None, [C.c_long]),
        "orc_rng_calls": (C.c_long, []),
        "orc_set_rng_calls": (C.c_int, [C.c_long]),
        "orc_build_pairs": (LL, [S]),
        "orc_num_pairs": (LL, [S]),
        "orc_get_pairs": (None, [S, PI]),
        "orc_set_pairs": (None, [S, LL, PI]),
        "orc_forces": (None, [S, C.c_int, C.c_int]),
        "orc_step": (None, [S, C.c_int, C.c_int, C.c_int]),
        "orc_time_nb": (C.c_double, [S, LL, LL, C.c_int]),
        "orc_merge": (None, [S, PI]),
        "orc_energies": (None, [S, PD]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if kind == "port":
        lib.orc_glibc_rand_stream.restype = None
        lib.orc_glibc_rand_stream.argtypes = [C.c_uint, C.c_int, PI]
    _LIBS[kind] = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(PD)


class Oracle:
    """CPU reference of one EdmdSystem (an msc-project_b200 EdmdSystem instance)."""

    def __init__(self, sys, kind: str = "port"):
        self.lib = _load(kind)
        self.kind = kind
        self.sys = sys
        self.n = sys.n
        self._na = np.ascontiguousarray(sys.num_atoms, dtype=np.int32)
        self._nev = np.ascontiguousarray(sys.num_evecs, dtype=np.int32)
        self.off3 = sys.off3
        self.h = self.lib.orc_create(self.n, self._na.ctypes.data_as(PI), self._nev.ctypes.data_as(PI))
        self.set_params(sys.params)
        for t in range(self.n):
            ps = sys.psets[int(sys.param_id[t])]
            self.lib.orc_set_static(self.h, t, _p(ps.avg), _p(ps.masses3), _p(ps.abq), _p(ps.evals), _p(ps.evecs))
        self.set_state(sys.crd, sys.vel)

    def close(self):
        if self.h:
            self.lib.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, params):
        p = np.ascontiguousarray(params, dtype=np.float64)
        self.lib.orc_set_params(self.h, _p(p))

    def set_nb_consts(self, krep=100.0, qfac=0.0):
        return self.lib.orc_set_nb_consts(self.h, krep, qfac)

    def set_state(self, crd=None, vel=None):
        c = None if crd is None else np.ascontiguousarray(crd, dtype=np.float64)
        v = None if vel is None else np.ascontiguousarray(vel, dtype=np.float64)
        self.lib.orc_set_state(self.h, _p(c), _p(v))

    def get_state(self):
        tot = int(self.off3[-1])
        crd = np.zeros(tot); vel = np.zeros(tot)
        self.lib.orc_get_state(self.h, _p(crd), _p(vel))
        return crd, vel

    def get_forces(self):
        tot = int(self.off3[-1]); n = self.n
        ed = np.zeros(tot); eed = np.zeros(n); rnd = np.zeros(tot); nb = np.zeros(tot)
        enb = np.zeros(2 * n); temp = np.zeros(n)
        self.lib.orc_get_forces(self.h, _p(ed), _p(eed), _p(rnd), _p(nb), _p(enb), _p(temp))
        return dict(ed=ed, ed_energy=eed, rnd=rnd, nb=nb, nb_energy=enb.reshape(n, 2), temperature=temp)

    def set_forces(self, ed=None, rnd=None, nb=None):
        a = [None if x is None else np.ascontiguousarray(x, dtype=np.float64) for x in (ed, rnd, nb)]
        self.lib.orc_set_forces(self.h, _p(a[0]), _p(a[1]), _p(a[2]))

    # the five methods
    def ed_forces(self, t=None):
        for k in (range(self.n) if t is None else [t]):
            self.lib.orc_ed_forces(self.h, k)

    def random_terms(self, t=None, rank=1):
        for k in (range(self.n) if t is None else [t]):
            self.lib.orc_random_terms(self.h, k, rank)

    def nb_forces_pair(self, t1, t2):
        self.lib.orc_nb_forces(self.h, t1, t2)

    def update_velocities(self, t=None):
        for k in (range(self.n) if t is None else [t]):
            self.lib.orc_update_velocities(self.h, k)

    def update_coordinates(self, t=None):
        for k in (range(self.n) if t is None else [t]):
            self.lib.orc_update_coordinates(self.h, k)

    def qcp(self, ref_xyz, x_xyz):
        a = np.ascontiguousarray(ref_xyz, dtype=np.float64).reshape(-1)
        b = np.ascontiguousarray(x_xyz, dtype=np.float64).reshape(-1)
        rot = np.zeros(9)
        rms = self.lib.orc_qcp(_p(a), _p(b), a.size // 3, _p(rot))
        return rot.reshape(3, 3), rms

    # rng pinning
    def set_time(self, t0):
        self.lib.orc_set_time(int(t0))

    def rng_calls(self):
        return self.lib.orc_rng_calls()

    def set_rng_calls(self, calls):
        return self.lib.orc_set_rng_calls(int(calls))

    # protocol
    def build_pairs(self) -> int:
        return self.lib.orc_build_pairs(self.h)

    def get_pairs(self) -> np.ndarray:
        npairs = self.lib.orc_num_pairs(self.h)
        out = np.zeros((npairs, 2), dtype=np.int32)
        if npairs:
            self.lib.orc_get_pairs(self.h, out.ctypes.data_as(PI))
        return out

    def set_pairs(self, pairs):
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        self.lib.orc_set_pairs(self.h, pairs.shape[0], pairs.ctypes.data_as(PI))

    def forces(self, random_on=False, nthreads=1):
        self.lib.orc_forces(self.h, int(random_on), nthreads)

    def step(self, nsteps=1, random_on=False, nthreads=1):
        self.lib.orc_step(self.h, nsteps, int(random_on), nthreads)

    def time_nb(self, p0, p1, nthreads=1) -> float:
        return self.lib.orc_time_nb(self.h, p0, p1, nthreads)

    def merge(self):
        bp = np.ascontiguousarray(self.sys.bp_atoms, dtype=np.int32)
        self.lib.orc_merge(self.h, bp.ctypes.data_as(PI))

    def energies(self):
        e = np.zeros(4)
        self.lib.orc_energies(self.h, _p(e))
        return e


def glibc_rand_stream(seed: int, count: int) -> np.ndarray:
    lib = _load("port")
    out = np.zeros(count, dtype=np.int32)
    lib.orc_glibc_rand_stream(seed & 0xffffffff, count, out.ctypes.data_as(PI))
    return out
