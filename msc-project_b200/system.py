"""Host-side description of an EDMD system and synthetic-input generators.

An :class:`EdmdSystem` is what the reference's ``IO::read_Prm`` / ``read_Crd`` /
``initialise_Tetrad_Crds`` produce (/root/reference/src/io.cpp:120-288): per-tetrad static
parameters (average structure, masses, (A,B,q), eigenvalues, eigenvectors) and dynamic
state (coordinates, velocities), plus the eight EDMD parameters in internal units
(io.cpp:85-91).  Static parameters are stored once per *parameter set*; ``param_id`` maps a
tetrad to its set, so replicated DNA of 10^5-10^6 tetrads does not need 90 KB per tetrad of
host memory.  Per-tetrad arrays are xyz-interleaved, tetrads concatenated (tetrad ``t`` owns
``[off3[t], off3[t+1])``), exactly the reference's ``Tetrad`` field layout.

No GPU or oracle code is imported here.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

TIMEFAC = 20.455      # Constants.timefac, edmd.cpp:26
BOLTZMANN = 0.002     # Constants.Boltzmann, edmd.cpp:25
GENERATOR_SEED = 20261019

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

# Amber residue atom order found in test/GC90_6c.crd: even base pairs are
# DG (33 atoms) then DC (30 atoms), odd base pairs DC then DG.
_DG = ("P O O O C H H C H O C H N C H N C C O N H C N H H N C C H C H H O").split()
_DC = ("P O O O C H H C H O C H N C H C H C N H H N C O C H C H H O").split()
_MASS = {"P": 30.97, "O": 16.00, "C": 12.01, "H": 1.008, "N": 14.01}
assert len(_DG) == 33 and len(_DC) == 30


def config_params(dt_ps=0.002, gamma_ps=2.0, tautp_ps=0.2, temperature=300.0,
                  mole_cutoff=30.0, atom_cutoff=10.0, mole_least=5.0) -> np.ndarray:
    """The eight EDMD parameters after the unit conversion of io.cpp:85-91."""
    return np.array([dt_ps * TIMEFAC, gamma_ps / TIMEFAC, tautp_ps * TIMEFAC, temperature,
                     BOLTZMANN * temperature, mole_cutoff, atom_cutoff, mole_least], dtype=np.float64)


@dataclass
class ParamSet:
    num_atoms: int
    num_evecs: int
    avg: np.ndarray       # [3N]
    masses3: np.ndarray   # [3N] each mass replicated x3 (io.cpp:149-152)
    abq: np.ndarray       # [3N]
    evals: np.ndarray     # [nev]
    evecs: np.ndarray     # [nev, 3N] C-contiguous


@dataclass
class EdmdSystem:
    psets: list
    param_id: np.ndarray      # [n] int32
    bp_atoms: np.ndarray      # [n+3] int32, may be None
    crd: np.ndarray           # flat [sum 3N]
    vel: np.ndarray           # flat [sum 3N]
    params: np.ndarray = field(default_factory=config_params)
    name: str = "system"

    @property
    def n(self) -> int:
        return int(self.param_id.shape[0])

    @property
    def num_atoms(self) -> np.ndarray:
        return np.array([self.psets[p].num_atoms for p in self.param_id], dtype=np.int32)

    @property
    def num_evecs(self) -> np.ndarray:
        return np.array([self.psets[p].num_evecs for p in self.param_id], dtype=np.int32)

    @property
    def off3(self) -> np.ndarray:
        o = np.zeros(self.n + 1, dtype=np.int64)
        np.cumsum(3 * self.num_atoms.astype(np.int64), out=o[1:])
        return o

    def copy(self) -> "EdmdSystem":
        return EdmdSystem(self.psets, self.param_id.copy(),
                          None if self.bp_atoms is None else self.bp_atoms.copy(),
                          self.crd.copy(), self.vel.copy(), self.params.copy(), self.name)

    def subset(self, tetrads) -> "EdmdSystem":
        """A system made of the listed tetrads (no base-pair bookkeeping: merge unavailable)."""
        tetrads = np.asarray(tetrads, dtype=np.int64)
        off = self.off3
        idx = np.concatenate([np.arange(off[t], off[t + 1]) for t in tetrads])
        return EdmdSystem(self.psets, self.param_id[tetrads].copy(), None, self.crd[idx].copy(),
                          self.vel[idx].copy(), self.params.copy(), self.name + "-subset")


def _gc90():
    z = np.load(os.path.join(_DATA, "gc90_6c.npz"))
    bp_atoms = z["bp_atoms"].astype(np.int32)
    xyz = z["xyz"].reshape(-1, 3)
    assert np.all(bp_atoms == 63)
    return bp_atoms, xyz.reshape(bp_atoms.shape[0], 63, 3)


def _bp_masses(bp_index: int) -> np.ndarray:
    names = (_DG + _DC) if bp_index % 2 == 0 else (_DC + _DG)
    return np.array([_MASS[e] for e in names], dtype=np.float64)


def _make_psets(shapes: np.ndarray, nev: int, seed: int, perturb: float) -> list:
    """Synthetic .prm content (the reference's test/GC90c12.prm is absent):
    avg = tetrad shape + N(0, perturb) Angstrom, standard atomic masses in Amber order,
    charges U(-0.8, 0.8), orthonormal eigenvectors (QR of a seeded Gaussian matrix),
    decreasing positive eigenvalues."""
    rng = np.random.default_rng(seed)
    psets = []
    for p, shape in enumerate(shapes):
        na = shape.shape[0]
        avg = shape.reshape(-1) + rng.normal(0.0, perturb, 3 * na)
        m = np.concatenate([_bp_masses(p + k) for k in range(4)])
        abq = np.zeros(3 * na)
        abq[2::3] = rng.uniform(-0.8, 0.8, na)
        q, _ = np.linalg.qr(rng.normal(size=(3 * na, nev)))
        evecs = np.ascontiguousarray(q.T)
        evals = np.sort(rng.uniform(0.5, 8.0, nev))[::-1].copy()
        psets.append(ParamSet(na, nev, avg, np.repeat(m, 3), abq, evals, evecs))
    return psets


def make_gc90(nev: int = 12, seed: int = GENERATOR_SEED, perturb: float = 0.05) -> EdmdSystem:
    """config[0]: the reference's test/GC90_6c.crd (93 bp, 90 tetrads of 252 atoms) with a
    synthetic parameter file; every tetrad has its own parameter set, as a real .prm would."""
    bp_atoms, bp = _gc90()
    n = bp.shape[0] - 3
    shapes = np.stack([bp[t:t + 4].reshape(-1, 3) for t in range(n)])
    psets = _make_psets(shapes, nev, seed, perturb)
    crd = np.concatenate([shapes[t].reshape(-1) for t in range(n)])
    return EdmdSystem(psets, np.arange(n, dtype=np.int32), bp_atoms, crd, np.zeros_like(crd),
                      config_params(), "GC90")


def _rot_y(theta: np.ndarray) -> np.ndarray:
    c, s = np.cos(theta), np.sin(theta)
    r = np.zeros(theta.shape + (3, 3))
    r[..., 0, 0] = c; r[..., 0, 2] = s; r[..., 1, 1] = 1.0; r[..., 2, 0] = -s; r[..., 2, 2] = c
    return r


def make_ring(n_tetrads: int, kind: str = "ring", nev: int = 12, seed: int = GENERATOR_SEED,
              perturb: float = 0.05, n_psets: int = 10, laps: int = 2, lap_gap: float = 9.0,
              mole_cutoff: float = 30.0) -> EdmdSystem:
    """Synthetic replicated GC DNA: a closed ring of ``n_tetrads`` base pairs (+3 repeated).

    Base pair k is the reference ring's base pair k mod 90 (test/GC90_6c.crd lies in the x-z
    plane around the y axis, radius 42.98 A), moved rigidly from ring angle 2*pi*(k mod 90)/90
    at radius R90 to angle theta_k at radius R = R90 * n/90 (same rise per base pair), so the
    local helix geometry of the shipped structure is kept.  n_tetrads must be a multiple of 10
    (helical repeat) for the ring to close.

    kind="ring": planar ring — with the shipped cutoffs every tetrad has 10 partners and no
    atom pair is closer than sqrt(2) A (NB forces are exactly zero, as in GC90).
    kind="coil": the chain makes ``laps`` turns around the y axis while its y offset follows
    lap_gap*sin(theta/laps): neighbouring laps interpenetrate where the offset crosses zero,
    which gives thousands of overlapping atom pairs — non-zero, partly clipped NB forces.

    Tetrads share ``n_psets`` parameter sets (param_id = t mod n_psets)."""
    if n_tetrads % 10 != 0 or n_tetrads < 20:
        raise ValueError("n_tetrads must be a multiple of 10 and >= 20")
    bp_atoms90, bp90 = _gc90()
    n = n_tetrads
    centre = bp90[:90].reshape(-1, 3).mean(axis=0)
    r90 = np.linalg.norm((bp90[:90].mean(axis=1) - centre)[:, [0, 2]], axis=1).mean()
    # angle of shipped base pair j around the y axis
    c90 = bp90[:90].mean(axis=1) - centre
    ang90 = np.arctan2(c90[:, 0], c90[:, 2])          # x = R sin, z = R cos
    ang90 = np.unwrap(ang90)
    k = np.arange(n + 3)
    j = (k % n) % 90
    if kind == "ring":
        nl = 1
    elif kind == "coil":
        nl = laps
    else:
        raise ValueError(kind)
    radius = r90 * n / (90.0 * nl)
    theta = 2.0 * np.pi * nl * (k % n) / n
    local = np.einsum("kab,knb->kna", _rot_y(-ang90[j]), bp90[j] - centre)   # bp at angle 0
    local[:, :, 2] += radius - r90
    if kind == "coil":
        local[:, :, 1] += (lap_gap * np.sin(theta / nl))[:, None]
    bp = np.einsum("kab,knb->kna", _rot_y(theta), local)
    # tetrad coordinates
    crd = np.empty((n, 252, 3))
    for q in range(4):
        crd[:, 63 * q:63 * (q + 1), :] = bp[q:q + n]
    shapes = np.stack([bp90[p:p + 4].reshape(-1, 3) for p in range(n_psets)])
    psets = _make_psets(shapes, nev, seed, perturb)
    params = config_params(mole_cutoff=mole_cutoff)
    return EdmdSystem(psets, (np.arange(n) % n_psets).astype(np.int32),
                      np.full(n + 3, 63, dtype=np.int32), crd.reshape(-1), np.zeros(n * 756),
                      params, f"{kind}{n}")


def make_ragged(n_tetrads: int = 40, seed: int = GENERATOR_SEED, nevs=(3, 12, 20, 29), lap_gap: float = 9.0,
                bp_lo: int = 58, bp_hi: int = 64, full_tetrad: bool = True) -> EdmdSystem:
    """Edge-case system: base pairs with DIFFERENT atom counts (58..64, one 64-atom base pair so a
    tetrad can reach the 256-atom limit), every tetrad with its own parameter set, and eigenvector
    counts below, at and above the 12-per-chunk register tile of the ED kernel.  Geometry: the
    double coil of make_ring (real atom overlaps).  ``bp_lo/bp_hi`` (exclusive) bound the atoms kept per
    base pair; ``full_tetrad=False`` leaves out the 256-atom tetrad, so small base pairs give a system
    whose atom stride is below the engine maximum."""
    base = make_ring(n_tetrads, kind="coil", laps=2, lap_gap=lap_gap)
    rng = np.random.default_rng(seed + 1)
    n = n_tetrads
    bp_atoms = rng.integers(bp_lo, bp_hi, size=n + 3).astype(np.int32)
    if full_tetrad:
        bp_atoms[5:9] = 64                  # tetrad 5 = 4 x 64 = 256 atoms (engine maximum)
    bp_atoms[n:n + 3] = bp_atoms[0:3]       # the ring repeats its first three base pairs
    crd_bp = base.crd.reshape(n, 4, 63, 3)
    extra = rng.normal(scale=0.5, size=(n + 3, 1, 3))   # 64th atom: near atom 0 of the base pair

    def bp_xyz(b):
        t, q = (b, 0) if b < n else (n - 1, b - n + 1)
        full = np.concatenate([crd_bp[t, q], crd_bp[t, q, :1] + extra[b % n] + 1.5], axis=0)
        return full[:bp_atoms[b]]

    bps = [bp_xyz(b) for b in range(n + 3)]
    for i in range(3):
        bps[n + i] = bps[i]
    psets, crd = [], []
    for t in range(n):
        shape = np.concatenate(bps[t:t + 4], axis=0)
        na = shape.shape[0]
        nev = int(nevs[t % len(nevs)])
        avg = shape.reshape(-1) + rng.normal(0.0, 0.05, 3 * na)
        m = rng.choice([1.008, 12.01, 14.01, 16.0, 30.97], size=na)
        abq = np.zeros(3 * na); abq[2::3] = rng.uniform(-0.8, 0.8, na)
        q, _ = np.linalg.qr(rng.normal(size=(3 * na, nev)))
        psets.append(ParamSet(na, nev, avg, np.repeat(m, 3), abq,
                              np.sort(rng.uniform(0.5, 8.0, nev))[::-1].copy(), np.ascontiguousarray(q.T)))
        crd.append(shape.reshape(-1))
    crd = np.concatenate(crd)
    return EdmdSystem(psets, np.arange(n, dtype=np.int32), bp_atoms, crd, np.zeros_like(crd),
                      config_params(), f"ragged{n}")
