"""Input-deck formats of the reference (config.txt, .prm, .crd) — writers and readers.

Formats follow /root/reference/src/io.cpp:60-232:
  config.txt  17 positional ``key = value`` lines (io.cpp:77-100)
  .prm        num_Tetrads; per tetrad: ``num_Atoms num_Evecs``; 3N avg; N masses; 3N (A,B,q);
              per eigenvector: eigenvalue then 3N coefficients (io.cpp:128-172)
  .crd        num_BP; num_BP atom counts; per bp 3n coordinates (+3n velocities iff irest != 0)
              (io.cpp:197-223)
Numbers are written with 17 significant digits so a deck round-trips bit-exactly.
"""
from __future__ import annotations

import os

import numpy as np

from .system import EdmdSystem, ParamSet, TIMEFAC, BOLTZMANN

CONFIG_KEYS = ["irest", "nsteps", "ntsync", "ntwt", "ntpr", "dt", "gamma", "tautp", "temperature",
               "mole_Cutoff", "atom_Cutoff", "mole_Least", "prm_File", "crd_File", "energy_File",
               "trj_File", "new_Crd_File"]


def _fmt(a) -> str:
    return " ".join(repr(float(x)) for x in np.asarray(a).reshape(-1))


def write_config(path, sys: EdmdSystem, prm_file, crd_file, energy_file="./energies.eng",
                 trj_file="./trajectory.trj", new_crd_file="./crd.crd", irest=0, nsteps=100,
                 ntsync=100, ntwt=100, ntpr=1000):
    p = sys.params
    vals = [irest, nsteps, ntsync, ntwt, ntpr, repr(float(p[0] / TIMEFAC)), repr(float(p[1] * TIMEFAC)),
            repr(float(p[2] / TIMEFAC)),
            repr(float(p[3])), repr(float(p[5])), repr(float(p[6])), repr(float(p[7])),
            prm_file, crd_file, energy_file, trj_file, new_crd_file]
    with open(path, "w") as f:
        for k, v in zip(CONFIG_KEYS, vals):
            f.write(f"{k:<12s} = {v}\n")


def write_prm(path, sys: EdmdSystem):
    with open(path, "w") as f:
        f.write(f"{sys.n}\n")
        for t in range(sys.n):
            ps = sys.psets[int(sys.param_id[t])]
            f.write(f"{ps.num_atoms} {ps.num_evecs}\n")
            f.write(_fmt(ps.avg) + "\n")
            f.write(_fmt(ps.masses3[::3]) + "\n")
            f.write(_fmt(ps.abq) + "\n")
            for j in range(ps.num_evecs):
                f.write(repr(float(ps.evals[j])) + "\n")
                f.write(_fmt(ps.evecs[j]) + "\n")


def write_crd(path, sys: EdmdSystem, with_velocities=False):
    """Base pair b is taken from the first tetrad that contains it (b < n: tetrad b, bp slot 0;
    the last three from tetrad n-1)."""
    assert sys.bp_atoms is not None
    n = sys.n
    nbp = n + 3
    off = sys.off3
    bp = sys.bp_atoms

    def bp_values(arr, b):
        t = min(b, n - 1)
        start = off[t] + 3 * int(bp[t:b].sum())
        return arr[start:start + 3 * int(bp[b])]

    with open(path, "w") as f:
        f.write(f"{nbp}\n")
        for b in range(nbp):
            f.write(f"{int(bp[b])}\n")
        for b in range(nbp):
            f.write(_fmt(bp_values(sys.crd, b)) + "\n")
            if with_velocities:
                f.write(_fmt(bp_values(sys.vel, b)) + "\n")


def read_deck(config_path) -> tuple:
    """Python restatement of IO::read_Cofig/read_Prm/read_Crd/initialise_Tetrad_Crds -> (EdmdSystem, cfg)."""
    base = os.path.dirname(os.path.abspath(config_path))
    cfg = {}
    with open(config_path) as f:
        for k, line in zip(CONFIG_KEYS, f):
            cfg[k] = line.split()[2]
    for k in CONFIG_KEYS[:5]:
        cfg[k] = int(cfg[k])
    for k in CONFIG_KEYS[5:12]:
        cfg[k] = float(cfg[k])
    params = np.array([cfg["dt"] * TIMEFAC, cfg["gamma"] / TIMEFAC, cfg["tautp"] * TIMEFAC, cfg["temperature"],
                       BOLTZMANN * cfg["temperature"], cfg["mole_Cutoff"], cfg["atom_Cutoff"], cfg["mole_Least"]])
    tok = open(os.path.join(base, cfg["prm_File"])).read().split()
    pos = 0
    n = int(tok[pos]); pos += 1
    psets = []
    for _ in range(n):
        na, nev = int(tok[pos]), int(tok[pos + 1]); pos += 2
        avg = np.array(tok[pos:pos + 3 * na], dtype=np.float64); pos += 3 * na
        m = np.array(tok[pos:pos + na], dtype=np.float64); pos += na
        abq = np.array(tok[pos:pos + 3 * na], dtype=np.float64); pos += 3 * na
        evals = np.zeros(nev); evecs = np.zeros((nev, 3 * na))
        for j in range(nev):
            evals[j] = float(tok[pos]); pos += 1
            evecs[j] = np.array(tok[pos:pos + 3 * na], dtype=np.float64); pos += 3 * na
        psets.append(ParamSet(na, nev, avg, np.repeat(m, 3), abq, evals, evecs))
    crd_path = cfg["crd_File"] if cfg["irest"] == 0 else cfg["new_Crd_File"]
    tok = open(os.path.join(base, crd_path)).read().split()
    nbp = int(tok[0])
    bp = np.array(tok[1:1 + nbp], dtype=np.int32)
    vals = np.array(tok[1 + nbp:], dtype=np.float64)
    displs = np.zeros(nbp + 1, dtype=np.int64)
    np.cumsum(3 * bp.astype(np.int64), out=displs[1:])
    flat_c = np.zeros(displs[-1]); flat_v = np.zeros(displs[-1])
    pos = 0
    for b in range(nbp):
        k = 3 * int(bp[b])
        flat_c[displs[b]:displs[b + 1]] = vals[pos:pos + k]; pos += k
        if cfg["irest"] != 0:
            flat_v[displs[b]:displs[b + 1]] = vals[pos:pos + k]; pos += k
    crd = np.concatenate([flat_c[displs[t]:displs[t + 4]] for t in range(n)])
    vel = np.concatenate([flat_v[displs[t]:displs[t + 4]] for t in range(n)])
    return EdmdSystem(psets, np.arange(n, dtype=np.int32), bp, crd, vel, params, "deck"), cfg
