"""Generate tests/golden/*.npz from the UNMODIFIED reference build (oracle/_ref/libedmd_ref.so).

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The vectors pin the plain-C port (oracle/edmd_oracle.c) on machines where the reference build
is absent, and are the fixed targets of the GPU parity tests.

golden_gc90.npz   config[0] geometry (test/GC90_6c.crd, synthetic .prm, shipped cutoffs):
                  pair list; per-function outputs after one force evaluation; state after
                  5 steps + merge (random off) and after 3 steps (random on, time pinned).
golden_coil40.npz 40-tetrad double coil with interpenetrating laps (non-zero, clipped NB forces).
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle as po  # noqa: E402

pkg = importlib.import_module("msc-project_b200")
T0 = 1000000000
STRIDE = 7


def run(system, path):
    out = {}
    o = po.Oracle(system, "reference")
    assert o.lib.orc_kind() == b"reference"
    out["npairs"] = np.array([o.build_pairs()])
    out["pairs"] = o.get_pairs()
    o.forces(random_on=False)
    f = o.get_forces()
    c, v = o.get_state()
    out.update(f1_ed=f["ed"], f1_ed_energy=f["ed_energy"], f1_nb=f["nb"], f1_nb_energy=f["nb_energy"], f1_crd=c)
    o.update_velocities(); o.update_coordinates()
    c, v = o.get_state()
    out.update(s1_crd=c, s1_vel=v, s1_temp=o.get_forces()["temperature"])
    o.step(4)
    o.merge()
    c, v = o.get_state()
    out.update(s5m_crd=c, s5m_vel=v, s5m_energies=o.energies())
    # stochastic: fresh system, pinned time; the reference's RNG_Seed counter cannot be reset, so
    # record how many calls this process had made before
    o2 = po.Oracle(system, "reference")
    o2.set_time(T0)
    out["rng_calls_before"] = np.array([o2.rng_calls()])
    o2.build_pairs()
    o2.step(3, random_on=True)
    c, v = o2.get_state()
    out.update(r3_crd=c, r3_vel=v, r3_rnd=o2.get_forces()["rnd"])
    # keep the fixtures small: every STRIDE-th element of the big arrays plus their plain sums
    small = {}
    for k, v in out.items():
        v = np.asarray(v)
        if v.ndim == 1 and v.size > 1000:
            small[k] = v[::STRIDE].copy()
            small[k + "_sum"] = np.array([v.sum()])
        else:
            small[k] = v
    out = small
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    # NOTE order matters for rng_calls_before: gc90 first (0 calls), coil second (3*90 calls).
    run(pkg.make_gc90(), os.path.join(here, "golden_gc90.npz"))
    run(pkg.make_ring(40, kind="coil", laps=2, lap_gap=9.0), os.path.join(here, "golden_coil40.npz"))
