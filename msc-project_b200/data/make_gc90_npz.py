"""Pack the reference's test input test/GC90_6c.crd (93 bp x 63 atoms, format io.cpp:197-223)
into msc-project_b200/data/gc90_6c.npz so that config[0] (GC90) and the synthetic ring
generator work on the GPU box, where /root/reference does not exist.

Run once in the build container:  python msc-project_b200/data/make_gc90_npz.py
"""
import os
import sys
import numpy as np

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/test/GC90_6c.crd"
tok = open(src).read().split()
nbp = int(tok[0])
bp_atoms = np.array([int(x) for x in tok[1:1 + nbp]], dtype=np.int32)
xyz = np.array(tok[1 + nbp:], dtype=np.float64)
assert xyz.size == 3 * bp_atoms.sum()
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gc90_6c.npz")
np.savez_compressed(out, bp_atoms=bp_atoms, xyz=xyz)
print("wrote", out, nbp, "bp", xyz.size // 3, "atoms")
