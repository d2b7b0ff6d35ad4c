"""The C-ABI library loads and exports every symbol include/edmd_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "edmd_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(edmd_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(edmd):
    lib = edmd.load_library()
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/edmd_b200.h but not exported"
    assert lib.edmd_abi_version() == 2


def test_python_binding_covers_the_header(edmd):
    lib = edmd.load_library()
    assert set(_declared()) == set(lib._edmd_symbols)


def test_tetrad_struct_matches_reference_layout(edmd):
    # tetrad.hpp:36-60 on LP64: 2 ints, 8 pointers, 1 double, 2 pointers = 96 bytes
    T = edmd.Tetrad
    assert ctypes.sizeof(T) == 96
    assert T.avg.offset == 8 and T.eigenvectors.offset == 40 and T.temperature.offset == 72
    assert T.velocities.offset == 80 and T.coordinates.offset == 88


def test_no_cpu_fallback(edmd):
    """Without a CUDA device the engine must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(edmd.EngineError, match="no CPU fallback"):
        edmd.Engine(edmd.make_ring(20))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "msc-project_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in src and "edmd_oracle" not in src and "libedmd_ref" not in src, f
