"""msc-project_b200 — B200-native EDMD force-and-integrate engine (host-side Python mirror).

The product is the CUDA library ``libedmd_b200.so`` behind the C ABI of
``include/edmd_b200.h``; this package is the thin ctypes layer tests, ``bench.py`` and the
multi-GPU driver use.  Importing it never touches ``oracle/``; constructing an
:class:`Engine` without the CUDA library or without a GPU raises — there is no CPU fallback.

The directory name contains a hyphen, so import it with
``importlib.import_module("msc-project_b200")`` (see ``tests/conftest.py``).
"""
from .system import (EdmdSystem, ParamSet, config_params, make_gc90, make_ring, make_ragged,  # noqa: F401
                     TIMEFAC, BOLTZMANN, GENERATOR_SEED)
from .engine import Engine, EngineGroup, EngineError, load_library, library_path, Tetrad, build_tetrads  # noqa: F401
