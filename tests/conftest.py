import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def edmd():
    """The product package (directory name has a hyphen, hence importlib)."""
    return importlib.import_module("msc-project_b200")


@pytest.fixture(scope="session")
def po():
    """oracle/pyoracle.py — the CPU checkers (tests are allowed to use them)."""
    import pyoracle
    if not pyoracle.available("port"):
        pyoracle.build("port")
    return pyoracle


@pytest.fixture(scope="session")
def have_ref(po):
    if not po.available("reference") and os.path.isdir("/root/reference"):
        po.build("reference")
    return po.available("reference")
