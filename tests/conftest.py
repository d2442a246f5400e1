import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_ops():
    from oracle import cpu_ops
    cpu_ops.build()
    return cpu_ops


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def ref_ext(name):
    """Import one of the compiled UNMODIFIED reference extensions from oracle/_ref (GPU box only)."""
    import importlib
    import torch  # noqa: F401  (libtorch must be loaded first)
    d = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(d, name + ".so")):
        return None
    if d not in sys.path:
        sys.path.insert(0, d)
    try:
        return importlib.import_module(name)
    except Exception:  # noqa: BLE001
        return None
