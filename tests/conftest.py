import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def scenes():
    import scenes as S
    return S


def _cuda_device_present():
    """True when librast_b200.so loads and can create a context on device 0 (no torch import: this runs at collection)."""
    try:
        import ctypes as C
        from rasteriser_b200 import _lib
        lib = _lib.load()
        h = C.c_void_p()
        if lib.rast_create(0, C.byref(h)) != 0:
            return False
        lib.rast_destroy(h)
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without a GPU skips the gpu-marked tests instead of failing in rast_create.
    On a GPU box nothing is skipped (a missing library there must fail loudly, not skip)."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items or os.path.exists("/dev/nvidia0") or _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device: the frame path has no CPU fallback")
    for it in gpu_items:
        it.add_marker(skip)
