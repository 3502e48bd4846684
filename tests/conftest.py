import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _have_cuda_device():
    try:
        import ctypes
        rt = ctypes.CDLL("libcudart.so")
        n = ctypes.c_int(0)
        return rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        try:
            import torch
            return torch.cuda.is_available()
        except Exception:
            return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a machine without a GPU skips the gpu-marked tests instead of failing in wm_create."""
    if _have_cuda_device():
        return
    skip = pytest.mark.skip(reason="no CUDA device (gpu-marked tests run on the B200 box: pytest -m gpu)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def wmlib():
    import wumingpic2d_b200 as wm
    return wm.load_library()
