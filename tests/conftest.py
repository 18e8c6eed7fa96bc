import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


HAVE_GPU = _have_gpu()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no GPU in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def zg():
    """The product, built if needed.  Import failure is a test failure, not a skip."""
    lib = os.path.join(ROOT, "zignal_b200", "libzignal_b200.so")
    if not os.path.exists(lib):
        import runpy
        runpy.run_path(os.path.join(ROOT, "zignal_b200", "build.py"))["build"]()
    import zignal_b200
    return zignal_b200


def _ref_so(name):
    import subprocess
    so = os.path.join(ROOT, "oracle", "_ref", name)
    if not os.path.exists(so) and os.path.exists("/root/reference/test/benchmark.cpp"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return so


@pytest.fixture(scope="session")
def ref_lib():
    """oracle/_ref/libzg_ref_custom.so: the hand-written loops of the reference's own benchmark.cpp compiled where it
    lies (checker only; linked without the product)."""
    import ctypes
    lib = ctypes.CDLL(_ref_so("libzg_ref_custom.so"))
    lib.zg_ref_sum_dirac_custom.restype = ctypes.c_float
    return lib


@pytest.fixture(scope="session")
def ref_flow_lib(zg):
    """oracle/_ref/libzg_ref_flow.so: the reference's make_flow() graphs (the same unmodified translation unit)
    compiled against this repository's flowz shim and linked with libzignal_b200 -- the drop-in check."""
    import ctypes
    lib = ctypes.CDLL(_ref_so("libzg_ref_flow.so"))
    lib.zg_ref_sum_dirac_flow.restype = ctypes.c_float
    return lib
