import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import loader as oracle_loader  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
REF_IMAGES = os.path.join(GOLDEN_DIR, "reference_images")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Make sure the native libraries exist (the driver runs build() first; this covers a bare pytest)."""
    from vviewer_b200 import capi
    need = [capi.HOST_LIB, capi.CUDA_LIB, oracle_loader.ORACLE_LIB]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()
    return capi


@pytest.fixture(scope="session")
def capi(built):
    return built


@pytest.fixture(scope="session")
def oracle_lib(capi):
    return oracle_loader.load_oracle()


@pytest.fixture(scope="session")
def cuda_lib(capi):
    return capi.load_cuda()
