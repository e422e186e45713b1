import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def built():
    """Build native pieces once per session if they are missing (CPU box: nvcc cross-compiles)."""
    from lld_slam_b200 import capi
    if not (os.path.exists(capi.LIB_PATH) and os.path.exists(capi.ORACLE_PATH)
            and os.path.exists(os.path.join(ROOT, "tests", "hostcheck", "libdevmath_host.so"))):
        import __graft_entry__ as g
        g.build()
    return True


@pytest.fixture(scope="session")
def gpu_ctx(built):
    from lld_slam_b200 import capi
    ctx = capi.Context(0)  # raises when no CUDA device: GPU tests must not silently fall back
    yield ctx
    ctx.close()
