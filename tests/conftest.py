import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the product library and the checkers once per session (no-op when up to date; nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g

    g.build()


@pytest.fixture(scope="session")
def product():
    from dxmclib_b200 import scene

    return scene.product_lib()


@pytest.fixture(scope="session")
def reference():
    from dxmclib_b200 import scene

    if not os.path.exists(scene.REFERENCE_LIB):
        pytest.skip("oracle/_ref/libdxmc_ref.so not built (needs /root/reference at build time)")
    return scene.reference_lib()


@pytest.fixture(scope="session")
def gpu():
    from dxmclib_b200 import cabi

    if cabi.device_count() < 1:
        pytest.fail("this test is marked gpu but no CUDA device is visible; the product has no CPU fallback")
    return 0
