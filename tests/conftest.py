import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O  # oracle/oracle.py: test infrastructure
    O.build()
    return O


@pytest.fixture(scope="session")
def reference():
    """oracle/ref.py: the reference's own function bodies (extracted at build time) behind ctypes; test infrastructure.
    Built where /root/reference exists; elsewhere the prebuilt oracle/_ref/libsrukf_ref.so is used, else skip."""
    import ref as R
    if not R.available():
        pytest.skip("oracle/_ref/libsrukf_ref.so absent and no reference tree to build it from")
    R.lib()
    return R


@pytest.fixture(scope="session")
def built_lib():
    """libsrukf_b200.so, built in-tree (nvcc cross-compiles without a GPU)."""
    from cv_monoslam_b200 import build, capi
    build.build()
    return capi.load_library()


def relmax(a, b):
    import numpy as np
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))
