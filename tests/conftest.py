import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.lib()
    return O


@pytest.fixture(scope="session")
def ctx():
    import voidin_b200 as vb

    try:
        return vb.Context(0)
    except (vb.BvhCudaError, OSError, ImportError) as e:  # CPU-only box: the gpu-marked tests are skipped, not errors
        pytest.skip(f"no CUDA device / libbvh_cuda.so unavailable: {e}")
