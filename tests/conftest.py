import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import enumcpu
    enumcpu.lib()
    return enumcpu


@pytest.fixture(scope="session")
def gpu_lib():
    import simplexmethod_b200 as sm
    L = sm.lib()
    if L.enumgpu_device_count() < 1:
        pytest.fail("GPU test selected but libenumgpu sees no CUDA device (no CPU fallback exists)")
    return L
