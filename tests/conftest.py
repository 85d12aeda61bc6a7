import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import hbo
    hbo.lib()
    return hbo


@pytest.fixture(scope="session")
def gpu_ctx():
    import hunter_bipedal_control_b200 as hb
    ctx = hb.Context(horizon_N=100, dt=0.01, max_batch=1024, device=0)
    yield ctx
    ctx.close()
