import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    import helpers
    return helpers.load_port()


@pytest.fixture(scope="session")
def ref():
    import helpers
    lib = helpers.load_ref()
    if lib is None:
        pytest.skip("oracle/_ref/liboracle_ref.so not built (needs /root/reference)")
    return lib


@pytest.fixture(scope="session")
def sim():
    import helpers
    return helpers.load_sim()


@pytest.fixture(scope="session")
def gpu_ctx():
    import crunch2_b200 as crn
    ctx = crn.Context(0)      # raises loudly if the nvcc library or the device is missing
    yield ctx
    ctx.close()
