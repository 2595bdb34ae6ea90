import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the libraries and the oracle once per session if they are missing (make is incremental)."""
    from ntrtsim_b200 import _capi as capi
    if not (os.path.exists(capi.lib_path()) and os.path.exists(capi.host_lib_path())
            and os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so"))):
        from ntrtsim_b200.build import build_all
        build_all()
    yield
