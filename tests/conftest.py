import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", params=["rank", "stream"])
def engine(request):
    """One handle on cuda:0 per counting path for the whole GPU session; creating it fails loudly
    without a GPU.  "rank" = default dispatch (rank path where it applies), "stream" = every
    case through the streaming kernel."""
    from nullranges_b200 import Engine
    e = Engine(0, force_stream=request.param == "stream")
    yield e
    e.close()
