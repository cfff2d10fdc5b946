import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", params=["auto", "general", "lanes", "tiles"])
def ctx(request):
    """One fgpu context per search-kernel family for the whole GPU session; fails loudly when the library or
    the device is missing.  "auto" is the product's own choice (the warp-cooperative kernels wherever the grid
    is regular, with the NeighborList mapping picked from the expected bonds per query), "general" forces the
    thread-per-query family that covers the remaining cases; "tiles" and "lanes" force the two mappings of the
    warp-cooperative NeighborList search (lanes over the candidates of a home tile / one query per lane): all must
    agree with the oracle bit for bit."""
    from freud_b200 import _capi

    c = _capi.Context(0)
    c.force_general_search(request.param == "general")
    if request.param == "lanes":
        c.set_tuning("lanes_over_queries", 1)
    if request.param == "tiles":
        c.set_tuning("lanes_over_queries", 0)
    yield c
    c.close()
