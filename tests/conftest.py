import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def checker():
    """The strongest CPU checker available: the reference's own object code, else the C port."""
    import oracle
    oracle.build()
    return oracle.best()


@pytest.fixture(scope="session")
def port():
    import oracle
    oracle.build()
    return oracle.port()


@pytest.fixture(autouse=True)
def fresh_capacity_estimates(request):
    """Every GPU test starts without learned pair-capacity estimates: they are keyed by (B, N, stride, voxel) only, so
    a test that reuses a shape with denser clouds would otherwise trip the deferred overflow check of NeighborPlan
    (documented behaviour: poison + Conv3pError + raised estimate) depending on the order the tests run in."""
    if "gpu" in request.keywords:
        from pointwise_b200 import ops
        ops._capacity_hint.clear()
        ops._generic_capacity_hint.clear()
    yield
