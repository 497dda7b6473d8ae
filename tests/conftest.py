import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by `pytest -m gpu` on the GPU box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """Build the oracle, the host check twin and the product library if they are missing.
    (On the GPU box the prebuilt in-tree .so files travel with the snapshot.)"""
    need = [
        os.path.join(ROOT, "oracle", "liboracle.so"),
        os.path.join(ROOT, "tests", "native", "libhostcheck.so"),
        os.path.join(ROOT, "doppler_b200", "libdoppler_b200.so"),
        os.path.join(ROOT, "doppler_b200", "bin", "doppler"),
    ]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()
    yield


@pytest.fixture(scope="session")
def oracle():
    from tests import oracle_lib
    return oracle_lib.Oracle()


@pytest.fixture(scope="session", params=["default", "bulk-async kernels only", "bulk-async, 4-warp segmented pipelines"])
def mixer(request):
    """The suite runs twice: with the product's thresholds (short inputs take the latency-shaped small kernel and the
    zero-copy host path) and with both disabled, so that every input -- however short or ragged -- also goes through
    the persistent bulk-async kernels and the staged host pipeline."""
    import doppler_b200
    m = doppler_b200.Mixer(0)
    if request.param != "default":
        m.tune(small_max_samples=0, tiny_host_bytes=0, seg_variant=int("4-warp" in request.param))
    yield m
    m.close()
