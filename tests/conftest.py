import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def _gpu_ready():
    try:
        import torch
        if not torch.cuda.is_available():
            return False, "no CUDA device"
    except Exception as exc:   # noqa: BLE001
        return False, "torch unavailable: %s" % exc
    return True, ""      # (a GPU box without the built .so must FAIL these tests, not skip them)


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device: skip them (with the reason) on a CPU box instead of failing 30 times.
    On a box WITH a GPU they always run - a missing libqpth_b200.so is then a loud failure, never a skip."""
    ok, why = _gpu_ready()
    if ok:
        return
    skip = pytest.mark.skip(reason="gpu test: " + why)
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
