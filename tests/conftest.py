import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200 box)")
    config.addinivalue_line("markers", "slow: long-running exhaustive checks (SYLT_SLOW=1)")


def _has_gpu():
    """Asks the product itself (cudaGetDeviceCount behind sylt_device_count): no torch needed, and a box whose torch build
    cannot see the GPU does not silently skip the parity tests."""
    try:
        from sylt2d_b200 import lib
        return lib().sylt_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    if os.environ.get("SYLT_REQUIRE_GPU", "0") not in ("", "0"):  # the GPU CI job: a missing device is a failure, not a skip
        raise pytest.UsageError("SYLT_REQUIRE_GPU is set but libsylt2d_b200.so sees no CUDA device")
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
