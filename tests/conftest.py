import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_devices() -> int:
    """Number of CUDA devices the library sees; 0 when it is not built or no driver is present."""
    try:
        from rf_inv_b200 import capi
        return max(0, int(capi.load().rfinv_device_count()))
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a GPU: the gpu-marked tests are skipped instead of failing in rfinv_create.  (With
    `-m gpu` on such a box they are skipped too -- loudly: the library has no CPU path to fall back to.)"""
    if not any("gpu" in item.keywords for item in items) or _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device: rf_inv_b200 has no CPU implementation of the path")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
