import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _gpu_available():
    try:
        from zippy_b200 import _native
        return _native.lib().zb200_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a box without a CUDA device or without the library."""
    if not any("gpu" in it.keywords for it in items) or _gpu_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device and libzippy_b200.so (run on the GPU box)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    from tests import util
    return util.load_golden()


@pytest.fixture(scope="session")
def corpus():
    from tests import util
    return util.load_corpus()
