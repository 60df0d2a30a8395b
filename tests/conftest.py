import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "needs_reference: needs /root/reference mounted (build container only)")


def pytest_collection_modifyitems(config, items):
    import ref_shims
    if not ref_shims.reference_available():
        skip = pytest.mark.skip(reason="/root/reference not mounted")
        for it in items:
            if "needs_reference" in it.keywords:
                it.add_marker(skip)
