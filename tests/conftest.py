import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "needs_reference: needs the reference tree at /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    import refharness
    if not refharness.reference_available():
        skip = pytest.mark.skip(reason="reference tree not present")
        for it in items:
            if "needs_reference" in it.keywords:
                it.add_marker(skip)


@pytest.fixture(scope="session")
def tables():
    from ranslice_b200.tables import load_tables
    return load_tables()


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    return load
