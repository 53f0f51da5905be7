import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with open(os.path.join(ROOT, "tests", "golden", f"{name}.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_pure():
    return load_golden("pure")


@pytest.fixture(scope="session")
def golden_mix():
    return load_golden("mix")


@pytest.fixture(scope="session")
def golden_gc():
    return load_golden("gc")


@pytest.fixture(scope="session")
def golden_state():
    """Gradients of the reference's `derivatives` (made by oracle/make_golden.py state_grad)."""
    return load_golden("state_grad")
