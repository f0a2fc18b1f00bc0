import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def dr3():
    """The product package (its name starts with a digit)."""
    return importlib.import_module("3dr_b200")


@pytest.fixture(scope="session")
def ctx(dr3):
    """A dr3lk context on cuda:0 -- gpu tests only. Fails loudly (no CPU fallback) when the library or device is missing."""
    c = dr3.Context(0)
    yield c
    c.close()
