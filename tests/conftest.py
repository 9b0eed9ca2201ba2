import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def nrs():
    """The product package (hyphenated directory name -> importlib)."""
    return importlib.import_module("navtech-radar-slam_b200")


@pytest.fixture(scope="session")
def synth(nrs):
    return importlib.import_module("navtech-radar-slam_b200.synth")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def ctx(nrs):
    """One CUDA context for the whole GPU session; skips loudly-failing is NOT allowed:
    -m gpu tests error out if the library or the device is missing."""
    c = nrs.Context(0)
    yield c
    c.close()
