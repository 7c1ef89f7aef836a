import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "tests")):
    if p_ not in sys.path:
        sys.path.insert(0, p_)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def kw():
    """The product package (ctypes binding of libkwave_b200.so)."""
    return importlib.import_module("k-wave-fluid-cuda_b200")


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module("k-wave-fluid-cuda_b200.synth")
