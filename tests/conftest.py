import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def vo():
    """the CPU oracle (test infrastructure)"""
    import oracle

    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def vb():
    import vers_b200

    vers_b200.lib()  # fails loudly if libvers_b200.so is missing
    return vers_b200


@pytest.fixture(scope="session")
def ctx(vb):
    return vb.Context(0)
