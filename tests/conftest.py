import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def port():
    from oracle.oracle import Checker
    return Checker("port")


@pytest.fixture(scope="session")
def reference():
    from oracle.oracle import Checker, have_reference
    if not have_reference():
        pytest.skip("oracle/_ref/libloosref.so not built (needs /root/reference)")
    return Checker("reference")


@pytest.fixture(scope="session")
def checker():
    """The strongest CPU checker available: the reference's own arithmetic when
    oracle/_ref was built, else our C restatement."""
    from oracle.oracle import Checker, have_reference
    return Checker("reference" if have_reference() else "port")
