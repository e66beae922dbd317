import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def ctx():
    """One device context for the whole GPU session; fails loudly without the CUDA library."""
    import starch3_b200 as s3
    c = s3.Context(0)
    yield c
    c.close()


def golden(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()
