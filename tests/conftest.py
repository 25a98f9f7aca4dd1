import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle

    pyoracle.build()
    return pyoracle.Oracle("brute")


@pytest.fixture(scope="session")
def oracle_nf():
    """Oracle whose kNN is the reference's own vendored nanoflann (oracle/_ref, prebuilt)."""
    from oracle import pyoracle

    pyoracle.build()
    try:
        return pyoracle.Oracle("nanoflann")
    except FileNotFoundError:
        pytest.skip("oracle/_ref/libdynfu_oracle_nf.so not built (reference tree absent)")
