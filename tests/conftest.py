import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracles():
    """Build the CPU oracles once (the reference one only where /root/reference exists; prebuilt files are kept)."""
    from oracle import pyoracle

    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine; fails loudly (no fallback) when the extension is missing or no GPU is present."""
    from geneticscre_b200 import api, build

    build.build()
    return api
