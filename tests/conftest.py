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
    """The plain-C restatement (oracle/ms_oracle.c), built on demand — the checker of every parity test."""
    from oracle import pyoracle
    if not os.path.exists(pyoracle.ORACLE_SO):
        pyoracle.build(ref=False)
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def reference():
    """The reference's own code (oracle/_ref/libmsref.so). Present wherever /root/reference was at build time."""
    from oracle import pyoracle
    if not os.path.exists(pyoracle.REF_SO):
        if os.path.isdir("/root/reference/libms"):
            pyoracle.build(ref=True)
        else:
            pytest.skip("oracle/_ref/libmsref.so not built and /root/reference absent")
    return pyoracle.Reference()


@pytest.fixture(scope="session")
def builder():
    """One msgpu context on cuda:0 for the whole session. Fails loudly (no CPU fallback) when the library or the GPU
    is missing."""
    from muchsalsa_b200.construct import GraphBuilder
    b = GraphBuilder(0)
    yield b
    b.close()
