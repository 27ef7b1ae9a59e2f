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
    """The CPU checkers (test infrastructure).  Builds liboracle_emitcpu.so on first use."""
    from oracle import oracle as O
    O.emit_lib()
    return O


@pytest.fixture(scope="session")
def ref(oracle):
    """The compiled reference DSL (oracle/_ref).  Present where /root/reference exists at build
    time, or prebuilt (it travels with the repo snapshot to the GPU box)."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here and no prebuilt library)")
    oracle.ref_lib()
    return oracle


@pytest.fixture(scope="session")
def hb():
    """The product: ctypes front of libhipacc_b200.so.  Fails loudly if the CUDA library is missing."""
    import hipacc_b200
    hipacc_b200.lib()
    return hipacc_b200
