import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ora():
    """The CPU oracle (test infrastructure): builds oracle/libpfx_oracle.so on first use."""
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def pf():
    """The product package; building the CUDA library if it is not there yet."""
    lib = os.path.join(ROOT, "plateflex_b200", "libplateflex_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()
    import plateflex_b200
    return plateflex_b200
