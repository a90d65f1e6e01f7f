import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def O():
    """The CPU parity oracle (test infrastructure)."""
    from oracle import oracle

    return oracle


@pytest.fixture(scope="session")
def gs():
    import gridsplines_b200

    return gridsplines_b200


@pytest.fixture(scope="session")
def ctx(gs):
    return gs.default_context()
