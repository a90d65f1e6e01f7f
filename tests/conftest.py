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


def pytest_sessionfinish(session, exitstatus):
    """Measured errors of every tolerance comparison (helpers.check_tol) -> gpurun_out/parity_tests.json."""
    import json

    from helpers import PARITY_LOG

    out = os.path.join(ROOT, "gpurun_out")
    if PARITY_LOG and os.path.isdir(out):
        with open(os.path.join(out, "parity_tests.json"), "w") as f:
            json.dump(PARITY_LOG, f, indent=1)
