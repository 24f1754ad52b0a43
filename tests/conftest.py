import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build_oracle()
    return oracle


# Property tests draw the same examples on every run (a red driver run must be reproducible here);
# QSMRT_HYPOTHESIS_RANDOM=1 explores fresh examples instead.
HYPOTHESIS_DERANDOMIZE = os.environ.get("QSMRT_HYPOTHESIS_RANDOM", "0") != "1"
