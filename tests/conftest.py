import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure). Built on demand from oracle/Makefile."""
    from facebook360_dep_b200 import capi
    if not os.path.exists(capi.ORACLE_LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return capi.load_oracle()


@pytest.fixture(scope="session")
def cuda():
    """The product library; GPU tests only. No fallback: missing library = failure."""
    from facebook360_dep_b200 import capi
    return capi.load_cuda()
