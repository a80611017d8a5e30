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
    from tests import oracle_libs
    return oracle_libs.load_oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference's own sources compiled into oracle/_ref (test infrastructure); skips when it is not built."""
    from tests import oracle_libs
    lib = oracle_libs.load_ref()
    if lib is None:
        pytest.skip("oracle/_ref/libderp_ref.so not built (needs /root/reference)")
    return lib


@pytest.fixture(scope="session")
def cuda():
    """The product library; GPU tests only. No fallback: missing library = failure."""
    from facebook360_dep_b200 import capi
    return capi.load_cuda()
