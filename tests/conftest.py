import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/libcn_oracle.so) wrapped for numpy.  Test infrastructure only."""
    import _oracle
    return _oracle.Oracle()


@pytest.fixture(scope="session")
def cn():
    """The product: the C-ABI library through its Python binding.  Fails loudly if not built."""
    import cute_nucleotides_b200 as pkg
    from cute_nucleotides_b200 import _lib
    _lib.load()
    return pkg
