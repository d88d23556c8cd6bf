import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Build (or find prebuilt) native pieces once per session."""
    import __graft_entry__ as g
    g.build_checkers()
    return g


@pytest.fixture(scope="session")
def oracle(built):
    from oracle import bindings as checkers
    return checkers.Oracle()


@pytest.fixture(scope="session")
def reference(built):
    """The reference's own sources (oracle/_ref); only where it was built."""
    from oracle import bindings as checkers
    path = os.path.join(ROOT, "oracle", "_ref", "libremy_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libremy_ref.so not built (needs /root/reference)")
    return checkers.Reference(path)


@pytest.fixture(scope="session")
def device(built):
    """The product: the nvcc-built library on cuda:0.  No fallback: a missing library or
    device fails the test."""
    from remy_b200 import abi
    built.build_cuda()
    return abi.Device(0)
