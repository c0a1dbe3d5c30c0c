import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def product_lib():
    from treeqp_b200 import capi
    return capi.TreeQpLib(capi.PRODUCT_LIB)


@pytest.fixture(scope="session")
def ref_lib():
    """the unmodified reference compiled by oracle/Makefile (travels to the GPU box as a .so)"""
    from treeqp_b200 import capi
    import oracle_port
    if not os.path.exists(oracle_port.REF_LIB):
        pytest.skip("oracle/_ref/libtreeqp_ref.so not built (needs /root/reference)")
    return capi.TreeQpLib(oracle_port.REF_LIB)
