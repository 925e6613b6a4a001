import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    """the plain-C restatement of the reference path (test infrastructure)"""
    from oracle.pyoracle import Oracle, build, PORT_SO
    if not os.path.exists(PORT_SO):
        build(want_ref=False)
    return Oracle("port")


@pytest.fixture(scope="session")
def ref():
    """the reference's own code compiled from /root/reference (skips where it was never built)"""
    from oracle.pyoracle import Oracle, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref/libhrm_ref.so not built (needs /root/reference)")
    return Oracle("ref")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible -- there is no CPU fallback to test")
    import hashreadmapper_b200.api as api
    return api
