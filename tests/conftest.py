import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def _cuda_available():
    try:
        import monty_b200
        return monty_b200._lib.lib.mb200_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def port():
    from oracle_lib import Oracle, build_oracles
    build_oracles()
    return Oracle("port")


@pytest.fixture(scope="session")
def reference():
    from oracle_lib import Oracle, have_reference
    if not have_reference():
        pytest.skip("oracle/_ref/libmonty_ref.so not built (needs /root/reference)")
    return Oracle("reference")


@pytest.fixture(scope="session")
def oracle(port):
    """The best available checker: the compiled reference if present, else the
    C restatement (which test_oracle.py pins against the reference's vectors)."""
    from oracle_lib import Oracle, have_reference
    return Oracle("reference") if have_reference() else port


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(HERE, "golden", "reference_vectors.npz"))
