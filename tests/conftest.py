import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device AND the built sm_100a library; on a box without either they are skipped, not failed
    (the product itself still fails loudly without the library: nrc_hpm_renderer_b200._lib raises)."""
    reason = None
    try:
        import torch
        if not torch.cuda.is_available():
            reason = "no CUDA device"
    except Exception as e:      # pragma: no cover
        reason = f"torch unavailable: {e}"
    if reason is None and not os.path.exists(os.path.join(ROOT, "nrc_hpm_renderer_b200", "libnrchpm_b200.so")):
        reason = "libnrchpm_b200.so not built (python __graft_entry__.py)"
    if reason:
        skip = pytest.mark.skip(reason=reason)
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle
    oracle.build()
    return oracle


def golden(name):
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", name))
