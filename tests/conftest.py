import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (C++ restatement of innr 0.6.3). Checker only."""
    from oracle import innr_oracle
    innr_oracle.build()
    return innr_oracle


def _api_params():
    return [pytest.param("oracle", id="oracle"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=_api_params())
def api(request):
    """The reference-shaped API under test: the oracle (CPU, pins the oracle against the reference's own
    tests) or the CUDA product (innr_b200, through the C-ABI)."""
    if request.param == "oracle":
        from oracle import innr_oracle
        innr_oracle.build()
        return innr_oracle
    import innr_b200
    return innr_b200
