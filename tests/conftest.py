import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure).  Builds its C range coder on first use."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import crdr_oracle
    from compressai import ans
    ans.build_lib()
    return crdr_oracle


@pytest.fixture(scope="session")
def crdr_opt():
    from crdr_b200.config import BaseConfig
    return BaseConfig.fromfile(os.path.join(ROOT, "config", "crdr.yaml"), device="cuda:0", is_train=False)
