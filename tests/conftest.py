import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import spcbpt_loader
    return spcbpt_loader.load()


@pytest.fixture(scope="session")
def orc():
    import spcbpt_loader
    o = spcbpt_loader.load_oracle()
    o.lib()
    return o


@pytest.fixture(scope="session")
def gpu_ctx(pkg):
    """one context on cuda:0 for the whole GPU session; fails loudly when the extension is missing"""
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    assert os.path.exists(pkg.LIB_PATH), "libspcbpt_b200.so missing: the CUDA path must be the one that runs"
    torch.cuda.init()
    return pkg
