import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_NAME = "3d_adapt_auto_driving_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load(sub=None):
    return importlib.import_module(PKG_NAME + ("." + sub if sub else ""))


@pytest.fixture(scope="session")
def pkg():
    return load()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.lib()
    return orc


@pytest.fixture(scope="session")
def legacy():
    """The reference's own CUDA kernels (oracle/_ref/libpn2_legacy.so); skip when absent."""
    from oracle import legacy as leg
    if not leg.available():
        pytest.skip("oracle/_ref/libpn2_legacy.so not available")
    leg.lib()
    return leg


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    load("cabi").lib()  # fail loudly if the extension is not built
    return torch.device("cuda:0")
