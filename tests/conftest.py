import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Compile the C-ABI library + the CPU logic emulator once per session."""
    import __graft_entry__ as g

    g.build()
    return g


@pytest.fixture()
def cuda(built):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    old = torch.get_default_device()
    torch.set_default_device("cuda:0")
    yield torch.device("cuda:0")
    torch.set_default_device(old)
