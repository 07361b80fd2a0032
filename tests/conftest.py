import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")


def free_port():
    """A TCP port that is free right now on 127.0.0.1 (bind to 0, read it back): the multi-process tests used
    pid-derived ports, some of them inside the kernel's ephemeral range (32768+), where a connection of an earlier
    test could already sit -- one rendezvous in three failed with 'address already in use'."""
    import socket

    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_ok():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_ok():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure both native libraries exist (nvcc cross-compiles without a GPU)."""
    from gaussian_splatting_3d_b200 import _build
    from oracle import gs_oracle

    _build.build()
    gs_oracle.build()
