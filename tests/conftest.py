import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    """Safety net for the GPU box: a GPU test that hangs (a kernel that never returns blocks inside C code, where a signal
    cannot interrupt it) is ended by pytest-timeout's thread method instead of holding the box until the driver's limit.
    The validated GPU suite takes about a minute in total; 900 s per test is far above any legitimate run."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("gpu") is not None and item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(900, method="thread"))


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (test infrastructure; builds oracle/liboracle.so on first use)."""
    from oracle import oracle as orc
    orc.load()
    # the oracle forks an OpenMP team per product / solve; on a many-core GPU box the fork-join cost of a 100+ thread team
    # dominates the small test problems (measured: the QR parity tests took 5 minutes instead of seconds)
    orc.set_num_threads(max(1, min(os.cpu_count() or 1, 16)))
    return orc


@pytest.fixture(scope="session")
def fb():
    """The product: ctypes binding + host mirror over libfaer_b200.so (no CPU fallback)."""
    import faer_b200
    faer_b200.load()
    return faer_b200


@pytest.fixture(scope="session")
def cuda_dev(fb):
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    assert fb.load().faer_b200_device_count() > 0
    return torch.device("cuda:0")
