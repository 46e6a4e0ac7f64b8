import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


_gpu_memory = []


@pytest.fixture(autouse=True, scope="module")
def _release_gpu_memory_between_modules(request):
    """GPU test modules build estimators at several image sizes, each with per-shape plans and buffers (tens of GB at 640x512 with
    32-frame chunks): drop them when the module is done and report what stayed allocated."""
    yield
    torch = sys.modules.get("torch")
    if torch is None or not torch.cuda.is_available():
        return
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    _gpu_memory.append((request.module.__name__, torch.cuda.max_memory_allocated() / 2 ** 30, torch.cuda.memory_allocated() / 2 ** 30))
    torch.cuda.reset_peak_memory_stats()


def pytest_terminal_summary(terminalreporter):
    for name, peak, held in _gpu_memory:
        terminalreporter.write_line(f"gpu memory, {name}: peak {peak:.1f} GiB allocated, {held:.1f} GiB still held after the module")
