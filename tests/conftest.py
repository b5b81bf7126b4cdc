import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# Parity statistics the GPU tests want on the record (sel mismatches and their fp64 adjudication, L1 kinks, outlier
# counts): collected here and printed in the terminal summary, so they show up under -q as well.
PARITY_STATS = []


def record_parity(what, **kw):
    PARITY_STATS.append((what, kw))


def pytest_terminal_summary(terminalreporter):
    if not PARITY_STATS:
        return
    terminalreporter.section("parity statistics (CUDA path vs oracle)")
    for what, kw in PARITY_STATS:
        terminalreporter.write_line(what + ": " + ", ".join(f"{k}={v}" for k, v in kw.items()))
