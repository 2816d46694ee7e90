"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` tests cover the oracle against the golden vectors, the host logic and
the C-ABI surface (no compute calls).  `-m gpu` tests are the parity tests proper: they
call the CUDA path through the C ABI and compare it with the oracle / golden vectors.
Nothing here reads /root/reference (it does not exist on the GPU box).
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(GOLDEN / name)

    return load


@pytest.fixture(scope="session", autouse=True)
def _built_oracle():
    import oracle

    oracle.build()


def rel_max(a, ref):
    """max_i |a_i - ref_i| / |ref_i| with vector norms for (N,3) input."""
    a, ref = np.asarray(a, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    if a.ndim == 2:
        return float((np.linalg.norm(a - ref, axis=1) / np.linalg.norm(ref, axis=1)).max())
    return float((np.abs(a - ref) / np.abs(ref)).max())


def rel_rms(a, ref):
    a, ref = np.asarray(a, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return float(np.sqrt(((a - ref) ** 2).sum() / (ref ** 2).sum()))
