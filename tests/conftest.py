"""pytest configuration: path setup, the `gpu` marker and shared synthetic-data helpers."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "sift-based-od_b200"
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Make sure libsod_b200.so exists (nvcc cross-compiles without a GPU)."""
    from sod_b200.build import build
    build()


def sift_like(rng: np.random.Generator, n: int) -> np.ndarray:
    """u8 descriptors with SIFT statistics: |N(0,1)|, L2-normalise, clip 0.2, renormalise, x512."""
    x = np.abs(rng.standard_normal((n, 128)))
    x /= np.linalg.norm(x, axis=1, keepdims=True) + 1e-12
    x = np.minimum(x, 0.2)
    x /= np.linalg.norm(x, axis=1, keepdims=True) + 1e-12
    return np.clip(np.rint(x * 512.0), 0, 255).astype(np.uint8)
