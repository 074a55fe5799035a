import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

REFERENCE = Path("/root/reference")
SAMPLES_REF = REFERENCE / "vision-common" / "samples"
SAMPLES_LOCAL = ROOT / "tests" / "_samples"   # git-ignored copies that travel to the GPU box
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the checker (C oracle) and the product library once per session."""
    import __graft_entry__ as g
    g.build()


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def sample_path(name: str):
    for d in (SAMPLES_LOCAL, SAMPLES_REF):
        p = d / name
        if p.exists():
            return p
    return None


def load_png_bgra(path):
    import cv2
    import numpy as np
    img = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
    assert img is not None, path
    if img.ndim == 2:
        img = cv2.cvtColor(img, cv2.COLOR_GRAY2BGRA)
    elif img.shape[2] == 3:
        img = np.dstack([img, np.full(img.shape[:2], 255, np.uint8)])
    return np.ascontiguousarray(img)
