"""The pixel pass's cut of a map into work items (csrc/smh_px_plan.h — the same header the kernel and the launch code
use): every (8-row group, strip) cell of the map is owned by exactly one item, and every pixel of the OCR quadrant by a
QUAD item.  CPU only."""
import ctypes

import numpy as np
import pytest

import smh_b200  # noqa: F401

from oracle import oracle_py as O
from squad_mortar_helper_b200 import build


@pytest.fixture(scope="module")
def hostcheck():
    L = ctypes.CDLL(str(build.build_hostcheck()))
    u32 = ctypes.c_uint32
    L.smh_hostcheck_px_plan.argtypes = [u32] * 6 + [ctypes.c_void_p] * 3
    L.smh_hostcheck_px_plan.restype = u32
    return L


SIZES = [(160, 120), (200, 150), (320, 240), (640, 480), (1000, 1000), (1280, 720), (1279, 717), (1281, 721), (1366, 768),
         (1600, 900), (1920, 1080), (1920, 1200), (2560, 1080), (2560, 1440), (3440, 1440), (3840, 2160), (5120, 1440),
         (7680, 4320), (16384, 2160)]


def _plan(L, pad, w, h, bqw, bqh, band):
    nstrips = (pad + w + 127) // 128
    groups = (h + 7) // 8
    cover = np.zeros((groups, nstrips), np.uint32)
    quad = np.zeros((groups, nstrips), np.uint8)
    counts = np.zeros(6, np.uint32)
    got = L.smh_hostcheck_px_plan(pad, w, h, bqw, bqh, band, cover.ctypes.data, quad.ctypes.data, counts.ctypes.data)
    assert got == nstrips, (got, nstrips)
    return cover, quad, counts


@pytest.mark.parametrize("band", [16, 64, 128])
@pytest.mark.parametrize("W,H", SIZES)
def test_items_partition_the_map(hostcheck, W, H, band):
    r = O.roi(W, H)
    x, _, w, h = r["map"]
    _, _, bqw, bqh = r["brq"]
    cover, quad, counts = _plan(hostcheck, x & 3, w, h, bqw, bqh, band)
    assert (cover == 1).all(), f"{W}x{H} band {band}: cells owned {np.unique(cover)} times"
    # every pixel of the quadrant (map rows [bqh, 2 bqh), columns [bqw, 2 bqw)) belongs to a QUAD item
    pad = x & 3
    for g in range(bqh // 8, (2 * bqh - 1) // 8 + 1):
        for s in range((bqw + pad) // 128, (2 * bqw - 1 + pad) // 128 + 1):
            assert quad[g, s], (W, H, band, g, s)
    nq, npl, qr0, qr1, qs0, qs1 = (int(v) for v in counts)
    assert qr0 % 32 == 0 and (qr1 % 32 == 0 or qr1 == h) and qr0 <= bqh and qr1 >= min(2 * bqh, h)
    assert nq == ((qr1 - qr0 + band - 1) // band) * (qs1 - qs0) and nq + npl > 0


def test_random_geometries(hostcheck):
    """Not only the reference's ROI arithmetic: any map / quadrant size and pad."""
    rng = np.random.default_rng(5)
    for _ in range(400):
        w, h = int(rng.integers(1, 3000)), int(rng.integers(1, 2500))
        pad = int(rng.integers(0, 4))
        bqw, bqh = w // 2 if rng.random() < 0.8 else int(rng.integers(0, w // 2 + 1)), h // 2 if rng.random() < 0.8 else int(rng.integers(0, h // 2 + 1))
        band = int(rng.choice([8, 16, 32, 64, 128]))
        cover, quad, _ = _plan(hostcheck, pad, w, h, bqw, bqh, band)
        assert (cover == 1).all(), (w, h, pad, bqw, bqh, band)
        if bqw and bqh:
            for g in range(bqh // 8, (2 * bqh - 1) // 8 + 1):
                for s in range((bqw + pad) // 128, (2 * bqw - 1 + pad) // 128 + 1):
                    assert quad[g, s], (w, h, pad, bqw, bqh, band, g, s)
