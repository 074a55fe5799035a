"""Soundness of k_lsd's exact ray filter (csrc/smh_lsd_filter.h), checked on the CPU: for seeds of
synthetic and hand-made masks, every ray whose squared length exceeds lsd.rs:95's 2500 must be in
the set the filter keeps (so skipping the others cannot change find_lines' output)."""
import ctypes

import numpy as np
import pytest

import smh_b200  # noqa: F401
from helpers import oracle_run
from oracle import oracle_py as O
from squad_mortar_helper_b200 import build, synth


@pytest.fixture(scope="module")
def hostcheck():
    L = ctypes.CDLL(str(build.build_hostcheck()))
    L.smh_hostcheck_lsd_filter.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float,
                                           ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
    L.smh_hostcheck_lsd_filter_quarters.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float,
                                                    ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
    return L


def marked_rays(L, img, x, y, gap):
    """Rays kept by the per-pixel form AND by the quarter form k_lsd2 runs (every bitmap phase 0..7 of the
    8-pixel quarters): both must be supersets of the rays that matter, the quarter form of the per-pixel one."""
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros(3600, np.uint8)
    nw = L.smh_hostcheck_lsd_filter(img.ctypes.data, img.shape[1], img.shape[0], x, y, gap, out.ctypes.data)
    keep = out.astype(bool)
    for pad in (0, 3, 5):
        q = np.zeros(3600, np.uint8)
        L.smh_hostcheck_lsd_filter_quarters(img.ctypes.data, img.shape[1], img.shape[0], pad, x, y, gap, q.ctypes.data)
        q = q.astype(bool)
        assert not (keep & ~q).any(), f"quarter marking (pad {pad}) drops rays the per-pixel marking keeps at seed {(x, y)} gap {gap}"
        keep_q = q
    return nw, keep & keep_q


def check_seeds(L, img, seeds, gap):
    kept_total = 0
    for (sx, sy) in seeds:
        cx, cy = O.get_centre_in(img, float(sx), float(sy))
        lens = O.ray_lengths_in(img, cx, cy, float(gap))
        nw, keep = marked_rays(L, img, cx, cy, gap)
        long_rays = lens > 2500.0
        missed = long_rays & ~keep
        assert not missed.any(), (f"seed {(sx, sy)} -> centre {(cx, cy)} gap {gap}: filter drops rays "
                                  f"{np.nonzero(missed)[0][:8].tolist()} with len^2 {lens[missed][:8].tolist()}")
        kept_total += int(keep.sum())
    return kept_total


def pick_seeds(img, rng, n):
    ys, xs = np.nonzero(img == 255)
    if len(xs) == 0:
        return []
    idx = rng.choice(len(xs), size=min(n, len(xs)), replace=False)
    return [(int(xs[i]), int(ys[i])) for i in idx]


@pytest.mark.parametrize("W,H,kw,gap", [
    (1280, 720, dict(n_segments=6, n_blips=20, n_labels=2), 15),
    (1920, 1080, synth.DENSE, 15),
    (1920, 1080, synth.DENSE, 22),
    (1280, 720, dict(n_segments=40, n_blips=0, n_labels=1), 15),
    (1280, 720, dict(n_segments=6, n_blips=20, n_labels=2), 5),
    (1280, 720, dict(n_segments=6, n_blips=20, n_labels=2), 30),
])
def test_filter_keeps_every_long_ray_synthetic(hostcheck, W, H, kw, gap):
    O.set_threads(8)
    rng = np.random.default_rng(7)
    frame = synth.frames_host(synth.SynthParams(W, H, **kw), 5, 1)[0]
    img = oracle_run(frame, True, 0, gap)["lsd_image"]
    seeds = pick_seeds(img, rng, 120)
    kept = check_seeds(hostcheck, img, seeds, gap)
    assert kept < 0.5 * 3600 * len(seeds), "the filter should prune most rays on these scenes"


def test_filter_border_and_hand_made_shapes(hostcheck):
    """Lines hugging / leaving through the left and top borders (the D7 exit quirk), crossing
    lines, a disc, dashed lines whose dashes are exactly max_gap apart."""
    O.set_threads(8)
    h, w = 300, 400
    img = np.zeros((h, w), np.uint8)
    img[3:6, 0:200] = 255          # along the top border
    img[0:250, 2:5] = 255          # along the left border
    for t in range(0, 260):        # diagonal into the top-left corner
        img[min(h - 1, t):min(h, t + 3), min(w - 1, t):min(w, t + 3)] = 255
    img[150:153, 100:399] = 255    # to the right border
    yy, xx = np.mgrid[0:h, 0:w]
    img[(yy - 220) ** 2 + (xx - 300) ** 2 <= 15 ** 2] = 255
    for x in range(20, 380, 19):   # dashes: 4 white, 15 black
        img[280:283, x:x + 4] = 255
    rng = np.random.default_rng(3)
    seeds = pick_seeds(img, rng, 250) + [(0, 4), (3, 0), (199, 4), (3, 249), (1, 1), (398, 151)]
    for gap in (15, 22, 3, 4, 32, 1, 2):
        check_seeds(hostcheck, img, seeds, gap)


def test_filter_random_noise(hostcheck):
    """Dense random speckle: many short accidental chains."""
    O.set_threads(8)
    rng = np.random.default_rng(11)
    img = (rng.random((240, 320)) < 0.06).astype(np.uint8) * 255
    seeds = pick_seeds(img, rng, 80)
    check_seeds(hostcheck, img, seeds, 15)


@pytest.mark.parametrize("NW", [4, 8])
def test_ray_bitmap_words_to_warps_is_a_bijection(NW):
    """k_lsd2 (csrc/kernels_lsd.cuh): lane l of warp p takes word `wmine` of the 113-word ray bitmap — the first 56 words by
    p + NW l, the others shifted by half a CTA so that the two opposite directions of a line (56 words apart) fall to
    different warps.  Restated here: every word is taken exactly once, and words w, w + 56 never share a warp."""
    WORDS, HALF = 113, 56
    owner = {}
    for p in range(NW):
        for lane in range(32):
            if lane < 16:
                w = p + NW * lane
                w = w if w < HALF else WORDS
            else:
                w = HALF + (p + NW - NW // 2) % NW + NW * (lane - 16)
            if w < WORDS:
                assert w not in owner, (w, owner[w], p)
                owner[w] = p
    assert sorted(owner) == list(range(WORDS))
    for w in range(HALF):
        assert owner[w] != owner[w + HALF], w
        assert owner[w] != owner[(w + 1) % HALF] or NW == 1
