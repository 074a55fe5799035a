"""CPU tests of the rows either side of the trait: the C oracle's restatements of
calc_meters_to_px_ratio (src/vision/mpx_ratio.rs) and find_minimap (src/vision/find_minimap.rs) against
hand-derived known answers and the independent numpy restatement."""
import numpy as np
import pytest

import smh_b200  # noqa: F401
from oracle import np_oracle as NP
from oracle import oracle_py as O
from squad_mortar_helper_b200 import synth
from conftest import sample_path, load_png_bgra


def bar_image(w=200, h=120, bars=()):
    """White image with black scale bars: (x0, x1, y, tick) = horizontal run [x0, x1] on row y and end ticks
    going DOWN `tick` pixels (the reference's test looks at rows y .. y+3)."""
    img = np.full((h, w), 255, np.uint8)
    for x0, x1, y, tick in bars:
        img[y, x0:x1 + 1] = 0
        img[y:y + tick, x0] = 0
        img[y:y + tick, x1] = 0
    return img


def test_scale_width_known_answers():
    img = bar_image(bars=[(20, 80, 30, 5)])
    # label bottom at y=25 above the bar's middle: the walk meets the bar on row 30, the ticks bound it
    assert O.calc_meters_to_px_ratio(img, [(100, (50, 25))]) == 100.0 / 58.0   # right = 79, left = 21
    r, dbg = O.calc_meters_to_px_ratio(img, [(100, (50, 25))], with_debug=True)
    assert dbg == [(100, 21, 79, 30)]
    assert O.calc_meters_to_px_ratio(img, [(100, (50, 3))]) is None            # y < MIN_SCALE_VERTICAL_BAR_HEIGHT
    assert O.calc_meters_to_px_ratio(img, [(100, (50, 5))]) is None            # bar beyond max_scale_y_offset (6 rows)
    assert O.calc_meters_to_px_ratio(img, []) is None
    # ticks of 3 px fail the 4-pixel test -> no vertical bar -> None
    assert O.calc_meters_to_px_ratio(bar_image(bars=[(20, 80, 30, 3)]), [(100, (50, 25))]) is None
    # narrower than MIN_SCALE_WIDTH
    assert O.calc_meters_to_px_ratio(bar_image(bars=[(40, 50, 30, 5)]), [(100, (45, 26))]) is None
    # mean of the scales that are found, in order (the "rayon ladder", mpx_ratio.rs:98-124)
    img = bar_image(bars=[(20, 80, 30, 5), (100, 171, 60, 6), (10, 40, 90, 4)])
    a, b, c = 100.0 / 58.0, 300.0 / 69.0, 50.0 / 28.0
    assert O.calc_meters_to_px_ratio(img, [(100, (50, 28)), (300, (120, 57)), (50, (25, 88))]) == (a + b + c) / 3.0
    assert O.calc_meters_to_px_ratio(img, [(100, (50, 28)), (7, (190, 57)), (50, (25, 88))]) == (a + c) / 2.0
    assert O.calc_meters_to_px_ratio(img, [(7, (190, 57)), (300, (120, 57))]) == b


def test_scale_width_out_of_image_reads_are_non_black():
    """D16: ticks that run off the bottom of the image do not count as a vertical bar."""
    img = bar_image(w=100, h=40, bars=[(20, 80, 38, 2)])
    assert O.calc_meters_to_px_ratio(img, [(100, (50, 36))]) is None
    assert O.calc_meters_to_px_ratio(img, [(100, (150, 36))]) is None  # x outside the image


@pytest.mark.parametrize("seed", range(6))
def test_scale_width_matches_numpy_restatement_on_random_images(seed):
    rng = np.random.default_rng(seed)
    w, h = int(rng.integers(60, 300)), int(rng.integers(40, 160))
    img = np.where(rng.random((h, w)) < 0.35, 0, 255).astype(np.uint8)
    # plant bars so that some probes succeed
    for _ in range(12):
        x0 = int(rng.integers(1, w - 20)); x1 = min(w - 2, x0 + int(rng.integers(8, 80))); y = int(rng.integers(5, h - 6))
        img[y, x0:x1 + 1] = 0; img[y:y + 5, x0] = 0; img[y:y + 5, x1] = 0
        img[y:y + 4, x0 + 1:x1] |= np.where(rng.random((min(4, h - y), x1 - x0 - 1)) < 0.9, 255, 0).astype(np.uint8)
        img[y, x0:x1 + 1] = 0
    n_some = 0
    for _ in range(300):
        sc = [(int(rng.integers(1, 900)), (int(rng.integers(0, w + 3)), int(rng.integers(0, h + 3)))) for _ in range(int(rng.integers(1, 4)))]
        got, want = O.calc_meters_to_px_ratio(img, sc), NP.calc_meters_to_px_ratio(img, sc)
        assert got == want, (sc, got, want)
        n_some += got is not None
    assert n_some > 10


def test_edginess_threshold_is_an_integer_bound():
    """`max as f32 / 765.0 <= 0.01` (find_minimap.rs:45,66) <=> max <= SMH_MINIMAP_EDGINESS_MAX_SUM = 7."""
    s = np.arange(766, dtype=np.float32)
    assert np.array_equal((s / np.float32(765.0)) <= np.float32(0.01), np.arange(766) <= 7)


def _views():
    out = []
    for name in ("point_intersect.png", "full_1280x1024.png", "full_1024x768.png", "difficult.png"):
        p = sample_path(name)
        if p is not None:
            f = load_png_bgra(p)
            out.append((name, f, O.roi(f.shape[1], f.shape[0])["map"]))
    for k, (W, H) in enumerate(((1280, 720), (1920, 1080))):
        f = synth.frames_host(synth.SynthParams(W, H), k, 1)[0]
        out.append((f"synth{W}", f, O.roi(W, H)["map"]))
    rng = np.random.default_rng(3)
    f = rng.integers(0, 256, (300, 400, 4), dtype=np.uint8)
    out.append(("noise", f, (10, 20, 301, 211)))
    flat = np.full((200, 260, 4), 90, np.uint8)
    out.append(("flat", flat, (5, 5, 240, 180)))
    # a flat "minimap" rectangle inside noise: the edges are where the flat runs end
    g = rng.integers(0, 256, (400, 500, 4), dtype=np.uint8)
    g[120:330, 90:420] = (60, 70, 80, 255)
    out.append(("flat_in_noise", g, (0, 0, 500, 400)))
    return out


@pytest.mark.parametrize("name,frame,view", _views(), ids=[v[0] for v in _views()])
def test_find_minimap_matches_numpy_restatement(name, frame, view):
    x, y, w, h = view
    got = O.find_minimap_in(frame, view)
    want = NP.find_minimap(frame[y:y + h, x:x + w])
    assert got is not None and tuple(got.values()) == want, (got, want)


def test_find_minimap_known_answers():
    flat = np.full((200, 260, 4), 90, np.uint8)
    # everything is below the threshold: the very first step in each direction qualifies and its run
    # ((other extent - centre) / 2 - 1 pixels) fits -> edge = first step minus the step = the centre itself
    assert O.find_minimap_in(flat, (5, 5, 240, 180)) == {"left": 120, "right": 120, "top": 90, "bottom": 90}
    rng = np.random.default_rng(1)
    noise = rng.integers(0, 256, (120, 160, 4), dtype=np.uint8)
    # nothing qualifies: left/top run into `< 3` -> 0, right/bottom past c_max = extent - 3 -> extent - 1
    assert O.find_minimap_in(noise, (0, 0, 160, 120)) == {"left": 0, "right": 159, "top": 0, "bottom": 119}
    assert O.find_minimap_in(noise, (0, 0, 2, 50)) is None
