"""Parity tests proper: the CUDA backend, called through the C ABI, against the CPU oracle —
bit-exact masks, images and line endpoints (f32 bit patterns)."""
import json
import threading

import numpy as np
import pytest
import torch

import smh_b200  # noqa: F401
from conftest import GOLDEN, sample_path, load_png_bgra
from helpers import assert_same, gpu_run, oracle_run, sha
from oracle import oracle_py as O
from squad_mortar_helper_b200 import capi, synth
from squad_mortar_helper_b200.vision import B200Vision, DebugView, VisionError

pytestmark = pytest.mark.gpu
GOLD = json.loads((GOLDEN / "golden.json").read_text())


@pytest.fixture(scope="module")
def vision():
    O.set_threads(8)
    v = B200Vision.init(0)
    yield v
    v.close()


def _golden_frame(name):
    rec = GOLD["synthetic"][name]
    p = rec["params"]
    sp = synth.SynthParams(p["W"], p["H"], seed=int(p["seed"], 16), n_segments=p["n_segments"], n_blips=p["n_blips"],
                           n_labels=p["n_labels"], map_open=p["map_open"], blip_band_pct=p["blip_band_pct"])
    return synth.frames_host(sp, p["frame"], 1)[0], p, rec


# ---- exhaustive colour tables evaluated on the device ---------------------------------------
def test_device_colour_tables_exhaustive(vision):
    assert np.array_equal(vision.colour_table(0), O.table_marker()), "marker predicate (kernel path: fast-reject + table) != oracle"
    assert np.array_equal(vision.colour_table(4), O.table_marker()), "marker predicate (float formula on device) != oracle"
    assert np.array_equal(vision.colour_table(1), O.table_luma()), "luma8 (kernel path) != oracle"
    assert np.array_equal(vision.colour_table(5), O.table_luma()), "luma8 (float formula on device) != oracle"
    assert np.array_equal(vision.colour_table(6), O.table_luma()), "luma8 (biased-word path) != oracle"
    c = np.arange(1 << 24, dtype=np.int64)
    r, g, b = (c >> 16) & 255, (c >> 8) & 255, c & 255
    red = (np.abs(217 - r) <= 25) & (np.abs(67 - g) <= 25) & (np.abs(49 - b) <= 25)
    assert np.array_equal(vision.colour_table(2), red.astype(np.uint8))


# ---- the committed golden cases (BASELINE configs 2-4 frames + edge cases) -------------------
@pytest.mark.parametrize("name", list(GOLD["synthetic"]))
def test_golden_synthetic_cases(vision, name):
    frame, p, rec = _golden_frame(name)
    got = gpu_run(vision, frame, p["grayscale"], p["start_y"], p["max_gap"])
    want = oracle_run(frame, p["grayscale"], p["start_y"], p["max_gap"])
    assert_same(got, want, name, scales_from=p["start_y"])
    # and against the committed hashes (catches an oracle that drifted together with the kernels)
    assert got["red_pixels"] == rec["red_pixels"] and got["map_open"] == rec["map_open"]
    if rec["map_open"]:
        # scales of a reused backend keep stale rows above start_y: compare the fresh part only
        for k in ("ui_map", "lsd_image", "ocr_out"):
            assert sha(got[k]) == rec[k + "_sha"], k
        assert [[f"{int(x):08x}" for x in row] for row in got["lines"].view(np.uint32)] == rec["lines_bits"]


def test_config2_1080p_full_pipeline_fresh_backend():
    """BASELINE config 2: one 1920x1080 synthetic frame, every output bit-exact."""
    frame, p, rec = _golden_frame("1080p_dense_f0")
    with B200Vision.init(0) as v:
        got = gpu_run(v, frame, True, 0, 15)
    assert_same(got, oracle_run(frame, True, 0, 15), "config2")
    assert sha(got["scales"]) == rec["scales_sha"]


@pytest.mark.parametrize("name", ["point_intersect.png", "points_intersect.png", "full_1024x768.png",
                                  "full_1280x1024.png", "full_1600x1024.png", "difficult.png"])
def test_real_screenshots(vision, name):
    path = sample_path(name)
    if path is None:
        pytest.skip("screenshot not shipped to this box")
    frame = load_png_bgra(path)
    for gray, gap in ((True, 15), (False, 22)):
        got = gpu_run(vision, frame, gray, 0, gap)
        assert_same(got, oracle_run(frame, gray, 0, gap), f"{name} gray={gray} gap={gap}")
    rec = GOLD["samples"][name]
    got = gpu_run(vision, frame, True, 0, 15)
    assert [[f"{int(x):08x}" for x in row] for row in got["lines"].view(np.uint32)] == rec["lines_bits"]
    assert sha(got["lsd_image"]) == rec["lsd_image_sha"]


@pytest.mark.parametrize("W,H", [(1280, 720), (1366, 768), (1600, 900), (1920, 1200), (2560, 1080), (3440, 1440),
                                 (1281, 721), (1279, 717), (1000, 1000), (5120, 1440)])
def test_odd_resolutions(vision, W, H):
    frame = synth.frames_host(synth.SynthParams(W, H, n_segments=5, n_blips=4, n_labels=3), 1, 1)[0]
    assert_same(gpu_run(vision, frame, True, 3, 15), oracle_run(frame, True, 3, 15), f"{W}x{H}", scales_from=3)


@pytest.mark.parametrize("W,H", [(160, 120), (200, 150), (320, 240), (640, 480), (7680, 4320), (16384, 2160)])
def test_extreme_sizes(vision, W, H):
    """The smallest frames the reference's ROI arithmetic accepts (map 56 x 91: narrower than one strip of the
    pixel pass, smaller than the line search's window), 8K, and the widest frame the ABI takes."""
    frame = synth.frames_host(synth.SynthParams(W, H, n_segments=3, n_blips=3, n_labels=2), 2, 1)[0]
    assert_same(gpu_run(vision, frame, True, 0, 15), oracle_run(frame, True, 0, 15), f"{W}x{H}")
    # and through the batch form, twice the same frame
    dev = torch.from_numpy(np.stack([frame, frame])).cuda()
    b = vision.process_batch(dev.data_ptr(), 2, W, H, max_gap=15, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want = oracle_run(frame, True, 0, 15)
    for i in range(2):
        assert np.array_equal(vision.download_plane(b.lsd_image, i, 1)[0], want["lsd_image"])
        assert np.array_equal(vision.download_plane(b.ui_map, i, 1)[0], want["ui_map"])


def test_random_noise_frames(vision):
    """Uniform random colours: exercises the exact HSV path on ~5 % of pixels, every OCR class
    and a saturated mask (32-line early exit)."""
    rng = np.random.default_rng(11)
    for W, H in ((1280, 720), (1024, 768)):
        base = synth.frames_host(synth.SynthParams(W, H, **synth.PLAIN), 0, 1)[0]
        r = O.roi(W, H)
        x, y, w, h = r["map"]
        f = base.copy()
        f[y:y + h, x:x + w, :3] = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        # grey-ish bright pixels in the bottom-right quadrant for the OCR classes
        qx, qy, qw, qh = r["brq"]
        g = rng.integers(120, 256, (qh, qw, 1), dtype=np.uint8)
        jitter = rng.integers(0, 14, (qh, qw, 3), dtype=np.uint8) * (rng.random((qh, qw, 1)) < 0.5)
        f[qy:qy + qh, qx:qx + qw, :3] = np.clip(g.astype(np.int32) + jitter, 0, 255).astype(np.uint8)
        f[qy + 5:qy + 9, qx:qx + qw, :3] = 0   # luma == 0 rows for the scales threshold
        assert_same(gpu_run(vision, f, True, 0, 15), oracle_run(f, True, 0, 15), f"noise {W}x{H}")
        assert_same(gpu_run(vision, f, False, 2, 4), oracle_run(f, False, 2, 4), f"noise rgb {W}x{H}", scales_from=2)


@pytest.mark.parametrize("W,H", [(2560, 1440), (1920, 1080), (1281, 721)])
def test_strip_seams(vision, W, H):
    """The pixel pass works in strips of 128 (padded) columns and reads nothing of the neighbouring strip: the
    3x3 cross and the OCR 7x7 window reach across a seam through k_px_seams.  Content ONLY in the six columns around
    every seam (marker colours everywhere, bright grey / near-grey in the OCR quadrant), sparse enough that a
    pixel's only partners are across the seam."""
    rng = np.random.default_rng(W * 7 + H)
    base = synth.frames_host(synth.SynthParams(W, H, **synth.PLAIN), 0, 1)[0]
    r = O.roi(W, H)
    x, y, w, h = r["map"]
    qx, qy, qw, qh = r["brq"]
    pad = x & 3
    for density in (0.02, 0.3):
        f = base.copy()
        roi = f[y:y + h, x:x + w, :3]
        roi[:] = (40, 60, 50)  # nothing: dark, not grey
        noise = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        grey = rng.integers(125, 256, (h, w, 1), dtype=np.uint8)
        jitter = rng.integers(0, 14, (h, w, 3), dtype=np.uint8) * (rng.random((h, w, 1)) < 0.5)
        greyish = np.clip(grey.astype(np.int32) + jitter, 0, 255).astype(np.uint8)
        near = np.zeros(w, dtype=bool)
        for s in range(1, (pad + w + 127) // 128):
            c = 128 * s - pad  # map column of the strip's first pixel
            near[max(c - 3, 0):min(c + 3, w)] = True
        pick = near[None, :] & (rng.random((h, w)) < density)
        inq = np.zeros((h, w), dtype=bool)
        inq[qy - y:qy - y + qh, qx - x:qx - x + qw] = True
        roi[pick & ~inq] = noise[pick & ~inq]
        roi[pick & inq] = greyish[pick & inq]
        half = pick & inq & (rng.random((h, w)) < 0.3)
        roi[half] = noise[half]  # marker colours inside the quadrant as well
        assert_same(gpu_run(vision, f, True, 0, 15), oracle_run(f, True, 0, 15), f"seams {W}x{H} {density}")
        assert_same(gpu_run(vision, f, False, 1, 2), oracle_run(f, False, 1, 2), f"seams rgb {W}x{H} {density}", scales_from=1)


def test_map_closed_and_threshold_boundary(vision):
    p = synth.SynthParams(1280, 720, map_open=False)
    frame = synth.frames_host(p, 0, 1)[0]
    vision.load_frame(frame)
    assert vision.crop_to_map(True) is None
    with pytest.raises(VisionError) as e:
        vision.find_marker_lines(15)
    assert e.value.code == capi.SMH_E_STATE
    r = O.roi(1280, 720)
    bx, by, bw, bh = r["button"]
    need = int(np.ceil(0.65 * bw * bh))
    for k in (need - 1, need, need + 1):
        f = frame.copy()
        for i in range(k):
            f[by + i // bw, bx + i % bw] = (49, 67, 217, 255)
        got, want = gpu_run(vision, f), oracle_run(f)
        assert got["red_pixels"] == k == want["red_pixels"]
        assert got["map_open"] == want["map_open"]


def test_scales_stale_rows_and_start_y_range(vision):
    fa = synth.frames_host(synth.SynthParams(1280, 720), 0, 1)[0]
    fb = synth.frames_host(synth.SynthParams(1280, 720, n_labels=4), 9, 1)[0]
    o = O.Oracle()
    with B200Vision.init(0) as v:
        a_g, a_o = gpu_run(v, fa, start_y=0), oracle_run(fa, start_y=0, oracle=o)
        assert np.array_equal(a_g["scales"], a_o["scales"])
        b_g, b_o = gpu_run(v, fb, start_y=120), oracle_run(fb, start_y=120, oracle=o)
        assert np.array_equal(b_g["scales"], b_o["scales"])          # rows < 120 are frame A's
        qh = b_g["scales"].shape[0]
        assert np.array_equal(v.find_scales_preprocess(qh), b_o["scales"])   # start_y == h: no-op
        with pytest.raises(VisionError) as e:
            v.find_scales_preprocess(qh + 1)
        assert e.value.code == capi.SMH_E_RANGE
    with B200Vision.init(0) as v:                                     # fresh backend: zeros above
        g = gpu_run(v, fb, start_y=120)
        assert not g["scales"][:120].any()


def test_find_longest_line_matches_oracle(vision):
    frame, p, _ = _golden_frame("1080p_dense_f0")
    want = oracle_run(frame)
    gpu_run(vision, frame)
    mask = want["lsd_image"]
    ys, xs = np.nonzero(mask == 255)
    rng = np.random.default_rng(3)
    pts = [(float(xs[i]), float(ys[i])) for i in rng.choice(len(xs), 12, replace=False)]
    pts += [(0.0, 0.0), (mask.shape[1] - 1.0, mask.shape[0] - 1.0), (100.5, 200.25), (-3.0, 10.0), (5.0, 1e6)]
    for gap in (15.0, 22.0, 0.0, 2.5):
        for pt in pts:
            gl, gn = vision.find_longest_line(pt, gap)
            ol, on = O.find_longest_line_in(mask, pt[0], pt[1], gap)
            assert np.array_equal(gl.view(np.uint32), ol.view(np.uint32)), (pt, gap, gl, ol)
            assert np.float32(gn).view(np.uint32) == np.float32(on).view(np.uint32), (pt, gap)


def test_debug_views(vision):
    frame = synth.frames_host(synth.SynthParams(1280, 720, n_labels=3), 2, 1)[0]
    o = O.Oracle()
    want = oracle_run(frame, True, 7, 15, oracle=o)
    with B200Vision.init(0) as v:
        gpu_run(v, frame, True, 7, 15)
        assert v.get_debug_view(DebugView.NONE) is None
        for which in (DebugView.OCRInput, DebugView.FindScalesInput, DebugView.LSDPreprocess, DebugView.LSDInput,
                      DebugView.CroppedBRQ):
            got = v.get_debug_view(which)
            ref = o.get_debug_view(int(which))
            assert got.shape == ref.shape and np.array_equal(got, ref), which


def test_call_order_errors(vision):
    with B200Vision.init(0) as v:
        with pytest.raises(VisionError) as e:
            v._frame = np.zeros((720, 1280, 4), np.uint8)
            v.crop_to_map(True)
        assert e.value.code == capi.SMH_E_STATE
        with pytest.raises(VisionError) as e:
            v.load_frame(np.zeros((1080, 640, 4), np.uint8))   # the reference panics on this size
        assert e.value.code == capi.SMH_E_GEOMETRY


def test_two_branches_concurrently(vision):
    """markers || scales from two threads on one handle (src/vision/mod.rs:219-223)."""
    frame, p, _ = _golden_frame("1440p_dense_f0")
    want = oracle_run(frame)
    for _ in range(3):
        vision.load_frame(frame)
        assert vision.crop_to_map(True) is not None
        res = {}

        def markers():
            vision.thread_ctx(); vision.isolate_map_markers(); vision.mask_marker_lines()
            res["lines"] = vision.find_marker_lines(15)

        def scales():
            vision.thread_ctx(); res["ocr"] = vision.ocr_preprocess().copy()
            res["scales"] = vision.find_scales_preprocess(0).copy()

        ts = [threading.Thread(target=markers), threading.Thread(target=scales)]
        [t.start() for t in ts]; [t.join() for t in ts]
        assert np.array_equal(res["lines"].view(np.uint32), want["lines"].view(np.uint32))
        assert np.array_equal(res["ocr"], want["ocr_out"]) and np.array_equal(res["scales"], want["scales"])


def test_repeatability(vision):
    frame, p, _ = _golden_frame("1440p_dense_f7_gap22")
    a = gpu_run(vision, frame, True, 100, 22)
    for _ in range(3):
        b = gpu_run(vision, frame, True, 100, 22)
        assert np.array_equal(a["lines"].view(np.uint32), b["lines"].view(np.uint32))
        assert np.array_equal(a["lsd_image"], b["lsd_image"]) and np.array_equal(a["ui_map"], b["ui_map"])
