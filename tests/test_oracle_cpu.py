"""CPU suite for the checker itself: the C oracle against (i) the independent numpy restatement,
(ii) the committed self-generated goldens, (iii) exhaustive colour tables, (iv) edge cases."""
import json

import numpy as np
import pytest

import smh_b200  # noqa: F401
from conftest import GOLDEN, sample_path, load_png_bgra
from helpers import oracle_run, sha
from oracle import np_oracle as N
from oracle import oracle_py as O
from squad_mortar_helper_b200 import synth

GOLD = json.loads((GOLDEN / "golden.json").read_text())


def _check_record(res, rec, what):
    assert res["map_open"] == rec["map_open"], what
    assert res["red_pixels"] == rec["red_pixels"], what
    if not rec["map_open"]:
        return
    for k in ("ui_map", "lsd_image", "ocr_out", "scales"):
        assert sha(res[k]) == rec[k + "_sha"], f"{what}: {k}"
    bits = [[f"{int(b):08x}" for b in row] for row in res["lines"].view(np.uint32)]
    assert bits == rec["lines_bits"], what
    assert list(res["lsd_stats"]) == [rec["lsd_seeds"], rec["lsd_ray_steps"]], what


def test_exhaustive_tables_match_golden_and_numpy():
    O.set_threads(8)
    tm, tl = O.table_marker(), O.table_luma()
    h, s, v = O.table_hsv()
    g = GOLD["tables"]
    assert sha(tm) == g["marker_sha"] and int(tm.sum()) == g["marker_count"]
    assert sha(tl) == g["luma_sha"]
    assert [sha(h), sha(s), sha(v)] == g["hsv_sha"]
    assert sha(O.ray_table()) == g["ray_table_sha"], "libm cosf/sinf differ from the golden box"
    c = np.arange(1 << 24, dtype=np.uint32)
    rgb = np.stack([(c >> 16) & 255, (c >> 8) & 255, c & 255], -1).astype(np.uint8)
    assert np.array_equal(N.is_marker(rgb), tm.astype(bool))
    assert np.array_equal(N.luma8(rgb), tl)
    nh, ns, nv = N.hsv(rgb)
    assert np.array_equal(nh, h) and np.array_equal(ns, s) and np.array_equal(nv, v)
    assert np.array_equal(N.ray_table().view(np.uint32), O.ray_table().view(np.uint32))


@pytest.mark.parametrize("name", list(GOLD["synthetic"]))
def test_oracle_matches_golden_synthetic(name):
    rec = GOLD["synthetic"][name]
    p = rec["params"]
    sp = synth.SynthParams(p["W"], p["H"], seed=int(p["seed"], 16), n_segments=p["n_segments"], n_blips=p["n_blips"],
                           n_labels=p["n_labels"], map_open=p["map_open"], blip_band_pct=p["blip_band_pct"])
    frame = synth.frames_host(sp, p["frame"], 1)[0]
    assert sha(frame) == rec["frame_sha"], "synthetic generator drifted"
    O.set_threads(8)
    _check_record(oracle_run(frame, p["grayscale"], p["start_y"], p["max_gap"]), rec, name)


@pytest.mark.parametrize("name", ["point_intersect.png", "points_intersect.png", "full_1024x768.png",
                                  "full_1600x1024.png", "snowpoints.png"])
def test_oracle_matches_golden_samples(name):
    path = sample_path(name)
    if path is None:
        pytest.skip("screenshot not available")
    frame = load_png_bgra(path)
    rec = GOLD["samples"][name]
    assert sha(frame) == rec["frame_sha"]
    O.set_threads(8)
    _check_record(oracle_run(frame, True, 0, 15), rec, name)
    if "gap22" in rec:
        _check_record(oracle_run(frame, True, 0, 22), rec["gap22"], name + " gap22")


def _np_pipeline(frame, grayscale, start_y, max_gap):
    c = N.crop_to_map(frame, grayscale)
    if c is None:
        return None
    raw, dil = N.mask_marker_lines(c["cropped_map"])
    return {"ui_map": c["ui_map"], "lsd_image": dil, "raw": raw, "ocr_out": N.ocr_preprocess(c["cropped_brq"]),
            "scales": N.find_scales_preprocess(c["cropped_brq"], start_y), "lines": N.find_lines(dil, max_gap),
            "red_pixels": c["red_pixels"], "xywh": c["xywh"]}


@pytest.mark.parametrize("W,H,kw,idx,gray,sy", [
    (1280, 720, dict(n_segments=4, n_blips=6, n_labels=3), 0, True, 0),
    (1024, 768, synth.DENSE, 1, False, 11),
    (1366, 768, dict(n_segments=6, n_blips=3, n_labels=2), 5, True, 200),
    (1920, 1080, dict(n_segments=5, n_blips=2, n_labels=4), 2, True, 0),
])
def test_oracle_matches_numpy_restatement(W, H, kw, idx, gray, sy):
    frame = synth.frames_host(synth.SynthParams(W, H, **kw), idx, 1)[0]
    O.set_threads(4)
    a = oracle_run(frame, gray, sy, 15)
    b = _np_pipeline(frame, gray, sy, 15)
    assert a["map_open"] and b is not None
    assert a["red_pixels"] == b["red_pixels"] and a["xywh"] == b["xywh"]
    for k in ("ui_map", "lsd_image", "ocr_out", "scales"):
        assert np.array_equal(a[k], b[k]), k
    assert a["lines"].shape == b["lines"].shape
    assert np.array_equal(a["lines"].view(np.uint32), b["lines"].view(np.uint32))


def test_numpy_restatement_on_real_screenshot():
    path = sample_path("point_intersect.png")
    if path is None:
        pytest.skip("screenshot not available")
    frame = load_png_bgra(path)
    O.set_threads(8)
    a = oracle_run(frame, True, 0, 22)
    b = _np_pipeline(frame, True, 0, 22)
    for k in ("ui_map", "lsd_image", "ocr_out", "scales"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["lines"].view(np.uint32), b["lines"].view(np.uint32))


def test_threaded_oracle_is_identical_to_scalar():
    frame = synth.frames_host(synth.SynthParams(1280, 720), 4, 1)[0]
    O.set_threads(1)
    a = oracle_run(frame)
    O.set_threads(8)
    b = oracle_run(frame)
    for k in ("ui_map", "lsd_image", "ocr_out", "scales"):
        assert np.array_equal(a[k], b[k])
    assert np.array_equal(a["lines"].view(np.uint32), b["lines"].view(np.uint32))
    assert a["lsd_stats"] == b["lsd_stats"]


def test_map_closed_returns_none_and_threshold_boundary():
    p = synth.SynthParams(1280, 720, map_open=False)
    frame = synth.frames_host(p, 0, 1)[0]
    res = oracle_run(frame)
    assert not res["map_open"] and res["red_pixels"] == 0
    # ratio test `red/(w*h) < 0.65` (vision-cpu/src/lib.rs:131-134): paint exactly k pixels red
    r = O.roi(1280, 720)
    bx, by, bw, bh = r["button"]
    need = int(np.ceil(0.65 * bw * bh))
    for k, want_open in ((need - 1, False), (need, True)):
        f = frame.copy()
        flat = [(by + i // bw, bx + i % bw) for i in range(k)]
        for (yy, xx) in flat:
            f[yy, xx] = (49, 67, 217, 255)
        o = O.Oracle(); o.load_frame(f)
        got = o.crop_to_map(True)
        assert o.red_pixels == k
        # f32 arithmetic decides the exact boundary
        assert (got is not None) == (np.float32(k) / np.float32(bw * bh) >= np.float32(0.65)) == want_open


def test_scales_rows_above_start_y_are_stale():
    fa = synth.frames_host(synth.SynthParams(1280, 720), 0, 1)[0]
    fb = synth.frames_host(synth.SynthParams(1280, 720, n_labels=4), 9, 1)[0]
    o = O.Oracle()
    a = oracle_run(fa, start_y=0, oracle=o)
    b = oracle_run(fb, start_y=120, oracle=o)   # same backend object: rows < 120 keep frame A's content
    assert np.array_equal(b["scales"][:120], a["scales"][:120])
    fresh = oracle_run(fb, start_y=120)
    assert not fresh["scales"][:120].any()
    assert np.array_equal(b["scales"][120:], fresh["scales"][120:])
    qh = a["scales"].shape[0]
    c = oracle_run(fb, start_y=qh, oracle=o)     # start_y == h: no-op
    assert np.array_equal(c["scales"], b["scales"])


def test_degenerate_frame_sizes_are_rejected():
    # the reference panics in par_crop_into / underflows for these (D12)
    for W, H in ((640, 1080), (500, 600), (800, 1000), (86, 100)):
        assert O.roi(W, H) is None, (W, H)
    assert O.roi(1920, 1080) is not None


def test_find_longest_line_quirks():
    # D7: a ray that leaves through the right/bottom edge yields a zero-length line; through the
    # left/top edge it may end at the border.  D6: ties -> highest angle index.
    img = np.zeros((40, 60), np.uint8)
    img[20, :] = 255                      # full-width horizontal line
    line, ln = O.find_longest_line_in(img, 30.0, 20.0, 15.0)
    l2, n2 = N.find_longest_line(img, 30.0, 20.0, 15.0)
    assert np.array_equal(line.view(np.uint32), l2.view(np.uint32)) and ln == n2
    # leftwards ray (180 deg) exits at x<0 -> checked pixel (0,20) is 255 -> end = start; the
    # rightwards ray exits at x>=w -> None -> end = start.  Only slightly tilted rays survive.
    assert ln < 30.0 ** 2 + 1
    img2 = np.zeros((50, 50), np.uint8)
    img2[10:40, 25] = 255
    for pt in ((25.0, 25.0), (25.5, 10.0), (0.0, 0.0), (49.0, 49.0)):
        a, la = O.find_longest_line_in(img2, pt[0], pt[1], 15.0)
        b, lb = N.find_longest_line(img2, pt[0], pt[1], 15.0)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and la == lb, pt
    # empty image, seed far from the borders: every ray stops at its very first (black) sample,
    # so the end point is pt - d and every length is ~1; the tie-break decides the winner
    z = np.zeros((100, 100), np.uint8)
    a, la = O.find_longest_line_in(z, 50.0, 50.0, 15.0)
    b, lb = N.find_longest_line(z, 50.0, 50.0, 15.0)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and la == lb and abs(float(la) - 1.0) < 1e-5
    # seed outside the image: the loop never runs, pixel (sat(x), sat(y)) decides (D7)
    a, la = O.find_longest_line_in(z, -5.0, 200.0, 15.0)
    b, lb = N.find_longest_line(z, -5.0, 200.0, 15.0)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and la == lb


def test_find_lines_small_cases():
    rng = np.random.default_rng(7)
    for trial in range(6):
        img = np.zeros((120, 160), np.uint8)
        import cv2
        for _ in range(3):
            p0 = (int(rng.integers(5, 155)), int(rng.integers(5, 115)))
            p1 = (int(rng.integers(5, 155)), int(rng.integers(5, 115)))
            cv2.line(img, p0, p1, 255, int(rng.integers(1, 5)))
        for _ in range(4):
            x, y = int(rng.integers(0, 158)), int(rng.integers(0, 118))
            img[y:y + 2, x:x + 2] = 255
        a = O.find_lines_in(img, 15)
        b = N.find_lines(img, 15)
        assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), trial


def test_cross_dilation_against_opencv():
    """The vision-cpu-only dilation (imageproc L1, k = 1 = 3x3 cross, D4) has no reference object code to run;
    it is checked against an independent implementation (OpenCV's cross-shaped dilate) instead."""
    import cv2
    frame = synth.frames_host(synth.SynthParams(1280, 720, n_segments=6, n_blips=8, n_labels=2), 2, 1)[0]
    o = O.Oracle()
    o.load_frame(frame); o.crop_to_map(True); o.isolate_map_markers(); o.mask_marker_lines()
    want = cv2.dilate(o.raw_mask(), cv2.getStructuringElement(cv2.MORPH_CROSS, (3, 3)), borderType=cv2.BORDER_CONSTANT, borderValue=0)
    assert np.array_equal(o.lsd_image(), want)


def test_vision_cpu_and_vision_gpu_march_flavours_agree_outside_d5():
    """The restated K10 march (pinned bit for bit against the reference's object code on the GPU box,
    tests/test_gpu_refcuda.py) and the vision-cpu march differ only in the gap_count rule (D5): on one
    direction table they must agree on every ray that rule does not decide."""
    rng = np.random.default_rng(21)
    dirs = O.ray_table()
    img = np.where(rng.random((120, 160)) < 0.9, 255, 0).astype(np.uint8)
    img[:, ::5] = 255
    n_d5 = 0
    for _ in range(12):
        sx, sy = float(rng.integers(0, 160)) + 0.5, float(rng.integers(0, 120))
        for gap in (15.0, 3.0, 1.0):
            g, d5 = O.gpuflavour_rays_in(img, sx, sy, gap, dirs)
            c = O.rays_in_dirs(img, sx, sy, gap, dirs)
            keep = d5 == 0
            assert np.array_equal(g.view(np.uint32)[keep], c.view(np.uint32)[keep])
            n_d5 += int((~keep).sum())
            # and rays_in_dirs with the host table is find_longest_line's own march
            lens = O.ray_lengths_in(img, sx, sy, gap)
            mine = (np.float32(sx) - c[:, 0]) ** 2 + (np.float32(sy) - c[:, 1]) ** 2
            assert np.array_equal(lens.view(np.uint32), mine.astype(np.float32).view(np.uint32))
    assert n_d5 > 0


def test_gpuflavour_dilate_is_the_2x2_box_model():
    rng = np.random.default_rng(2)
    raw = np.where(rng.random((40, 53)) < 0.2, 255, 0).astype(np.uint8)
    h, w = raw.shape
    box = np.zeros_like(raw)
    box[1:h - 1, 1:w - 1] = np.maximum.reduce([raw[0:h - 2, 0:w - 2], raw[0:h - 2, 1:w - 1], raw[1:h - 1, 0:w - 2], raw[1:h - 1, 1:w - 1]])
    assert np.array_equal(O.gpuflavour_dilate(raw), box)


def test_marker_predicate_against_the_reference_kernels_own_table():
    """tests/golden/k8_marker_table.npz is an OUTPUT OF THE REFERENCE: its K8 (mask_marker_lines, vision-gpu/cuda/cuda.cu:616-635)
    compiled unmodified and run over all 2^24 colours on a B200 (tools/k8_table.py).  The oracle's restatement of
    vision-gpu's predicate must reproduce it exactly; vision-cpu's predicate (what the product is checked against)
    differs from it on exactly the colours where the two HSV formulas round apart (D1) — each of them one step of
    h, s or v away from flipping."""
    k8 = np.unpackbits(np.load(GOLDEN / "k8_marker_table.npz")["bits"]).astype(bool)
    assert np.array_equal(O.gpuflavour_table_marker() == 1, k8)
    cpu = O.table_marker() == 1
    diff = np.nonzero(cpu != k8)[0]
    assert len(diff) == 136
    th, ts, tv = O.table_hsv()
    hh, ss, vv = th[diff].astype(np.int32), ts[diff].astype(np.int32), tv[diff].astype(np.int32)

    def pred(h, s, v):
        ok = np.zeros(h.shape, bool)
        for mh, ms, mv in ((105, 100, 100), (285, 46, 85), (158, 60, 91)):
            ok |= (np.abs(h - mh) <= 15) & ((np.abs(s - ms) <= 15) | (np.abs(s - (ms - 50)) <= 15)) & (np.abs(v - mv) <= 15)
        return ok & (s >= 35)
    base = pred(hh, ss, vv)
    assert np.array_equal(base, cpu[diff])
    flips = np.zeros(len(diff), bool)
    for dh in (-1, 0, 1):
        for ds in (-1, 0, 1):
            for dv in (-1, 0, 1):
                flips |= pred((hh + dh) % 360, ss + ds, vv + dv) != base
    assert flips.all()
