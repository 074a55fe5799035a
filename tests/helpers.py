"""Shared test helpers: run a frame through the oracle / through the CUDA backend and compare."""
from __future__ import annotations

import hashlib

import numpy as np


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def oracle_run(frame, grayscale=True, start_y=0, max_gap=15, oracle=None):
    """Whole per-frame pipeline on the CPU oracle -> dict of every trait-visible output."""
    from oracle import oracle_py as O
    o = oracle or O.Oracle()
    o.load_frame(frame)
    got = o.crop_to_map(grayscale)
    out = {"red_pixels": o.red_pixels, "map_open": got is not None, "roi": o.roi}
    if got is None:
        return out
    out["ui_map"], out["xywh"] = got
    o.isolate_map_markers()
    out["cropped_map_isolated"] = o.cropped_map()
    o.mask_marker_lines()
    out["lsd_image"] = o.lsd_image()
    out["lines"] = o.find_marker_lines(max_gap)
    out["lsd_stats"] = o.lsd_stats()
    out["ocr_out"] = o.ocr_preprocess()
    out["scales"] = o.find_scales_preprocess(start_y)
    out["cropped_brq"] = o.cropped_brq()
    return out


def gpu_run(v, frame, grayscale=True, start_y=0, max_gap=15):
    """Same call order as VisionState::process (src/vision/mod.rs:66-223) on the CUDA backend."""
    v.load_frame(frame)
    got = v.crop_to_map(grayscale)
    out = {"red_pixels": v.red_pixels(), "map_open": got is not None}
    if got is None:
        return out
    out["ui_map"], out["xywh"] = got
    v.thread_ctx()
    v.isolate_map_markers()
    v.mask_marker_lines()
    out["lines"] = v.find_marker_lines(max_gap)
    out["lsd_image"] = v.lsd_image().copy()
    v.thread_ctx()
    out["ocr_out"] = v.ocr_preprocess().copy()
    out["scales"] = v.find_scales_preprocess(start_y).copy()
    return out


def assert_same(got: dict, want: dict, what="", scales_from=0):
    """scales_from: first row of `scales` to compare — a reused backend legitimately keeps an
    earlier frame's rows above start_y (D9); the stale-row semantics have their own test."""
    assert got["map_open"] == want["map_open"], f"{what}: map_open"
    assert got["red_pixels"] == want["red_pixels"], f"{what}: red_pixels {got['red_pixels']} != {want['red_pixels']}"
    if not want["map_open"]:
        return
    assert list(got["xywh"]) == list(want["xywh"]), f"{what}: xywh"
    for k in ("ui_map", "lsd_image", "ocr_out", "scales"):
        g, w = got[k], want[k]
        if k == "scales" and scales_from:
            g, w = g[scales_from:], w[scales_from:]
        assert g.shape == w.shape, f"{what}: {k} shape {g.shape} != {w.shape}"
        if not np.array_equal(g, w):
            bad = np.argwhere(g != w)
            raise AssertionError(f"{what}: {k} differs at {len(bad)} positions, first {bad[0].tolist()} "
                                 f"got {g[tuple(bad[0])]} want {w[tuple(bad[0])]}")
    gl, wl = got["lines"], want["lines"]
    assert gl.shape == wl.shape, f"{what}: {gl.shape[0]} lines, oracle has {wl.shape[0]}\n{gl}\n{wl}"
    # bit-exact: compare the raw f32 bit patterns
    assert np.array_equal(gl.view(np.uint32), wl.view(np.uint32)), f"{what}: line endpoints differ\n{gl}\n{wl}"


def oracle_process(frame, grayscale=True, detect_markers=True, max_gap=15, ocr=None, oracle=None):
    """VisionState::process (src/vision/mod.rs:36-240) composed from the oracle's pieces.
    ocr: None | callable(ocr_l8) -> ([(meters, (x, y))], scales_start_y).  -> None | dict like B200Vision.process"""
    from oracle import oracle_py as O
    o = oracle or O.Oracle()
    o.load_frame(frame)
    got = o.crop_to_map(grayscale)
    if got is None:
        return None
    ui, xywh = got
    res = {"map": ui, "xywh": xywh, "markers": np.zeros((0, 4), np.float32), "meters_to_px_ratio": None, "scales_debug": []}
    res["minimap_bounds"] = O.find_minimap_in(frame, xywh)
    if detect_markers:
        o.isolate_map_markers()
        o.mask_marker_lines()
        res["markers"] = o.find_marker_lines(max_gap)
    if ocr is not None:
        scales, start_y = ocr(o.ocr_preprocess())
        scales = list(scales)[:3]
        if scales:
            img = o.find_scales_preprocess(start_y)
            res["meters_to_px_ratio"], res["scales_debug"] = O.calc_meters_to_px_ratio(img, scales, with_debug=True)
    return res


def marker_templates(size=22):
    """The committed fixture of `load_markers(22)` (tests/golden/make_marker_templates.py)."""
    from pathlib import Path
    return np.load(Path(__file__).with_name("golden") / f"marker_templates_{size}.npy")


def stamp_icon(frame, map_xywh, template, x, y):
    """Draw a marker icon into a BGRA frame: template pixels with alpha != 0 replace the map pixel at
    map coordinates (x + mx, y + my)."""
    mx0, my0 = map_xywh[0], map_xywh[1]
    s = template.shape[0]
    sel = template[..., 3] != 0
    patch = frame[my0 + y:my0 + y + s, mx0 + x:mx0 + x + s]
    patch[sel] = np.concatenate([template[..., 2::-1], np.full((s, s, 1), 255, np.uint8)], axis=2)[sel]
