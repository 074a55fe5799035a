"""ctypes binding of the CPU ORACLE (oracle/libsmh_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (squad-mortar-helper_b200/) never does.

PARITY STATUS: pinned against the reference's own compiled CUDA kernels (oracle/_ref) — see oracle/smh_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    so = _HERE / "libsmh_oracle.so"
    src = [_HERE / "smh_oracle.c", _HERE / "smh_oracle.h", _HERE.parent / "include" / "smh_consts.h"]
    stale = (not so.exists()) or any(s.stat().st_mtime > so.stat().st_mtime for s in src)
    if force or stale:
        subprocess.run(["make", "-C", str(_HERE), "libsmh_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(str(build()))
    u8p, u32p, f32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_float)
    vp = C.c_void_p
    L.smh_oracle_new.restype = vp
    L.smh_oracle_free.argtypes = [vp]
    L.smh_oracle_roi.argtypes = [C.c_uint32, C.c_uint32, u32p]
    L.smh_oracle_luma8.restype = C.c_uint8
    L.smh_oracle_luma8.argtypes = [C.c_uint8] * 3
    L.smh_oracle_is_marker.argtypes = [C.c_uint8] * 3
    L.smh_oracle_hsv.argtypes = [C.c_uint8] * 3 + [C.POINTER(C.c_uint16), u8p, u8p]
    L.smh_oracle_table_marker.argtypes = [vp]
    L.smh_oracle_table_luma.argtypes = [vp]
    L.smh_oracle_table_hsv.argtypes = [vp, vp, vp]
    L.smh_oracle_ray_table.argtypes = [vp]
    L.smh_oracle_set_threads.argtypes = [C.c_int]
    L.smh_oracle_load_frame.argtypes = [vp, vp, C.c_uint32, C.c_uint32]
    L.smh_oracle_crop_to_map.argtypes = [vp, C.c_int, C.POINTER(C.c_int), u32p, C.POINTER(vp), u32p]
    L.smh_oracle_ocr_preprocess.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.smh_oracle_find_scales_preprocess.argtypes = [vp, C.c_uint32, C.POINTER(vp), u32p]
    L.smh_oracle_isolate_map_markers.argtypes = [vp]
    L.smh_oracle_mask_marker_lines.argtypes = [vp]
    L.smh_oracle_load_map_markers.argtypes = [vp, vp, C.c_uint32, C.c_uint32]
    L.smh_oracle_filter_map_marker_icons.argtypes = [vp, C.POINTER(C.c_int), u32p, u32p]
    L.smh_oracle_find_longest_line.argtypes = [vp, C.c_float, C.c_float, C.c_float, f32p, f32p]
    L.smh_oracle_find_longest_line_in.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_float, C.c_float,
                                                  C.c_float, f32p, f32p]
    L.smh_oracle_find_marker_lines.argtypes = [vp, C.c_uint32, vp, u32p]
    L.smh_oracle_find_lines_in.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, u32p]
    L.smh_oracle_ray_lengths_in.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_float, vp]
    L.smh_oracle_get_centre_in.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_float, C.c_float, f32p]
    L.smh_oracle_debug_view.argtypes = [vp, C.c_uint8, vp, u32p]
    for g in ("cropped_map", "cropped_brq", "lsd_image", "scales", "raw_mask"):
        f = getattr(L, "smh_oracle_" + g)
        f.restype = vp
        f.argtypes = [vp, u32p]
    L.smh_oracle_lsd_stats.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.smh_oracle_process_frame.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32,
                                           C.c_uint32, vp, u32p]
    L.smh_oracle_calc_meters_to_px_ratio.argtypes = [vp, C.c_uint32, C.c_uint32, vp, C.c_uint32, C.POINTER(C.c_double), vp, u32p]
    L.smh_oracle_find_minimap_in.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, u32p]
    L.smh_oracle_find_minimap.argtypes = [vp, u32p]
    L.smh_oracle_gpuflavour_table_marker.argtypes = [vp]
    L.smh_oracle_synth_frames.argtypes = [vp, C.c_uint32, C.c_uint32, vp]
    L.smh_oracle_gpuflavour_rays_in.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_float, vp, vp, vp]
    L.smh_oracle_rays_in_dirs.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_float, vp, vp]
    L.smh_oracle_gpuflavour_dilate.argtypes = [vp, C.c_uint32, C.c_uint32, vp]
    L.smh_oracle_gpuflavour_icon_sad.restype = C.c_uint16
    L.smh_oracle_gpuflavour_icon_sad.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, C.c_uint32]
    L.smh_oracle_gpuflavour_icon_clear.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    _LIB = L
    return L


def set_threads(n: int) -> None:
    lib().smh_oracle_set_threads(int(n))


def roi(W: int, H: int):
    """-> dict(button=(x,y,w,h), map=(x,y,w,h), brq=(x,y,w,h)) or None if the reference would panic."""
    out = (C.c_uint32 * 12)()
    if lib().smh_oracle_roi(W, H, out) != 0:
        return None
    v = list(out)
    return {"button": tuple(v[0:4]), "map": tuple(v[4:8]), "brq": tuple(v[8:12])}


def _view(ptr, shape):
    n = int(np.prod(shape))
    buf = (C.c_uint8 * n).from_address(ptr)
    return np.frombuffer(buf, dtype=np.uint8).reshape(shape).copy()


def table_marker() -> np.ndarray:
    out = np.empty(1 << 24, np.uint8)
    lib().smh_oracle_table_marker(out.ctypes.data)
    return out


def table_luma() -> np.ndarray:
    out = np.empty(1 << 24, np.uint8)
    lib().smh_oracle_table_luma(out.ctypes.data)
    return out


def table_hsv():
    h = np.empty(1 << 24, np.uint16)
    s = np.empty(1 << 24, np.uint8)
    v = np.empty(1 << 24, np.uint8)
    lib().smh_oracle_table_hsv(h.ctypes.data, s.ctypes.data, v.ctypes.data)
    return h, s, v


def ray_table() -> np.ndarray:
    out = np.empty((3600, 2), np.float32)
    lib().smh_oracle_ray_table(out.ctypes.data)
    return out


class Oracle:
    """Mirrors `CPUFallback: Vision` (vision-cpu/src/lib.rs) method for method."""

    def __init__(self):
        self.L = lib()
        self.h = self.L.smh_oracle_new()
        self.W = self.H = 0

    def close(self):
        if self.h:
            self.L.smh_oracle_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def load_frame(self, bgra: np.ndarray) -> None:
        assert bgra.dtype == np.uint8 and bgra.ndim == 3 and bgra.shape[2] == 4
        bgra = np.ascontiguousarray(bgra)
        self.H, self.W = bgra.shape[:2]
        rc = self.L.smh_oracle_load_frame(self.h, bgra.ctypes.data, self.W, self.H)
        if rc:
            raise ValueError(f"oracle load_frame rc={rc} (reference would panic)")
        self.roi = roi(self.W, self.H)

    def crop_to_map(self, grayscale: bool):
        """-> None (map closed) or (ui_map[h,w,4], [x,y,w,h]); self.red_pixels is set either way."""
        mo = C.c_int(0)
        xywh = (C.c_uint32 * 4)()
        p = C.c_void_p()
        red = C.c_uint32(0)
        rc = self.L.smh_oracle_crop_to_map(self.h, int(grayscale), C.byref(mo), xywh, C.byref(p), C.byref(red))
        assert rc == 0
        self.red_pixels = red.value
        if not mo.value:
            return None
        x, y, w, h = list(xywh)
        return _view(p.value, (h, w, 4)), [x, y, w, h]

    def ocr_preprocess(self) -> np.ndarray:
        p = C.c_void_p()
        n = C.c_size_t()
        assert self.L.smh_oracle_ocr_preprocess(self.h, C.byref(p), C.byref(n)) == 0
        _, _, w, h = self.roi["brq"]
        assert n.value == w * h
        return _view(p.value, (h, w))

    def find_scales_preprocess(self, start_y: int) -> np.ndarray:
        p = C.c_void_p()
        wh = (C.c_uint32 * 2)()
        assert self.L.smh_oracle_find_scales_preprocess(self.h, start_y, C.byref(p), wh) == 0
        return _view(p.value, (wh[1], wh[0]))

    def isolate_map_markers(self) -> None:
        assert self.L.smh_oracle_isolate_map_markers(self.h) == 0

    def mask_marker_lines(self) -> None:
        assert self.L.smh_oracle_mask_marker_lines(self.h) == 0

    def load_map_markers(self, templates: np.ndarray) -> None:
        """templates: [n, size, size, 4] RGBA (UnfilteredMarkers layout, markers/mod.rs:142-152)."""
        t = np.ascontiguousarray(templates, np.uint8)
        assert t.ndim == 4 and t.shape[1] == t.shape[2] and t.shape[3] == 4
        assert self.L.smh_oracle_load_map_markers(self.h, t.ctypes.data, t.shape[0], t.shape[1]) == 0

    def filter_map_marker_icons(self):
        """The disabled template-match stage (vision-cpu/src/lib.rs:282-355) -> None | (x, y, sad)."""
        found = C.c_int()
        xy = (C.c_uint32 * 2)()
        sad = C.c_uint32()
        assert self.L.smh_oracle_filter_map_marker_icons(self.h, C.byref(found), xy, C.byref(sad)) == 0
        return (int(xy[0]), int(xy[1]), int(sad.value)) if found.value else None

    def find_longest_line(self, x: float, y: float, max_gap: float):
        line = (C.c_float * 4)()
        ln = C.c_float()
        assert self.L.smh_oracle_find_longest_line(self.h, x, y, max_gap, line, C.byref(ln)) == 0
        return np.array(list(line), np.float32), np.float32(ln.value)

    def find_marker_lines(self, max_gap: int) -> np.ndarray:
        lines = np.zeros((32, 4), np.float32)
        n = C.c_uint32()
        assert self.L.smh_oracle_find_marker_lines(self.h, max_gap, lines.ctypes.data, C.byref(n)) == 0
        return lines[: n.value].copy()

    def lsd_stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        self.L.smh_oracle_lsd_stats(C.byref(a), C.byref(b))
        return a.value, b.value

    def get_debug_view(self, which: int):
        wh = (C.c_uint32 * 2)()
        rc = self.L.smh_oracle_debug_view(self.h, which, None, wh)
        if rc == 1:
            return None
        assert rc == 0
        out = np.empty((wh[1], wh[0], 4), np.uint8)
        self.L.smh_oracle_debug_view(self.h, which, out.ctypes.data, wh)
        return out

    def _img(self, name, ch):
        wh = (C.c_uint32 * 2)()
        p = getattr(self.L, "smh_oracle_" + name)(self.h, wh)
        shape = (wh[1], wh[0]) if ch == 1 else (wh[1], wh[0], ch)
        return _view(p, shape)

    def cropped_map(self): return self._img("cropped_map", 3)
    def cropped_brq(self): return self._img("cropped_brq", 3)
    def lsd_image(self): return self._img("lsd_image", 1)
    def raw_mask(self): return self._img("raw_mask", 1)
    def scales(self): return self._img("scales", 1)

    def process_frame(self, bgra: np.ndarray, grayscale=True, start_y=0, max_gap=15):
        """Whole pipeline in the orchestrator's order; -> (map_open, lines[n,4])."""
        bgra = np.ascontiguousarray(bgra)
        H, W = bgra.shape[:2]
        lines = np.zeros((32, 4), np.float32)
        n = C.c_uint32()
        rc = self.L.smh_oracle_process_frame(self.h, bgra.ctypes.data, W, H, int(grayscale), start_y,
                                             max_gap, lines.ctypes.data, C.byref(n))
        if rc < 0:
            raise ValueError(f"oracle process_frame rc={rc}")
        self.W, self.H = W, H
        self.roi = roi(W, H)
        return bool(rc), lines[: n.value].copy()


def calc_meters_to_px_ratio(img: np.ndarray, scales, with_debug: bool = False):
    """src/vision/mpx_ratio.rs:3-133.  scales: [(meters, (x, y)), ...] -> None | float (f64)."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    arr = np.zeros((3, 3), np.uint32)
    for k, (m, (x, y)) in enumerate(scales):
        arr[k] = (m, x, y)
    ratio = C.c_double()
    dbg = np.zeros((3, 4), np.uint32)
    nd = C.c_uint32()
    ok = lib().smh_oracle_calc_meters_to_px_ratio(img.ctypes.data, w, h, arr.ctypes.data, len(scales), C.byref(ratio),
                                                  dbg.ctypes.data, C.byref(nd))
    r = ratio.value if ok else None
    if with_debug:
        return r, [tuple(int(v) for v in dbg[k]) for k in range(nd.value)]
    return r


def find_minimap_in(bgra: np.ndarray, view):
    """src/vision/find_minimap.rs:51-146 on frame.view(x, y, w, h) -> None | dict(left, right, top, bottom)."""
    bgra = np.ascontiguousarray(bgra)
    H, W = bgra.shape[:2]
    x, y, w, h = view
    rect = (C.c_uint32 * 4)()
    ok = lib().smh_oracle_find_minimap_in(bgra.ctypes.data, W, x, y, w, h, rect)
    return dict(zip(("left", "right", "top", "bottom"), list(rect))) if ok else None


def find_lines_in(img: np.ndarray, max_gap: int) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    lines = np.zeros((32, 4), np.float32)
    n = C.c_uint32()
    lib().smh_oracle_find_lines_in(img.ctypes.data, w, h, max_gap, lines.ctypes.data, C.byref(n))
    return lines[: n.value].copy()


def find_longest_line_in(img: np.ndarray, x: float, y: float, max_gap: float):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    line = (C.c_float * 4)()
    ln = C.c_float()
    lib().smh_oracle_find_longest_line_in(img.ctypes.data, w, h, x, y, max_gap, line, C.byref(ln))
    return np.array(list(line), np.float32), np.float32(ln.value)


def ray_lengths_in(img: np.ndarray, x: float, y: float, max_gap: float) -> np.ndarray:
    """Squared length of each of the 3600 rays from (x, y) (test helper)."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.empty(3600, np.float32)
    lib().smh_oracle_ray_lengths_in(img.ctypes.data, w, h, x, y, max_gap, out.ctypes.data)
    return out


def get_centre_in(img: np.ndarray, x: float, y: float):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = (C.c_float * 2)()
    lib().smh_oracle_get_centre_in(img.ctypes.data, w, h, x, y, out)
    return float(out[0]), float(out[1])


# ---- vision-gpu flavour (restated reference CUDA kernels; see smh_oracle.h) -------------------------------

def gpuflavour_rays_in(img: np.ndarray, x: float, y: float, max_gap: float, dxdy: np.ndarray):
    """K10 restated: -> (ends[3600, 2] f32, d5[3600] u8: the gap_count rule decided the ray)."""
    img = np.ascontiguousarray(img, np.uint8)
    dxdy = np.ascontiguousarray(dxdy, np.float32)
    h, w = img.shape
    ends = np.empty((3600, 2), np.float32)
    d5 = np.empty(3600, np.uint8)
    lib().smh_oracle_gpuflavour_rays_in(img.ctypes.data, w, h, x, y, max_gap, dxdy.ctypes.data, ends.ctypes.data, d5.ctypes.data)
    return ends, d5


def rays_in_dirs(img: np.ndarray, x: float, y: float, max_gap: float, dxdy: np.ndarray) -> np.ndarray:
    """The vision-cpu march over a caller-supplied direction table -> ends[3600, 2]."""
    img = np.ascontiguousarray(img, np.uint8)
    dxdy = np.ascontiguousarray(dxdy, np.float32)
    h, w = img.shape
    ends = np.empty((3600, 2), np.float32)
    lib().smh_oracle_rays_in_dirs(img.ctypes.data, w, h, x, y, max_gap, dxdy.ctypes.data, ends.ctypes.data)
    return ends


def gpuflavour_dilate(raw: np.ndarray) -> np.ndarray:
    """NPP 2x2 dilation as vision-gpu calls it; the 1-px frame stays 0."""
    raw = np.ascontiguousarray(raw, np.uint8)
    h, w = raw.shape
    out = np.zeros((h, w), np.uint8)
    lib().smh_oracle_gpuflavour_dilate(raw.ctypes.data, w, h, out.ctypes.data)
    return out


def gpuflavour_icon_sad(map_rgb: np.ndarray, x: int, y: int, tpl_rgba: np.ndarray) -> int:
    map_rgb = np.ascontiguousarray(map_rgb, np.uint8)
    tpl_rgba = np.ascontiguousarray(tpl_rgba, np.uint8)
    return int(lib().smh_oracle_gpuflavour_icon_sad(map_rgb.ctypes.data, map_rgb.shape[1], x, y, tpl_rgba.ctypes.data, tpl_rgba.shape[0]))


def gpuflavour_icon_clear(map_rgb: np.ndarray, x: int, y: int, size: int) -> np.ndarray:
    out = np.ascontiguousarray(map_rgb, np.uint8).copy()
    lib().smh_oracle_gpuflavour_icon_clear(out.ctypes.data, out.shape[1], out.shape[0], x, y, size)
    return out


def gpuflavour_table_marker() -> np.ndarray:
    """K8's marker predicate (vision-gpu's HSV formulas) over all 2^24 colours."""
    out = np.empty(1 << 24, np.uint8)
    lib().smh_oracle_gpuflavour_table_marker(out.ctypes.data)
    return out


def synth_frames(W: int, H: int, first: int, n: int, seed: int = 0x534D4800_00000000, n_segments: int = 8, n_blips: int = 12,
                 n_labels: int = 2, map_open: bool = True, blip_band_pct: int = 100) -> np.ndarray:
    """The benchmarks' synthetic captures (same generator as the product's rasterisers, byte-identical) without
    loading the product library -> [n, H, W, 4] uint8 BGRA."""
    class P(C.Structure):
        _fields_ = [("seed", C.c_uint64)] + [(k, C.c_uint32) for k in ("W", "H", "n_segments", "n_blips", "n_labels", "map_open", "blip_band_pct", "reserved")]
    p = P(seed, W, H, n_segments, n_blips, n_labels, 1 if map_open else 0, blip_band_pct, 0)
    out = np.empty((n, H, W, 4), np.uint8)
    if lib().smh_oracle_synth_frames(C.byref(p), first, n, out.ctypes.data) != 0:
        raise ValueError("synth_frames: bad parameters")
    return out
