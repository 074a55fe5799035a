"""Host-side mirror of the reference's backend interface, `trait Vision`
(vision-common/src/lib.rs:30-61), over the C ABI of include/smh_b200.h.

The reference's host language is Rust, which this image cannot build; the Rust shim a
maintainer would add is in INTEGRATION.md.  This Python class keeps the trait's method names,
argument meaning, call-order contract and error behaviour so the parity tests read like the
reference's own smoke test (vision-gpu/src/lib.rs:562-622).  PyTorch is not involved here; the
product is the CUDA library.

    trait method                      here
    ------------------------------    -----------------------------------------------
    init() -> Result<Self>            B200Vision.init(device=0)
    thread_ctx(&self)                 thread_ctx()
    get_cpu_frame(&self) -> Arc<..>   get_cpu_frame()
    load_frame(&mut self, frame)      load_frame(bgra[H,W,4] uint8)
    crop_to_map(&self, grayscale)     crop_to_map(grayscale) -> None | (ui_map[h,w,4], [x,y,w,h])
    ocr_preprocess(&self)             ocr_preprocess() -> L8 [bqh,bqw] (borrowed view)
    find_scales_preprocess(&self, y)  find_scales_preprocess(y) -> L8 [bqh,bqw] (borrowed view)
    isolate_map_markers(&self)        isolate_map_markers()
    mask_marker_lines(&self)          mask_marker_lines()
    find_longest_line(&self, ..)      find_longest_line(pt, max_gap) -> (line[4], len_sq)
    find_marker_lines(&self, gap)     find_marker_lines(max_gap) -> lines[n,4] (n <= 32)
    get_debug_view(&self, choice)     get_debug_view(DebugView) -> None | rgba[h,w,4]
"""
from __future__ import annotations

import ctypes as C
import enum

import numpy as np

from . import capi
from .capi import VisionError, check  # noqa: F401


class DebugView(enum.IntEnum):
    """vision-common/src/debug.rs:31-40"""
    NONE = 0
    OCRInput = 1
    FindScalesInput = 2
    LSDPreprocess = 3
    LSDInput = 4
    CroppedBRQ = 5


def roi(W: int, H: int):
    """ROI rectangles of a W x H capture; None where the reference panics (D12). Host-only."""
    out = (C.c_uint32 * 12)()
    if capi.lib().smh_b200_roi(W, H, out) != 0:
        return None
    v = list(out)
    return {"button": tuple(v[0:4]), "map": tuple(v[4:8]), "brq": tuple(v[8:12])}


def _borrow(ptr: int, shape) -> np.ndarray:
    n = int(np.prod(shape))
    buf = (C.c_uint8 * n).from_address(ptr)
    a = np.frombuffer(buf, dtype=np.uint8).reshape(shape)
    a.flags.writeable = False
    return a


class B200Vision:
    def __init__(self, handle):
        self._L = capi.lib()
        self._h = handle
        self._frame = None

    # -- lifecycle ---------------------------------------------------------------------
    @classmethod
    def init(cls, device: int = 0) -> "B200Vision":
        L = capi.lib()
        h = C.c_void_p()
        rc = L.smh_b200_create(device, C.byref(h))
        if rc != 0:
            msg = L.smh_b200_last_error(None)
            raise VisionError(rc, msg.decode() if msg else "")
        return cls(h)

    def close(self) -> None:
        if self._h:
            self._L.smh_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self):
        return self._h

    def thread_ctx(self) -> None:
        check(self._h, self._L.smh_b200_bind_thread(self._h))

    # -- frame -------------------------------------------------------------------------
    def load_frame(self, bgra: np.ndarray) -> None:
        if bgra.dtype != np.uint8 or bgra.ndim != 3 or bgra.shape[2] != 4:
            raise ValueError("load_frame expects a [H, W, 4] uint8 BGRA image")
        if bgra.strides[2] != 1 or bgra.strides[1] != 4:
            bgra = np.ascontiguousarray(bgra)
        H, W = bgra.shape[:2]
        check(self._h, self._L.smh_b200_load_frame(self._h, bgra.ctypes.data, W, H, bgra.strides[0]))
        self._frame = bgra  # the shim keeps the Arc<VisionFrame> (get_cpu_frame)

    def get_cpu_frame(self) -> np.ndarray:
        return self._frame

    def crop_to_map(self, grayscale: bool):
        mo = C.c_int32(0)
        xywh = (C.c_uint32 * 4)()
        p = C.c_void_p()
        check(self._h, self._L.smh_b200_crop_to_map(self._h, int(bool(grayscale)), C.byref(mo), xywh, C.byref(p)))
        if not mo.value:
            return None
        x, y, w, h = list(xywh)
        # the trait returns an OWNED RgbaImage
        return _borrow(p.value, (h, w, 4)).copy(), [x, y, w, h]

    def red_pixels(self) -> int:
        c = C.c_uint32()
        check(self._h, self._L.smh_b200_red_pixels(self._h, C.byref(c)))
        return c.value

    # -- scales branch -------------------------------------------------------------------
    def ocr_preprocess(self) -> np.ndarray:
        p = C.c_void_p()
        n = C.c_size_t()
        check(self._h, self._L.smh_b200_ocr_preprocess(self._h, C.byref(p), C.byref(n)))
        wh = self._brq_wh()
        assert n.value == wh[0] * wh[1]
        return _borrow(p.value, (wh[1], wh[0]))

    def find_scales_preprocess(self, scales_start_y: int) -> np.ndarray:
        p = C.c_void_p()
        wh = (C.c_uint32 * 2)()
        check(self._h, self._L.smh_b200_find_scales_preprocess(self._h, scales_start_y, C.byref(p), wh))
        return _borrow(p.value, (wh[1], wh[0]))

    def _brq_wh(self):
        H, W = self._frame.shape[:2]
        r = roi(W, H)
        return r["brq"][2], r["brq"][3]

    # -- markers branch ------------------------------------------------------------------
    def isolate_map_markers(self) -> None:
        check(self._h, self._L.smh_b200_isolate_map_markers(self._h))

    def mask_marker_lines(self) -> None:
        check(self._h, self._L.smh_b200_mask_marker_lines(self._h))

    def load_map_markers(self, templates: np.ndarray) -> None:
        """load_map_markers(map_marker_size) — vision-common/src/lib.rs:46 (commented out in the reference).
        templates: [n, size, size, 4] RGBA as `markers::load_markers::<UnfilteredMarkers>(size)` returns them
        (vision-common/src/markers/mod.rs:111-119; host-side `image`-crate work)."""
        t = np.ascontiguousarray(templates, np.uint8)
        if t.ndim != 4 or t.shape[1] != t.shape[2] or t.shape[3] != 4:
            raise ValueError("load_map_markers expects [n, size, size, 4] uint8 RGBA templates")
        check(self._h, self._L.smh_b200_load_map_markers(self._h, t.ctypes.data, t.shape[0], t.shape[1]))

    def filter_map_marker_icons(self):
        """filter_map_marker_icons() — vision-common/src/lib.rs:53 (commented out in the reference), between
        isolate_map_markers and mask_marker_lines.  -> None | (x, y, sad) of the erased icon."""
        found = C.c_int32()
        xy = (C.c_uint32 * 2)()
        sad = C.c_uint32()
        check(self._h, self._L.smh_b200_filter_map_marker_icons(self._h, C.byref(found), xy, C.byref(sad)))
        return (int(xy[0]), int(xy[1]), int(sad.value)) if found.value else None

    def lsd_image(self) -> np.ndarray:
        p = C.c_void_p()
        wh = (C.c_uint32 * 2)()
        check(self._h, self._L.smh_b200_lsd_image(self._h, C.byref(p), wh))
        return _borrow(p.value, (wh[1], wh[0]))

    def find_longest_line(self, pt, max_gap: float):
        line = (C.c_float * 4)()
        ln = C.c_float()
        check(self._h, self._L.smh_b200_find_longest_line(self._h, float(pt[0]), float(pt[1]), float(max_gap), line,
                                                          C.byref(ln)))
        return np.array(list(line), np.float32), np.float32(ln.value)

    def find_marker_lines(self, max_gap: int) -> np.ndarray:
        lines = np.zeros((32, 4), np.float32)
        n = C.c_uint32()
        check(self._h, self._L.smh_b200_find_marker_lines(self._h, int(max_gap), lines.ctypes.data, C.byref(n)))
        return lines[: n.value].copy()

    def get_debug_view(self, choice: DebugView):
        p = C.c_void_p()
        wh = (C.c_uint32 * 2)()
        check(self._h, self._L.smh_b200_debug_view(self._h, int(choice), C.byref(p), wh))
        if not p.value:
            return None
        return _borrow(p.value, (wh[1], wh[0], 4)).copy()

    # -- the callers either side of the trait -------------------------
    def find_minimap(self):
        """find_minimap(vision.get_cpu_frame().view(x, y, w, h)) — src/vision/find_minimap.rs:51-146, on the
        device copy of the loaded frame.  -> None | dict(left, right, top, bottom)"""
        found = C.c_int32()
        rect = (C.c_uint32 * 4)()
        check(self._h, self._L.smh_b200_find_minimap(self._h, C.byref(found), rect))
        if not found.value:
            return None
        return dict(zip(("left", "right", "top", "bottom"), list(rect)))

    def calc_meters_to_px_ratio(self, scales, with_debug: bool = False):
        """calc_meters_to_px_ratio(scales, image) — src/vision/mpx_ratio.rs:3-133 — on the device copy of
        what find_scales_preprocess last returned.  scales: [(meters, (x, y)), ...] (<= 3).
        -> Option<f64> as None | float; with_debug also returns the scales_debug lines [(meters, left, right, y)]."""
        arr = (capi.Scale * 3)()
        for k, (m, (x, y)) in enumerate(scales):
            arr[k] = capi.Scale(int(m), int(x), int(y))
        found, ratio = C.c_int32(), C.c_double()
        dbg = ((C.c_uint32 * 4) * 3)()
        nd = C.c_uint32()
        check(self._h, self._L.smh_b200_calc_meters_to_px_ratio(self._h, arr, len(scales), C.byref(found), C.byref(ratio),
                                                                C.cast(dbg, C.c_void_p), C.byref(nd)))
        r = ratio.value if found.value else None
        if with_debug:
            return r, [tuple(dbg[k]) for k in range(nd.value)]
        return r

    def process(self, bgra: np.ndarray, *, grayscale: bool = True, detect_markers: bool = True, max_gap: int = 15, ocr=None):
        """VisionState::process — src/vision/mod.rs:36-240 — as one call into the C ABI.
        ocr: None (a heightmap is set: no scales branch) or callable(ocr_l8[h,w]) -> ([(meters, (x, y))...], scales_start_y)
        standing where Tesseract does.  -> None (map closed) | dict of VisionResults."""
        if bgra.strides[2] != 1 or bgra.strides[1] != 4:
            bgra = np.ascontiguousarray(bgra)
        H, W = bgra.shape[:2]
        err = []

        def _cb(user, img, w, h, scales, n, start_y):
            try:
                got, sy = ocr(_borrow(img, (h, w)))
                got = list(got)[:3]
                for k, (m, (x, y)) in enumerate(got):
                    scales[k].meters, scales[k].x, scales[k].y = int(m), int(x), int(y)
                n[0] = len(got)
                start_y[0] = int(sy) if got else 0xFFFFFFFF
                return 0
            except Exception as e:  # noqa: BLE001
                err.append(e)
                return -100

        cb = capi.OCR_FN(_cb) if ocr is not None else capi.OCR_FN()
        res = capi.VisionResults()
        rc = self._L.smh_b200_process(self._h, bgra.ctypes.data, W, H, bgra.strides[0], int(bool(grayscale)),
                                      int(bool(detect_markers)), int(max_gap), cb, None, C.byref(res))
        if err:
            raise err[0]
        check(self._h, rc)
        self._frame = bgra
        if not res.map_open:
            return None
        x, y, w, h = list(res.xywh)
        lines = np.array([list(res.markers[k]) for k in range(res.n_markers)], np.float32).reshape(-1, 4)
        return {"map": _borrow(res.ui_map_rgba, (h, w, 4)).copy(), "xywh": [x, y, w, h], "markers": lines,
                "meters_to_px_ratio": res.meters_to_px_ratio if res.has_ratio else None,
                "minimap_bounds": dict(zip(("left", "right", "top", "bottom"), list(res.minimap))) if res.has_minimap else None,
                "scales_debug": [tuple(res.scales_debug[k]) for k in range(res.n_scales_debug)]}

    def batch_meters_to_px_ratio(self, scales, stream: int = 0):
        """calc_meters_to_px_ratio for every frame of the last process_batch.  scales: per frame a list of
        (meters, (x, y)) (<= 3).  -> (ratio[n] float64, found[n] bool)"""
        n = len(scales)
        arr = np.zeros((n, 3, 3), np.uint32)
        cnt = np.zeros(n, np.uint8)
        for f, sc in enumerate(scales):
            cnt[f] = len(sc)
            for k, (m, (x, y)) in enumerate(sc):
                arr[f, k] = (m, x, y)
        ratio = np.zeros(n, np.float64)
        found = np.zeros(n, np.uint8)
        check(self._h, self._L.smh_b200_batch_meters_to_px_ratio(self._h, arr.ctypes.data, cnt.ctypes.data, n, stream,
                                                                 ratio.ctypes.data, found.ctypes.data))
        return ratio, found.astype(bool)

    # -- batch extension (not in the trait) ------------------------------------------------
    def process_batch(self, d_frames: int, n: int, W: int, H: int, *, pitch: int | None = None,
                      frame_stride: int | None = None, start_y=None, max_gap: int = 15, grayscale: bool = True,
                      lsd_stats: bool = False, skip_lsd: bool = False, skip_pixels: bool = False,
                      lsd_exhaustive: bool = False, minimap: bool = False, serial: bool = False,
                      filter_icons: bool = False,
                      stream: int = 0) -> capi.Batch:
        """Enqueue the whole path for n device-resident frames on `stream`; returns the plane table."""
        pitch = W * 4 if pitch is None else pitch
        frame_stride = pitch * H if frame_stride is None else frame_stride
        flags = (capi.BATCH_GRAYSCALE if grayscale else 0) | (capi.BATCH_LSD_STATS if lsd_stats else 0) | \
                (capi.BATCH_SKIP_LSD if skip_lsd else 0) | (capi.BATCH_SKIP_PIXELS if skip_pixels else 0) | \
                (capi.BATCH_LSD_EXHAUSTIVE if lsd_exhaustive else 0) | (capi.BATCH_MINIMAP if minimap else 0) | \
                (capi.BATCH_SERIAL if serial else 0) | (capi.BATCH_FILTER_ICONS if filter_icons else 0)
        sy = None
        if start_y is not None:
            sy = np.ascontiguousarray(start_y, np.uint32)
            assert sy.shape == (n,)
        out = capi.Batch()
        check(self._h, self._L.smh_b200_process_batch(self._h, d_frames, n, W, H, pitch, frame_stride,
                                                      sy.ctypes.data if sy is not None else None, max_gap, flags,
                                                      stream, C.byref(out)))
        return out

    def process_host_batch(self, frames: np.ndarray, *, start_y=None, max_gap: int = 15, grayscale: bool = True,
                           chunk: int = 32, want=("ui_map", "ocr_out", "scales"), out: dict | None = None,
                           frames_ptr: int | None = None, shape=None, ui_l8: bool = False, _multi=None,
                           ring: int = 0, n_total: int | None = None) -> dict:
        """End-to-end batch from host frames [n,H,W,4]; returns tightly packed host arrays.  ui_l8 (with grayscale):
        ui_map comes back as its L plane [n,h,w] instead of RGBA."""
        if frames is not None:
            n, H, W = frames.shape[:3]
            assert frames.dtype == np.uint8 and frames.flags["C_CONTIGUOUS"]
            frames_ptr = frames.ctypes.data
        else:
            n, H, W = shape
        r = roi(W, H)
        _, _, w, h = r["map"]
        _, _, qw, qh = r["brq"]
        if out is None:
            out = {}
        res = capi.HostResults()
        shapes = {"ui_map": (n, h, w) if ui_l8 else (n, h, w, 4), "ocr_out": (n, qh, qw), "scales": (n, qh, qw), "lsd_image": (n, h, w)}
        for k in ("ui_map", "ocr_out", "scales", "lsd_image"):
            if k in want:
                if k not in out:
                    out[k] = np.zeros(shapes[k], np.uint8)
                setattr(res, k, out[k].ctypes.data if isinstance(out[k], np.ndarray) else int(out[k]))
        for k, shp, dt in (("lines", (n, 32, 4), np.float32), ("n_lines", (n,), np.uint32),
                           ("red_pixels", (n,), np.uint32), ("map_open", (n,), np.uint8)):
            if k not in out:
                out[k] = np.zeros(shp, dt)
            setattr(res, k, out[k].ctypes.data if isinstance(out[k], np.ndarray) else int(out[k]))
        if "minimap" in want:
            out.setdefault("minimap", np.zeros((n, 4), np.uint32))
            out.setdefault("minimap_found", np.zeros((n,), np.uint8))
            res.minimap = out["minimap"].ctypes.data if isinstance(out["minimap"], np.ndarray) else int(out["minimap"])
            res.minimap_found = out["minimap_found"].ctypes.data if isinstance(out["minimap_found"], np.ndarray) else int(out["minimap_found"])
        if "icons" in want:  # the reference-disabled template match (needs load_map_markers)
            for k, shp, dt in (("icon_found", (n,), np.uint8), ("icon_xy", (n, 2), np.uint32), ("icon_sad", (n,), np.uint32)):
                out.setdefault(k, np.zeros(shp, dt))
                setattr(res, k, out[k].ctypes.data if isinstance(out[k], np.ndarray) else int(out[k]))
        sy = None
        if start_y is not None:
            sy = np.ascontiguousarray(start_y, np.uint32)
        flags = (capi.BATCH_GRAYSCALE if grayscale else 0) | (capi.BATCH_MINIMAP if "minimap" in want else 0) | \
                (capi.BATCH_FILTER_ICONS if "icons" in want else 0) | (capi.BATCH_UI_L8 if ui_l8 else 0)
        if _multi is not None:  # the same call spread over several GPUs (B200VisionMulti)
            rc = self._L.smh_b200_multi_process_host_batch(_multi, frames_ptr, n, W, H, W * 4, W * H * 4,
                                                           sy.ctypes.data if sy is not None else None, max_gap, flags, chunk, C.byref(res))
            if rc != 0:
                raise capi.VisionError(rc, (self._L.smh_b200_multi_last_error(_multi) or b"").decode())
            return out
        if ring:  # streaming over a ring of `n` host buffers: n_total frames, results at index f % n
            check(self._h, self._L.smh_b200_process_host_stream(self._h, frames_ptr, n, n_total or n, W, H, W * 4, W * H * 4,
                                                                sy.ctypes.data if sy is not None else None, max_gap, flags,
                                                                chunk, C.byref(res)))
            return out
        check(self._h, self._L.smh_b200_process_host_batch(self._h, frames_ptr, n, W, H, W * 4, W * H * 4,
                                                           sy.ctypes.data if sy is not None else None, max_gap, flags,
                                                           chunk, C.byref(res)))
        return out

    def timeshares(self) -> dict:
        """The reference's Timeshares (vision-common/src/debug.rs:3-29) of the last frame through the trait calls:
        stage -> (host_ms, device_ms)."""
        hm = (C.c_float * len(capi.STAGES))()
        dm = (C.c_float * len(capi.STAGES))()
        check(self._h, self._L.smh_b200_timeshares(self._h, hm, dm))
        return {name: (float(hm[i]), float(dm[i])) for i, name in enumerate(capi.STAGES)}

    def download_plane(self, plane: capi.Plane, first: int, count: int, stream: int = 0) -> np.ndarray:
        """Copy frames [first, first+count) of a device plane into a packed host array (synchronises)."""
        bpp = plane.bytes_per_pixel
        shape = (count, plane.height, plane.width) + ((bpp,) if bpp > 1 else ())
        out = np.empty(shape, np.uint8)
        check(self._h, self._L.smh_b200_download_plane(self._h, C.byref(plane), first, count, out.ctypes.data, stream))
        self.synchronize(stream)
        return out

    def synchronize(self, stream: int = 0) -> None:
        check(self._h, self._L.smh_b200_stream_synchronize(self._h, stream))

    def launch_count(self) -> int:
        return int(self._L.smh_b200_launch_count(self._h))

    def colour_table(self, which: int) -> np.ndarray:
        out = np.empty(1 << 24, np.uint8)
        check(self._h, self._L.smh_b200_colour_table(self._h, which, out.ctypes.data))
        return out


class B200VisionMulti:
    """Several GPUs of one box behind one call (smh_b200_multi_*): a batch of host frames is split into contiguous
    frame ranges, one host thread + handle per GPU, no collective."""

    def __init__(self, devices):
        self._L = capi.lib()
        devs = (C.c_int32 * len(devices))(*devices)
        h = C.c_void_p()
        rc = self._L.smh_b200_multi_create(devs, len(devices), C.byref(h))
        if rc != 0:
            raise capi.VisionError(rc, (self._L.smh_b200_last_error(None) or b"").decode())
        self._m = h
        self.devices = list(devices)

    def close(self):
        if self._m:
            self._L.smh_b200_multi_destroy(self._m)
            self._m = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def shard_range(self, n: int, g: int):
        a, b = C.c_uint32(), C.c_uint32()
        self._L.smh_b200_shard_range(n, len(self.devices), g, C.byref(a), C.byref(b))
        return a.value, b.value

    def load_map_markers(self, templates: np.ndarray) -> None:
        t = np.ascontiguousarray(templates, np.uint8)
        rc = self._L.smh_b200_multi_load_map_markers(self._m, t.ctypes.data, t.shape[0], t.shape[1])
        if rc != 0:
            raise capi.VisionError(rc, (self._L.smh_b200_multi_last_error(self._m) or b"").decode())

    def process_host_batch(self, frames, **kw) -> dict:
        # argument handling is B200Vision.process_host_batch's; only the final call differs
        return B200Vision.process_host_batch(_MultiShim(self._L), frames, _multi=self._m, **kw)

    def last_times(self):
        out = (C.c_double * len(self.devices))()
        self._L.smh_b200_multi_last_times(self._m, out, len(self.devices))
        return list(out)


class _MultiShim:
    def __init__(self, L):
        self._L = L
        self._h = None
