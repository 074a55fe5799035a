"""ctypes binding of oracle/_ref/libsmh_refgpu.so — the REFERENCE's own CUDA kernels behind a
restatement of vision-gpu's launch sequence (oracle/ref_gpu_harness.cu).  TEST INFRASTRUCTURE ONLY.

The library is built from the reference sources where they lie (`make -C oracle refgpu`, only where
/root/reference is mounted) and travels to the GPU box as a prebuilt file; `available()` says
whether it is there.  Used to pin the CPU oracle against outputs of the reference itself and as the
"GPU kernel to beat" timing of bench.py's baseline leg.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
SO = _HERE / "_ref" / "libsmh_refgpu.so"
_LIB = None


def available() -> bool:
    return SO.exists()


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        from . import oracle_py
        oracle_py.lib()  # the harness links the oracle's host-side find_lines
        L = C.CDLL(str(SO))
        vp = C.c_void_p
        L.refgpu_last_error.restype = C.c_char_p
        L.refgpu_create.argtypes = [C.POINTER(vp)]
        L.refgpu_destroy.argtypes = [vp]
        L.refgpu_process_frame.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32,
                                           C.POINTER(C.c_int), C.POINTER(C.c_uint32)] + [vp] * 7 + \
                                          [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
        L.refgpu_find_longest_line_in.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_float,
                                                  C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.refgpu_time_frames.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.refgpu_kernel_times.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        L.refgpu_ray_dirs.argtypes = [vp, vp]
        L.refgpu_rays_per_seed.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, C.c_uint32, C.c_float, vp]
        L.refgpu_block_winners.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, C.c_uint32, C.c_float, vp]
        L.refgpu_mask_and_dilate.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, vp]
        L.refgpu_icons.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, C.c_uint32, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32, vp]
        _LIB = L
    return _LIB


class RefGpu:
    """The reference's GPU backend (its kernels, its launch order) for one frame at a time."""

    def __init__(self):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.refgpu_create(C.byref(h))
        if rc:
            raise RuntimeError(f"refgpu_create rc={rc}: {self.L.refgpu_last_error().decode()}")
        self.h = h

    def close(self):
        if self.h:
            self.L.refgpu_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def process_frame(self, bgra: np.ndarray, grayscale=True, start_y=0, max_gap=15) -> dict:
        from . import oracle_py as O
        bgra = np.ascontiguousarray(bgra)
        H, W = bgra.shape[:2]
        roi = O.roi(W, H)
        _, _, w, h = roi["map"]
        _, _, qw, qh = roi["brq"]
        out = {"ui_map": np.zeros((h, w, 4), np.uint8), "cropped_map_isolated": np.zeros((h, w, 3), np.uint8),
               "cropped_brq": np.zeros((qh, qw, 3), np.uint8), "raw_mask": np.zeros((h, w), np.uint8),
               "lsd_image": np.zeros((h, w), np.uint8), "ocr_out": np.zeros((qh, qw), np.uint8),
               "scales": np.zeros((qh, qw), np.uint8)}
        lines = np.zeros((32, 4), np.float32)
        mo, red, n, seeds = C.c_int(0), C.c_uint32(0), C.c_uint32(0), C.c_uint64(0)
        rc = self.L.refgpu_process_frame(self.h, bgra.ctypes.data, W, H, int(grayscale), start_y, max_gap, C.byref(mo),
                                         C.byref(red), *[out[k].ctypes.data for k in
                                                         ("ui_map", "cropped_map_isolated", "cropped_brq", "raw_mask",
                                                          "lsd_image", "ocr_out", "scales")],
                                         lines.ctypes.data, C.byref(n), C.byref(seeds))
        if rc:
            raise RuntimeError(f"refgpu_process_frame rc={rc}: {self.L.refgpu_last_error().decode()}")
        out.update(map_open=bool(mo.value), red_pixels=red.value, lines=lines[: n.value].copy(), seeds=seeds.value, roi=roi)
        return out

    def find_longest_line_in(self, img: np.ndarray, x: float, y: float, max_gap: float):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        line = (C.c_float * 4)()
        ln = C.c_float()
        rc = self.L.refgpu_find_longest_line_in(self.h, img.ctypes.data, w, h, x, y, max_gap, line, C.byref(ln))
        if rc:
            raise RuntimeError(f"refgpu_find_longest_line_in rc={rc}: {self.L.refgpu_last_error().decode()}")
        return np.array(list(line), np.float32), np.float32(ln.value)

    def time_frames(self, frames: np.ndarray, grayscale=True, max_gap=15, warm=8, passes=5) -> dict:
        """frames: [n, H, W, 4] host BGRA.  Wall-clock frames/s of the reference's per-frame sequence: `warm` untimed
        frames, then `passes` timed passes over all frames (median reported, all listed)."""
        frames = np.ascontiguousarray(frames)
        n, H, W = frames.shape[:3]
        fps = (C.c_double * passes)()
        seeds, launches = C.c_uint64(0), C.c_uint64(0)
        rc = self.L.refgpu_time_frames(self.h, frames.ctypes.data, n, W, H, int(grayscale), max_gap, warm, passes, fps,
                                       C.byref(seeds), C.byref(launches))
        if rc:
            raise RuntimeError(f"refgpu_time_frames rc={rc}: {self.L.refgpu_last_error().decode()}")
        vals = sorted(fps)
        return {"frames_per_s": vals[len(vals) // 2], "frames_per_s_passes": list(fps), "seeds_per_frame": seeds.value / (n * passes),
                "launches_per_frame": launches.value / (n * passes)}

    def kernel_times(self, frames: np.ndarray, grayscale=True, max_gap=15) -> dict:
        """Device time of every reference kernel (CUDA events around each launch) per frame, in microseconds."""
        frames = np.ascontiguousarray(frames)
        n, H, W = frames.shape[:3]
        us = (C.c_double * 11)()
        cnt = (C.c_uint64 * 11)()
        rc = self.L.refgpu_kernel_times(self.h, frames.ctypes.data, n, W, H, int(grayscale), max_gap, us, cnt)
        if rc:
            raise RuntimeError(f"refgpu_kernel_times rc={rc}: {self.L.refgpu_last_error().decode()}")
        names = ["upload (pageable H2D, 14.7 MB)", "K1 count_close_deployment_button_red_pixels", "K2 crop_to_map", "K3 crop_to_map_grayscale",
                 "K4 crop_to_map_brq", "K5 ocr_preprocess", "K6 find_scales_preprocess", "K7 isolate_map_markers", "K8 mask_marker_lines",
                 "K9 nppiDilate_8u_C1R", "K10 find_longest_line (all seeds of the frame)"]
        return {nm: {"us_per_frame": us[i] / n, "launches_per_frame": cnt[i] / n} for i, nm in enumerate(names) if cnt[i]}

    def _ck(self, rc, what):
        if rc:
            raise RuntimeError(f"{what} rc={rc}: {self.L.refgpu_last_error().decode()}")

    def ray_dirs(self) -> np.ndarray:
        """The 3600 (dx, dy) of K10, computed on the device with the kernel's own expression."""
        out = np.empty((3600, 2), np.float32)
        self._ck(self.L.refgpu_ray_dirs(self.h, out.ctypes.data), "refgpu_ray_dirs")
        return out

    def rays_per_seed(self, img: np.ndarray, seeds: np.ndarray, max_gap: float) -> np.ndarray:
        """K10 as <<<3600, 1>>>: every ray's Line of every seed -> [n, 3600, 4]."""
        img = np.ascontiguousarray(img, np.uint8)
        seeds = np.ascontiguousarray(seeds, np.float32).reshape(-1, 2)
        h, w = img.shape
        out = np.empty((len(seeds), 3600, 4), np.float32)
        self._ck(self.L.refgpu_rays_per_seed(self.h, img.ctypes.data, w, h, seeds.ctypes.data, len(seeds), max_gap, out.ctypes.data),
                 "refgpu_rays_per_seed")
        return out

    def block_winners(self, img: np.ndarray, seeds: np.ndarray, max_gap: float) -> np.ndarray:
        """K10 in vision-gpu's launch shape <<<8, 450>>>: the 8 block winners of every seed -> [n, 8, 4]."""
        img = np.ascontiguousarray(img, np.uint8)
        seeds = np.ascontiguousarray(seeds, np.float32).reshape(-1, 2)
        h, w = img.shape
        out = np.empty((len(seeds), 8, 4), np.float32)
        self._ck(self.L.refgpu_block_winners(self.h, img.ctypes.data, w, h, seeds.ctypes.data, len(seeds), max_gap, out.ctypes.data),
                 "refgpu_block_winners")
        return out

    def mask_and_dilate(self, rgb: np.ndarray):
        """K8 + NPP dilate on an RGB crop -> (raw_mask, lsd_image)."""
        rgb = np.ascontiguousarray(rgb, np.uint8)
        h, w = rgb.shape[:2]
        raw = np.empty((h, w), np.uint8); out = np.empty((h, w), np.uint8)
        self._ck(self.L.refgpu_mask_and_dilate(self.h, rgb.ctypes.data, w, h, raw.ctypes.data, out.ctypes.data), "refgpu_mask_and_dilate")
        return raw, out

    def icons(self, map_rgb: np.ndarray, templates: np.ndarray, cand_xy: np.ndarray, clear_xy=None):
        """K11 per (template, candidate) and K12: -> (sads[n_markers, n_cand] u16, crop after the clear or None)."""
        map_rgb = np.ascontiguousarray(map_rgb, np.uint8)
        templates = np.ascontiguousarray(templates, np.uint8)
        cand_xy = np.ascontiguousarray(cand_xy, np.uint32)
        h, w = map_rgb.shape[:2]
        n, size = templates.shape[0], templates.shape[1]
        sads = np.zeros((n, len(cand_xy)), np.uint16)
        out = np.empty_like(map_rgb)
        self._ck(self.L.refgpu_icons(self.h, map_rgb.ctypes.data, w, h, templates.ctypes.data, n, size, cand_xy.ctypes.data, len(cand_xy),
                                     sads.ctypes.data, 0xFFFFFFFF if clear_xy is None else int(clear_xy), out.ctypes.data), "refgpu_icons")
        return sads, (out if clear_xy is not None else None)
