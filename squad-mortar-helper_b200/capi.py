"""ctypes declarations of the C ABI in include/smh_b200.h — the same symbols the Rust shim binds
(INTEGRATION.md).  Loading fails loudly when libsmh_b200.so is missing: there is no fallback.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
import os

# SMH_B200_LIB lets tuning runs load an alternative build of the same library
LIB_PATH = Path(os.environ.get("SMH_B200_LIB", PKG / "libsmh_b200.so"))

SMH_OK, SMH_E_ARG, SMH_E_GEOMETRY, SMH_E_STATE, SMH_E_ALIGN, SMH_E_NOGPU, SMH_E_RANGE = 0, -1, -2, -3, -4, -5, -6
BATCH_GRAYSCALE, BATCH_LSD_STATS, BATCH_SKIP_LSD, BATCH_SKIP_PIXELS, BATCH_LSD_EXHAUSTIVE, BATCH_MINIMAP, BATCH_SERIAL, BATCH_FILTER_ICONS, BATCH_UI_L8 = 1, 2, 4, 8, 16, 32, 64, 128, 256

# every symbol include/smh_b200.h declares (tests assert the library exports all of them)
SYMBOLS = [
    "smh_b200_create", "smh_b200_destroy", "smh_b200_last_error", "smh_b200_bind_thread", "smh_b200_roi",
    "smh_b200_load_frame", "smh_b200_crop_to_map", "smh_b200_ocr_preprocess", "smh_b200_find_scales_preprocess",
    "smh_b200_isolate_map_markers", "smh_b200_mask_marker_lines", "smh_b200_find_longest_line",
    "smh_b200_find_marker_lines", "smh_b200_debug_view", "smh_b200_lsd_image", "smh_b200_red_pixels",
    "smh_b200_process_batch", "smh_b200_process_host_batch", "smh_b200_download_plane", "smh_b200_stream_synchronize", "smh_b200_launch_count",
    "smh_b200_synth_frames_host", "smh_b200_synth_frames_device", "smh_b200_colour_table",
    "smh_b200_find_minimap", "smh_b200_calc_meters_to_px_ratio", "smh_b200_process", "smh_b200_batch_meters_to_px_ratio",
    "smh_b200_load_map_markers", "smh_b200_filter_map_marker_icons",
    "smh_b200_timeshares", "smh_b200_stage_name", "smh_b200_shard_range", "smh_b200_process_host_stream",
    "smh_b200_multi_create", "smh_b200_multi_destroy", "smh_b200_multi_last_error", "smh_b200_multi_device_count",
    "smh_b200_multi_process_host_batch", "smh_b200_multi_load_map_markers", "smh_b200_multi_last_times",
]
STAGES = ["load_frame", "crop_to_map", "find_minimap", "ocr_preprocess", "ocr", "find_scales_preprocess", "calc_meters_to_px_ratio",
          "isolate_map_markers", "filter_map_marker_icons", "mask_marker_lines", "find_marker_lines"]


class Plane(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("frame_stride", C.c_uint64), ("pitch", C.c_uint32), ("pad", C.c_uint32),
                ("width", C.c_uint32), ("height", C.c_uint32), ("bytes_per_pixel", C.c_uint32),
                ("reserved", C.c_uint32)]


class Batch(C.Structure):
    _fields_ = [("n_frames", C.c_uint32), ("W", C.c_uint32), ("H", C.c_uint32), ("reserved", C.c_uint32),
                ("roi", C.c_uint32 * 12),
                ("ui_map", Plane), ("lsd_image", Plane), ("ocr_out", Plane), ("scales", Plane),
                ("lines", C.c_void_p), ("n_lines", C.c_void_p), ("red_pixels", C.c_void_p),
                ("map_open", C.c_void_p), ("lsd_stats", C.c_void_p),
                ("minimap", C.c_void_p), ("minimap_found", C.c_void_p),
                ("icon_found", C.c_void_p), ("icon_xy", C.c_void_p), ("icon_sad", C.c_void_p)]


class HostResults(C.Structure):
    _fields_ = [("ui_map", C.c_void_p), ("ocr_out", C.c_void_p), ("scales", C.c_void_p), ("lsd_image", C.c_void_p),
                ("lines", C.c_void_p), ("n_lines", C.c_void_p), ("red_pixels", C.c_void_p), ("map_open", C.c_void_p),
                ("minimap", C.c_void_p), ("minimap_found", C.c_void_p),
                ("icon_found", C.c_void_p), ("icon_xy", C.c_void_p), ("icon_sad", C.c_void_p)]


class Scale(C.Structure):
    """(meters, (x, y)) of one OCR'd scale label — src/vision/mod.rs:183"""
    _fields_ = [("meters", C.c_uint32), ("x", C.c_uint32), ("y", C.c_uint32)]


class VisionResults(C.Structure):
    """VisionResults — src/vision/mod.rs:22-29 (plus the map ROI and the scales_debug lines)"""
    _fields_ = [("map_open", C.c_int32), ("xywh", C.c_uint32 * 4), ("ui_map_rgba", C.c_void_p),
                ("markers", (C.c_float * 4) * 32), ("n_markers", C.c_uint32),
                ("has_ratio", C.c_int32), ("meters_to_px_ratio", C.c_double),
                ("has_minimap", C.c_int32), ("minimap", C.c_uint32 * 4),
                ("scales_debug", (C.c_uint32 * 4) * 3), ("n_scales_debug", C.c_uint32), ("reserved", C.c_uint32)]


# int32_t (*)(void *user, const uint8_t *ocr_l8, uint32_t w, uint32_t h, smh_b200_scale scales[3], uint32_t *n, uint32_t *start_y)
OCR_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(Scale), C.POINTER(C.c_uint32),
                     C.POINTER(C.c_uint32))


class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("W", C.c_uint32), ("H", C.c_uint32), ("n_segments", C.c_uint32),
                ("n_blips", C.c_uint32), ("n_labels", C.c_uint32), ("map_open", C.c_uint32),
                ("blip_band_pct", C.c_uint32), ("reserved", C.c_uint32)]


_LIB = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                           "This backend has no CPU or PyTorch fallback.")
    L = C.CDLL(str(LIB_PATH))
    vp, u32, i32, u64, f32 = C.c_void_p, C.c_uint32, C.c_int32, C.c_uint64, C.c_float
    P = C.POINTER
    sig = {
        "smh_b200_create": (i32, [i32, P(vp)]),
        "smh_b200_destroy": (None, [vp]),
        "smh_b200_last_error": (C.c_char_p, [vp]),
        "smh_b200_bind_thread": (i32, [vp]),
        "smh_b200_roi": (i32, [u32, u32, P(u32)]),
        "smh_b200_load_frame": (i32, [vp, vp, u32, u32, u32]),
        "smh_b200_crop_to_map": (i32, [vp, i32, P(i32), P(u32), P(vp)]),
        "smh_b200_ocr_preprocess": (i32, [vp, P(vp), P(C.c_size_t)]),
        "smh_b200_find_scales_preprocess": (i32, [vp, u32, P(vp), P(u32)]),
        "smh_b200_isolate_map_markers": (i32, [vp]),
        "smh_b200_mask_marker_lines": (i32, [vp]),
        "smh_b200_find_longest_line": (i32, [vp, f32, f32, f32, P(f32), P(f32)]),
        "smh_b200_find_marker_lines": (i32, [vp, u32, vp, P(u32)]),
        "smh_b200_debug_view": (i32, [vp, C.c_uint8, P(vp), P(u32)]),
        "smh_b200_lsd_image": (i32, [vp, P(vp), P(u32)]),
        "smh_b200_red_pixels": (i32, [vp, P(u32)]),
        "smh_b200_process_batch": (i32, [vp, vp, u32, u32, u32, u32, u64, vp, u32, u32, vp, P(Batch)]),
        "smh_b200_process_host_batch": (i32, [vp, vp, u32, u32, u32, u32, u64, vp, u32, u32, u32, P(HostResults)]),
        "smh_b200_download_plane": (i32, [vp, P(Plane), u32, u32, vp, vp]),
        "smh_b200_stream_synchronize": (i32, [vp, vp]),
        "smh_b200_launch_count": (u64, [vp]),
        "smh_b200_synth_frames_host": (i32, [P(SynthParams), u32, u32, vp]),
        "smh_b200_synth_frames_device": (i32, [vp, P(SynthParams), u32, u32, vp, vp]),
        "smh_b200_colour_table": (i32, [vp, u32, vp]),
        "smh_b200_find_minimap": (i32, [vp, P(i32), P(u32)]),
        "smh_b200_calc_meters_to_px_ratio": (i32, [vp, P(Scale), u32, P(i32), P(C.c_double), vp, P(u32)]),
        "smh_b200_process": (i32, [vp, vp, u32, u32, u32, i32, i32, u32, OCR_FN, vp, P(VisionResults)]),
        "smh_b200_batch_meters_to_px_ratio": (i32, [vp, vp, vp, u32, vp, vp, vp]),
        "smh_b200_load_map_markers": (i32, [vp, vp, u32, u32]),
        "smh_b200_filter_map_marker_icons": (i32, [vp, P(i32), P(u32), P(u32)]),
        "smh_b200_timeshares": (i32, [vp, P(f32), P(f32)]),
        "smh_b200_process_host_stream": (i32, [vp, vp, u32, u32, u32, u32, u32, u64, vp, u32, u32, u32, P(HostResults)]),
        "smh_b200_stage_name": (C.c_char_p, [u32]),
        "smh_b200_shard_range": (None, [u32, u32, u32, P(u32), P(u32)]),
        "smh_b200_multi_create": (i32, [P(i32), u32, P(vp)]),
        "smh_b200_multi_destroy": (None, [vp]),
        "smh_b200_multi_last_error": (C.c_char_p, [vp]),
        "smh_b200_multi_device_count": (u32, [vp]),
        "smh_b200_multi_process_host_batch": (i32, [vp, vp, u32, u32, u32, u32, u64, vp, u32, u32, u32, P(HostResults)]),
        "smh_b200_multi_load_map_markers": (i32, [vp, vp, u32, u32]),
        "smh_b200_multi_last_times": (i32, [vp, P(C.c_double), u32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


class VisionError(RuntimeError):
    """Mirrors the trait's `Self::Error` (anyhow::Error in vision-gpu): code + message."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"smh_b200 error {code}: {msg}")
        self.code = code


def check(handle, rc: int) -> None:
    if rc != 0:
        msg = lib().smh_b200_last_error(handle)
        raise VisionError(rc, msg.decode() if msg else "")
