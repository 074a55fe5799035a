//! Raw bindings of include/smh_b200.h.  One `extern "C"` item per entry point; names, argument order and types
//! are exactly the header's (tests/test_capi_cpu.py keeps the header, the library's exports and the Python
//! mirror in step, and that this file binds every entry point the header declares — by name: no Rust toolchain in the
//! build image checks the types).
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_void};

#[repr(C)]
pub struct smh_b200 {
    _private: [u8; 0],
}
#[repr(C)]
pub struct smh_b200_multi {
    _private: [u8; 0],
}

pub const SMH_OK: i32 = 0;
pub const SMH_STAGE_COUNT: usize = 11;
pub const SMH_BATCH_GRAYSCALE: u32 = 1;
pub const SMH_BATCH_MINIMAP: u32 = 32;
pub const SMH_BATCH_UI_L8: u32 = 256;

/// `(meters, (x, y))` of one "NNNm" label — src/vision/mod.rs:183
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct smh_b200_scale {
    pub meters: u32,
    pub x: u32,
    pub y: u32,
}

/// VisionResults — src/vision/mod.rs:22-29
#[repr(C)]
pub struct smh_b200_vision_results {
    pub map_open: i32,
    pub xywh: [u32; 4],
    pub ui_map_rgba: *const u8,
    pub markers: [[f32; 4]; 32],
    pub n_markers: u32,
    pub has_ratio: i32,
    pub meters_to_px_ratio: f64,
    pub has_minimap: i32,
    pub minimap: [u32; 4],
    pub scales_debug: [[u32; 4]; 3],
    pub n_scales_debug: u32,
    pub reserved: u32,
}

pub type smh_b200_ocr_fn = Option<
    unsafe extern "C" fn(user: *mut c_void, ocr_l8: *const u8, w: u32, h: u32, scales: *mut smh_b200_scale, n_scales: *mut u32, scales_start_y: *mut u32) -> i32,
>;

/// Host-resident results of a batch, tightly packed (smh_b200_host_results)
#[repr(C)]
pub struct smh_b200_host_results {
    pub ui_map: *mut u8,
    pub ocr_out: *mut u8,
    pub scales: *mut u8,
    pub lsd_image: *mut u8,
    pub lines: *mut f32,
    pub n_lines: *mut u32,
    pub red_pixels: *mut u32,
    pub map_open: *mut u8,
    pub minimap: *mut u32,
    pub minimap_found: *mut u8,
    pub icon_found: *mut u8,
    pub icon_xy: *mut u32,
    pub icon_sad: *mut u32,
}

/// One output plane of a device batch (smh_b200_plane): rows `pitch` bytes apart, `pad` pixels of left padding
#[repr(C)]
#[derive(Clone, Copy)]
pub struct smh_b200_plane {
    pub ptr: *mut u8,
    pub frame_stride: u64,
    pub pitch: u32,
    pub pad: u32,
    pub width: u32,
    pub height: u32,
    pub bytes_per_pixel: u32,
    pub reserved: u32,
}

/// Device-resident results of smh_b200_process_batch (smh_b200_batch); owned by the handle
#[repr(C)]
pub struct smh_b200_batch {
    pub n_frames: u32,
    pub w: u32,
    pub h: u32,
    pub reserved: u32,
    pub roi: [u32; 12],
    pub ui_map: smh_b200_plane,
    pub lsd_image: smh_b200_plane,
    pub ocr_out: smh_b200_plane,
    pub scales: smh_b200_plane,
    pub lines: *mut f32,
    pub n_lines: *mut u32,
    pub red_pixels: *mut u32,
    pub map_open: *mut u8,
    pub lsd_stats: *mut u64,
    pub minimap: *mut u32,
    pub minimap_found: *mut u8,
    pub icon_found: *mut u8,
    pub icon_xy: *mut u32,
    pub icon_sad: *mut u32,
}

/// Parameters of the synthetic capture generator (smh_b200_synth_params; tests and benchmarks)
#[repr(C)]
#[derive(Clone, Copy)]
pub struct smh_b200_synth_params {
    pub seed: u64,
    pub w: u32,
    pub h: u32,
    pub n_segments: u32,
    pub n_blips: u32,
    pub n_labels: u32,
    pub map_open: u32,
    pub blip_band_pct: u32,
    pub reserved: u32,
}

extern "C" {
    // ---- lifecycle + per-frame trait methods ----
    pub fn smh_b200_create(device: i32, out: *mut *mut smh_b200) -> i32;
    pub fn smh_b200_destroy(h: *mut smh_b200);
    pub fn smh_b200_last_error(h: *mut smh_b200) -> *const c_char;
    pub fn smh_b200_bind_thread(h: *mut smh_b200) -> i32;
    pub fn smh_b200_roi(w: u32, h: u32, roi: *mut u32) -> i32;
    pub fn smh_b200_load_frame(h: *mut smh_b200, bgra: *const u8, w: u32, hgt: u32, pitch_bytes: u32) -> i32;
    pub fn smh_b200_crop_to_map(h: *mut smh_b200, grayscale: i32, map_open: *mut i32, xywh: *mut u32, ui_map_rgba: *mut *const u8) -> i32;
    pub fn smh_b200_ocr_preprocess(h: *mut smh_b200, out: *mut *const u8, len: *mut usize) -> i32;
    pub fn smh_b200_find_scales_preprocess(h: *mut smh_b200, scales_start_y: u32, out: *mut *const u8, wh: *mut u32) -> i32;
    pub fn smh_b200_isolate_map_markers(h: *mut smh_b200) -> i32;
    pub fn smh_b200_mask_marker_lines(h: *mut smh_b200) -> i32;
    pub fn smh_b200_load_map_markers(h: *mut smh_b200, rgba: *const u8, n_markers: u32, size: u32) -> i32;
    pub fn smh_b200_filter_map_marker_icons(h: *mut smh_b200, found: *mut i32, xy: *mut u32, sad: *mut u32) -> i32;
    pub fn smh_b200_find_longest_line(h: *mut smh_b200, x: f32, y: f32, max_gap: f32, line: *mut f32, len_sq: *mut f32) -> i32;
    pub fn smh_b200_find_marker_lines(h: *mut smh_b200, max_gap: u32, lines: *mut [f32; 4], n: *mut u32) -> i32;
    pub fn smh_b200_debug_view(h: *mut smh_b200, which: u8, rgba: *mut *const u8, wh: *mut u32) -> i32;
    pub fn smh_b200_lsd_image(h: *mut smh_b200, l8: *mut *const u8, wh: *mut u32) -> i32;
    pub fn smh_b200_red_pixels(h: *mut smh_b200, count: *mut u32) -> i32;
    pub fn smh_b200_stream_synchronize(h: *mut smh_b200, stream: *mut c_void) -> i32;
    // ---- the callers either side of the trait (N1-N3) ----
    pub fn smh_b200_find_minimap(h: *mut smh_b200, found: *mut i32, rect: *mut u32) -> i32;
    pub fn smh_b200_calc_meters_to_px_ratio(h: *mut smh_b200, scales: *const smh_b200_scale, n_scales: u32, found: *mut i32, ratio: *mut f64,
                                            debug_lines: *mut [u32; 4], n_debug: *mut u32) -> i32;
    pub fn smh_b200_process(h: *mut smh_b200, bgra: *const u8, w: u32, hgt: u32, pitch_bytes: u32, grayscale: i32, detect_markers: i32,
                            max_gap: u32, ocr: smh_b200_ocr_fn, user: *mut c_void, out: *mut smh_b200_vision_results) -> i32;
    // ---- device-resident batches, their planes, the launch counter; generators and tables of the tests ----
    pub fn smh_b200_process_batch(h: *mut smh_b200, d_frames: *const u8, n: u32, w: u32, hgt: u32, pitch_bytes: u32, frame_stride: u64,
                                  start_y: *const u32, max_gap: u32, flags: u32, stream: *mut c_void, out: *mut smh_b200_batch) -> i32;
    pub fn smh_b200_batch_meters_to_px_ratio(h: *mut smh_b200, scales: *const smh_b200_scale, n_scales: *const u8, n: u32, stream: *mut c_void,
                                             ratio: *mut f64, found: *mut u8) -> i32;
    pub fn smh_b200_download_plane(h: *mut smh_b200, p: *const smh_b200_plane, first: u32, count: u32, host_out: *mut u8, stream: *mut c_void) -> i32;
    pub fn smh_b200_launch_count(h: *mut smh_b200) -> u64;
    pub fn smh_b200_synth_frames_host(p: *const smh_b200_synth_params, first: u32, n: u32, out: *mut u8) -> i32;
    pub fn smh_b200_synth_frames_device(h: *mut smh_b200, p: *const smh_b200_synth_params, first: u32, n: u32, d_out: *mut u8, stream: *mut c_void) -> i32;
    pub fn smh_b200_colour_table(h: *mut smh_b200, which: u32, host_out: *mut u8) -> i32;
    // ---- Timeshares (vision-common/src/debug.rs:3-29) ----
    pub fn smh_b200_timeshares(h: *mut smh_b200, host_ms: *mut f32, device_ms: *mut f32) -> i32;
    pub fn smh_b200_stage_name(stage: u32) -> *const c_char;
    // ---- batches of host frames: one GPU, a ring of buffers, several GPUs ----
    pub fn smh_b200_process_host_batch(h: *mut smh_b200, frames: *const u8, n: u32, w: u32, hgt: u32, pitch_bytes: u32, frame_stride: u64,
                                       start_y: *const u32, max_gap: u32, flags: u32, chunk: u32, res: *const smh_b200_host_results) -> i32;
    pub fn smh_b200_process_host_stream(h: *mut smh_b200, frames: *const u8, ring_n: u32, n: u32, w: u32, hgt: u32, pitch_bytes: u32,
                                        frame_stride: u64, start_y: *const u32, max_gap: u32, flags: u32, chunk: u32,
                                        res: *const smh_b200_host_results) -> i32;
    pub fn smh_b200_multi_create(devices: *const i32, n_dev: u32, out: *mut *mut smh_b200_multi) -> i32;
    pub fn smh_b200_multi_destroy(m: *mut smh_b200_multi);
    pub fn smh_b200_multi_last_error(m: *mut smh_b200_multi) -> *const c_char;
    pub fn smh_b200_multi_device_count(m: *mut smh_b200_multi) -> u32;
    pub fn smh_b200_shard_range(n: u32, n_dev: u32, g: u32, first: *mut u32, count: *mut u32);
    pub fn smh_b200_multi_process_host_batch(m: *mut smh_b200_multi, frames: *const u8, n: u32, w: u32, hgt: u32, pitch_bytes: u32,
                                             frame_stride: u64, start_y: *const u32, max_gap: u32, flags: u32, chunk: u32,
                                             res: *const smh_b200_host_results) -> i32;
    pub fn smh_b200_multi_load_map_markers(m: *mut smh_b200_multi, rgba: *const u8, n_markers: u32, size: u32) -> i32;
    pub fn smh_b200_multi_last_times(m: *mut smh_b200_multi, seconds: *mut f64, n: u32) -> i32;
}
