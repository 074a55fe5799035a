//! Raw bindings of include/smh_b200.h (the per-frame part).  One `extern "C"` item per entry
//! point; names, argument order and types are exactly the header's.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

#[repr(C)]
pub struct smh_b200 {
    _private: [u8; 0],
}

pub const SMH_OK: i32 = 0;

extern "C" {
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
    #[allow(dead_code)]
    pub fn smh_b200_stream_synchronize(h: *mut smh_b200, stream: *mut c_void) -> i32;
}
