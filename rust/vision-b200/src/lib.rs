//! `B200Vision` — a third implementation of `vision_common::Vision`
//! (vision-common/src/lib.rs:30-61) next to `CPUFallback` (vision-cpu/src/lib.rs:55-461) and
//! `CudaInstance` (vision-gpu/src/lib.rs:147-560).  Every trait method is a thin call into the
//! `extern "C"` library of include/smh_b200.h; the host stays Rust, the device code is
//! hand-written CUDA C++ for sm_100a compiled by build.rs.  No `cust`, no PTX, no NPP, no CPU
//! fallback: `init()` returns `Err` on a machine without a Blackwell GPU and the app's existing
//! `hardware::gpu()` error path (src/vision/hardware.rs:66-82) takes over.
//!
//! A SKETCH, NOT A TESTED CRATE: it has never been compiled (no Rust toolchain in the build image), so the
//! crate / macro names taken from the reference (`smh_vision_common`, `export_dylib_wrapper!`) and the trait conformance
//! are unverified; compile it against the reference workspace before relying on it.  What IS tested is the C ABI
//! underneath, through the ctypes mirror squad-mortar-helper_b200/vision.py (same calls, same order).
mod ffi;

use std::{ffi::CStr, ptr, sync::Arc};

use smh_vision_common::{debug::DebugView, prelude::*, Vision};

// Exported with the reference's own macro so that src/vision/hardware.rs:60-64 can
// `VisionDylibWrapper::wrap(logger, "smh_vision_b200")` (same-rustc caveat).
smh_vision_common::export_dylib_wrapper!(smh_vision_b200 => B200Vision);

/// Opaque handle to the backend's own dilated marker mask on the device.  The trait passes it
/// back into `find_longest_line`; the library always searches the handle's current mask.
pub struct B200LsdImage(());

pub struct B200Vision {
    h: *mut ffi::smh_b200,
    cpu_frame: Arc<VisionFrame>,
    /// `find_scales_preprocess` hands out `*const SusRefCell<GrayImage>` (lib.rs:50); the cell
    /// is refreshed from the library's pinned host image on every call.
    scales_host: SusRefCell<image::GrayImage>,
    lsd_image: B200LsdImage,
}
// &self methods are called from two rayon workers (src/vision/mod.rs:219-223); the C library
// serialises each branch on its own mutex + stream (include/smh_b200.h "Threading").
unsafe impl Send for B200Vision {}
unsafe impl Sync for B200Vision {}

impl B200Vision {
    fn check(&self, rc: i32) -> Result<(), AnyError> {
        if rc == ffi::SMH_OK {
            return Ok(());
        }
        let msg = unsafe { CStr::from_ptr(ffi::smh_b200_last_error(self.h)) }.to_string_lossy().into_owned();
        Err(anyhow::anyhow!("smh_b200 error {rc}: {msg}"))
    }
}

impl B200Vision {
    /// `load_map_markers` — vision-gpu/src/lib.rs:129-137,202-204: the templates are prepared on the host by
    /// `markers::load_markers::<UnfilteredMarkers>` (vision-common/src/markers/mod.rs:111-119, once that
    /// module is un-commented) and handed to the library as one packed RGBA block.
    #[allow(dead_code)]
    pub fn load_map_markers_rgba(&mut self, rgba: &[u8], n_markers: u32, map_marker_size: u32) -> Result<(), AnyError> {
        assert_eq!(rgba.len(), (n_markers * map_marker_size * map_marker_size * 4) as usize);
        self.check(unsafe { ffi::smh_b200_load_map_markers(self.h, rgba.as_ptr(), n_markers, map_marker_size) })
    }

    /// `filter_map_marker_icons` — vision-cpu/src/lib.rs:282-355; `Ok(Some((x, y, sad)))` of the erased icon.
    #[allow(dead_code)]
    pub fn filter_map_marker_icons(&self) -> Result<Option<(u32, u32, u32)>, AnyError> {
        let (mut found, mut xy, mut sad) = (0i32, [0u32; 2], 0u32);
        self.check(unsafe { ffi::smh_b200_filter_map_marker_icons(self.h, &mut found, xy.as_mut_ptr(), &mut sad) })?;
        Ok(if found != 0 { Some((xy[0], xy[1], sad)) } else { None })
    }
}

impl B200Vision {
    /// `find_minimap` on the device copy of the loaded frame — src/vision/find_minimap.rs:51-146 (called at mod.rs:85).
    #[allow(dead_code)]
    pub fn find_minimap(&self) -> Result<Option<[u32; 4]>, AnyError> {
        let (mut found, mut rect) = (0i32, [0u32; 4]);
        self.check(unsafe { ffi::smh_b200_find_minimap(self.h, &mut found, rect.as_mut_ptr()) })?;
        Ok(if found != 0 { Some(rect) } else { None }) // left, right, top, bottom
    }

    /// `calc_meters_to_px_ratio` on the image `find_scales_preprocess` last returned — src/vision/mpx_ratio.rs:3-133.
    #[allow(dead_code)]
    pub fn calc_meters_to_px_ratio(&self, scales: &[(u32, (u32, u32))]) -> Result<Option<f64>, AnyError> {
        let sc: Vec<ffi::smh_b200_scale> = scales.iter().take(3).map(|&(meters, (x, y))| ffi::smh_b200_scale { meters, x, y }).collect();
        let (mut found, mut ratio, mut n_debug) = (0i32, 0f64, 0u32);
        let mut debug = [[0u32; 4]; 3];
        self.check(unsafe {
            ffi::smh_b200_calc_meters_to_px_ratio(self.h, sc.as_ptr(), sc.len() as u32, &mut found, &mut ratio, debug.as_mut_ptr(), &mut n_debug)
        })?;
        Ok(if found != 0 { Some(ratio) } else { None })
    }

    /// The reference's `Timeshares` (vision-common/src/debug.rs:3-29) of the last frame: per stage (wall-clock ms of
    /// the call, CUDA-event ms of the device work it enqueued), in the reference's stage order.
    #[allow(dead_code)]
    pub fn timeshares(&self) -> Result<[(f32, f32); ffi::SMH_STAGE_COUNT], AnyError> {
        let (mut host, mut dev) = ([0f32; ffi::SMH_STAGE_COUNT], [0f32; ffi::SMH_STAGE_COUNT]);
        self.check(unsafe { ffi::smh_b200_timeshares(self.h, host.as_mut_ptr(), dev.as_mut_ptr()) })?;
        let mut out = [(0f32, 0f32); ffi::SMH_STAGE_COUNT];
        for i in 0..ffi::SMH_STAGE_COUNT {
            out[i] = (host[i], dev[i]);
        }
        Ok(out)
    }
}

impl Drop for B200Vision {
    fn drop(&mut self) {
        unsafe { ffi::smh_b200_destroy(self.h) }
    }
}

impl Vision for B200Vision {
    type LSDImage = B200LsdImage;
    type Error = AnyError;

    fn init() -> Result<Self, AnyError> {
        let mut h = ptr::null_mut();
        let rc = unsafe { ffi::smh_b200_create(0, &mut h) };
        if rc != ffi::SMH_OK {
            let msg = unsafe { CStr::from_ptr(ffi::smh_b200_last_error(ptr::null_mut())) }.to_string_lossy().into_owned();
            return Err(anyhow::anyhow!("smh_b200_create failed ({rc}): {msg}"));
        }
        Ok(Self {
            h,
            cpu_frame: Arc::new(VisionFrame::from_raw(0, 0, Vec::new().into_boxed_slice()).unwrap()),
            scales_host: SusRefCell::new(image::GrayImage::new(0, 0)),
            lsd_image: B200LsdImage(()),
        })
    }

    fn thread_ctx(&self) -> Result<(), AnyError> {
        self.check(unsafe { ffi::smh_b200_bind_thread(self.h) })
    }

    fn get_cpu_frame(&self) -> Arc<VisionFrame> {
        self.cpu_frame.clone()
    }

    fn load_frame(&mut self, frame: VisionFrame) -> Result<(), Self::Error> {
        let (w, h) = frame.dimensions();
        // VisionFrame is tightly packed B,G,R,A (vision-common/src/lib.rs:17): pitch = 4*w
        self.check(unsafe { ffi::smh_b200_load_frame(self.h, frame.as_raw().as_ptr(), w, h, w * 4) })?;
        self.cpu_frame = Arc::new(frame);
        Ok(())
    }

    fn crop_to_map(&self, grayscale: bool) -> Result<Option<(image::RgbaImage, [u32; 4])>, Self::Error> {
        let (mut open, mut xywh, mut ui) = (0i32, [0u32; 4], ptr::null());
        self.check(unsafe { ffi::smh_b200_crop_to_map(self.h, grayscale as i32, &mut open, xywh.as_mut_ptr(), &mut ui) })?;
        if open == 0 {
            return Ok(None); // map closed — not an error (vision-cpu/src/lib.rs:131-134)
        }
        let [_, _, w, h] = xywh;
        // the trait returns an OWNED image: copy out of the library's pinned mirror
        let bytes = unsafe { std::slice::from_raw_parts(ui, w as usize * h as usize * 4) }.to_vec();
        Ok(Some((image::RgbaImage::from_raw(w, h, bytes).unwrap(), xywh)))
    }

    fn ocr_preprocess(&self) -> Result<(*const u8, usize), Self::Error> {
        let (mut p, mut len) = (ptr::null(), 0usize);
        self.check(unsafe { ffi::smh_b200_ocr_preprocess(self.h, &mut p, &mut len) })?;
        Ok((p, len)) // borrowed from the backend, like vision-gpu/src/lib.rs:316-320
    }

    fn find_scales_preprocess(&self, scales_start_y: u32) -> Result<*const SusRefCell<image::GrayImage>, Self::Error> {
        let (mut p, mut wh) = (ptr::null(), [0u32; 2]);
        self.check(unsafe { ffi::smh_b200_find_scales_preprocess(self.h, scales_start_y, &mut p, wh.as_mut_ptr()) })?;
        let bytes = unsafe { std::slice::from_raw_parts(p, wh[0] as usize * wh[1] as usize) };
        let mut img = self.scales_host.borrow_mut();
        if img.dimensions() != (wh[0], wh[1]) {
            *img = image::GrayImage::new(wh[0], wh[1]);
        }
        // the library's image already carries the stale rows above start_y (D9)
        img.copy_from_slice(bytes);
        Ok(&self.scales_host)
    }

    fn isolate_map_markers(&self) -> Result<(), Self::Error> {
        self.check(unsafe { ffi::smh_b200_isolate_map_markers(self.h) })
    }

    fn mask_marker_lines(&self) -> Result<(), Self::Error> {
        self.check(unsafe { ffi::smh_b200_mask_marker_lines(self.h) })
    }

    // The two trait methods the reference carries commented out (vision-common/src/lib.rs:46,53) exist as the inherent
    // methods `load_map_markers_rgba` / `filter_map_marker_icons` above; with the trait lines restored they would forward
    // to those (the TGA decode / resize of `load_markers` stays host-side `image` work, markers/mod.rs:96-119).

    fn find_longest_line(&self, _image: &Self::LSDImage, pt: Point<f32>, max_gap: f32) -> Result<(Line<f32>, f32), Self::Error> {
        let (mut line, mut len) = ([0f32; 4], 0f32);
        self.check(unsafe { ffi::smh_b200_find_longest_line(self.h, pt.x, pt.y, max_gap, line.as_mut_ptr(), &mut len) })?;
        Ok((Line::new(Point::new(line[0], line[1]), Point::new(line[2], line[3])), len))
    }

    fn find_marker_lines(&self, max_gap: u32) -> Result<SmallVec<Line<f32>, 32>, Self::Error> {
        // The reference drives `lsd::find_lines` from the host with one backend call per seed
        // (vision-gpu/src/lib.rs:477-487); here the whole scan is one kernel launch.
        let (mut lines, mut n) = ([[0f32; 4]; 32], 0u32);
        self.check(unsafe { ffi::smh_b200_find_marker_lines(self.h, max_gap, lines.as_mut_ptr(), &mut n) })?;
        let mut out = SmallVec::new();
        for l in &lines[..n as usize] {
            out.push(Line::new(Point::new(l[0], l[1]), Point::new(l[2], l[3])));
        }
        let _ = &self.lsd_image;
        Ok(out)
    }

    fn get_debug_view(&self, choice: DebugView) -> Option<Arc<image::RgbaImage>> {
        // discriminants follow vision-common/src/debug.rs:31-40 (None = 0, OCRInput = 1, ...)
        let (mut p, mut wh) = (ptr::null(), [0u32; 2]);
        let rc = unsafe { ffi::smh_b200_debug_view(self.h, choice as u8, &mut p, wh.as_mut_ptr()) };
        if rc != ffi::SMH_OK || p.is_null() {
            return None;
        }
        let bytes = unsafe { std::slice::from_raw_parts(p, wh[0] as usize * wh[1] as usize * 4) }.to_vec();
        image::RgbaImage::from_raw(wh[0], wh[1], bytes).map(Arc::new)
    }
}
