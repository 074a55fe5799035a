/* smh_oracle.h — CPU ORACLE for the map-analysis vision path.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's `vision-cpu` backend (CPUFallback) and the
 * shared `lsd::find_lines` driver.  It exists to CHECK the CUDA path; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product library (libsmh_b200.so) never links or calls it.
 *
 * PARITY STATUS: **pinned against outputs of the reference itself, run here** — its own CUDA backend.
 * The reference holds no golden vectors or known-answer tests for this path and its
 * Rust crates cannot be built in this image (no cargo/rustc; crates.io deps un-vendored), but its
 * `vision-gpu` kernels compile unmodified for sm_100a from where they lie (oracle/Makefile `refgpu`
 * -> oracle/_ref/libsmh_refgpu.so, driven by oracle/ref_gpu_harness.cu).  The `smh_oracle_gpuflavour_*`
 * functions below restate those KERNELS and tests/test_gpu_refcuda.py demands bit equality with their
 * object code on a B200: every ray of K10 (1826 seeds x 3600 rays x four max_gap values), K8 over all
 * 2^24 colours (tests/golden/k8_marker_table.npz is K8's own output), the NPP dilation, K1-K6, K11/K12.
 * The `vision-cpu` functions the CUDA product is checked against share their code with that flavour
 * wherever the two backends agree, with the documented divergences switched;
 * what stays reader-verified — branches only `vision-cpu` has, which no reference object code can
 * execute — is listed in DESIGN.md section 2.  Beside that: an independent numpy restatement
 * (oracle/np_oracle.py), exhaustive 2^24-colour tables, self-generated goldens under tests/golden/.
 *
 * Third-party arithmetic restated from the published algorithm (sources NOT in
 * /root/reference): image 0.23.14 (`Pixel::to_luma`, Cargo.lock:1487), imageproc 0.22.0
 * (`morphology::dilate_mut(Norm::L1, 1)`, Cargo.lock:1503), rayon 1.5.3 (reduce order),
 * libm cosf/sinf (glibc here).
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 */
#ifndef SMH_ORACLE_H
#define SMH_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct smh_oracle smh_oracle;

/* ---- pure helpers -------------------------------------------------------------- */

/* vision-common/src/screen.rs:55-66 + consts/mod.rs:7-19 + vision-cpu/src/lib.rs:116,137-145.
 * out = {btn x,y,w,h, map x,y,w,h, brq x,y,w,h} (brq in FRAME coordinates).
 * Returns 0, or -1 when the reference would panic (crop out of bounds, D12). */
int smh_oracle_roi(uint32_t W, uint32_t H, uint32_t out[12]);

/* image 0.23.14 rgb_to_luma: ((0.2126f*r + 0.7152f*g) + 0.0722f*b) as u8, unfused f32 */
uint8_t smh_oracle_luma8(uint8_t r, uint8_t g, uint8_t b);
/* util/src/image.rs:159-187 */
void smh_oracle_hsv(uint8_t r, uint8_t g, uint8_t b, uint16_t *h, uint8_t *s, uint8_t *v);
/* vision-common/src/markers/mod.rs:17-19,40-54 */
int smh_oracle_is_marker(uint8_t r, uint8_t g, uint8_t b);
/* Exhaustive tables over all 2^24 colours, index = r<<16 | g<<8 | b. */
void smh_oracle_table_marker(uint8_t *out /* 2^24 */);
void smh_oracle_table_luma(uint8_t *out /* 2^24 */);
void smh_oracle_table_hsv(uint16_t *h, uint8_t *s, uint8_t *v /* 2^24 each */);
/* the 3600 (cos, sin) pairs of vision-cpu/src/lib.rs:398-399,437 (f32, libm) */
void smh_oracle_ray_table(float *dxdy /* 7200: dx0,dy0,dx1,dy1,... */);

/* ---- the backend object (mirrors CPUFallback, vision-cpu/src/lib.rs:11-26) ------ */

smh_oracle *smh_oracle_new(void);
void smh_oracle_free(smh_oracle *o);
/* Threads used by the pixel loops and the 3600-ray search (rayon's role). 1 = scalar. */
void smh_oracle_set_threads(int n);
int smh_oracle_get_threads(void);

/* vision-cpu/src/lib.rs:69-93.  bgra tightly packed W*H*4 (B,G,R,A).  The oracle copies it. */
int smh_oracle_load_frame(smh_oracle *o, const uint8_t *bgra, uint32_t W, uint32_t H);
/* vision-cpu/src/lib.rs:110-171.  *map_open=0 mirrors Ok(None).  ui_map: borrowed, w*h*4. */
int smh_oracle_crop_to_map(smh_oracle *o, int grayscale, int *map_open, uint32_t xywh[4],
                           const uint8_t **ui_map_rgba, uint32_t *red_pixels);
/* vision-cpu/src/lib.rs:173-231 */
int smh_oracle_ocr_preprocess(smh_oracle *o, const uint8_t **out, size_t *len);
/* vision-cpu/src/lib.rs:233-251 (rows < start_y keep their previous content, D9) */
int smh_oracle_find_scales_preprocess(smh_oracle *o, uint32_t scales_start_y, const uint8_t **out,
                                      uint32_t wh[2]);
/* vision-cpu/src/lib.rs:253-280 */
int smh_oracle_isolate_map_markers(smh_oracle *o);
/* The template-match stage the reference carries commented out (A20; vision-cpu/src/lib.rs:101-108,
 * 270-277, 282-355).  Templates: n * size*size RGBA in the UnfilteredMarkers layout
 * (vision-common/src/markers/mod.rs:142-152).  Call between isolate_map_markers and mask_marker_lines
 * (src/vision/mod.rs:96-107): erases the best-matching icon from cropped_map and leaves the 4x4 green
 * square.  D17: ties go to the first candidate in raster order. */
int smh_oracle_load_map_markers(smh_oracle *o, const uint8_t *rgba, uint32_t n, uint32_t size);
int smh_oracle_filter_map_marker_icons(smh_oracle *o, int *found, uint32_t xy[2], uint32_t *sad);
/* vision-cpu/src/lib.rs:357-375 (+ imageproc dilate L1,k=1) */
int smh_oracle_mask_marker_lines(smh_oracle *o);
/* vision-cpu/src/lib.rs:387-449 on the oracle's own lsd_image */
int smh_oracle_find_longest_line(smh_oracle *o, float x, float y, float max_gap, float line[4],
                                 float *len_sq);
/* same search on a caller-supplied L8 image (the trait passes the image explicitly) */
int smh_oracle_find_longest_line_in(const uint8_t *img, uint32_t w, uint32_t h, float x, float y,
                                    float max_gap, float line[4], float *len_sq);
/* test helpers: squared length of each of the 3600 rays of one seed; get_centre on an image */
int smh_oracle_ray_lengths_in(const uint8_t *img, uint32_t w, uint32_t h, float x, float y, float max_gap,
                              float *len_sq);
void smh_oracle_get_centre_in(const uint8_t *img, uint32_t w, uint32_t h, float x, float y, float out[2]);
/* vision-cpu/src/lib.rs:377-385 + vision-common/src/lsd.rs:60-105 */
int smh_oracle_find_marker_lines(smh_oracle *o, uint32_t max_gap, float lines[32][4], uint32_t *n);
/* lsd::find_lines on a caller-supplied L8 image */
int smh_oracle_find_lines_in(const uint8_t *img, uint32_t w, uint32_t h, uint32_t max_gap,
                             float lines[32][4], uint32_t *n);
/* lsd::find_lines with the backend's find_longest_line supplied by the caller (the closure of
 * vision-common/src/lsd.rs:63; how vision-gpu plugs its per-seed kernel launch in,
 * vision-gpu/src/lib.rs:477-487).  A non-zero return of `cb` aborts the scan and is passed on. */
typedef int (*smh_oracle_longest_fn)(void *user, const uint8_t *img, uint32_t w, uint32_t h, float x, float y,
                                     float max_gap, float line[4], float *len_sq);
int smh_oracle_find_lines_cb(const uint8_t *img, uint32_t w, uint32_t h, uint32_t max_gap,
                             smh_oracle_longest_fn cb, void *user, float lines[32][4], uint32_t *n);
/* vision-cpu/src/lib.rs:451-460; which = DebugView discriminant (vision-common/src/debug.rs:31-40).
 * Fills rgba (w*h*4 of that view) if non-NULL; always writes wh. Returns 1 for None. */
int smh_oracle_debug_view(smh_oracle *o, uint8_t which, uint8_t *rgba, uint32_t wh[2]);

/* borrowed views of the intermediate buffers (valid until the next load_frame/free) */
const uint8_t *smh_oracle_cropped_map(smh_oracle *o, uint32_t wh[2]);  /* RGB8 */
const uint8_t *smh_oracle_cropped_brq(smh_oracle *o, uint32_t wh[2]);  /* RGB8 */
const uint8_t *smh_oracle_lsd_image(smh_oracle *o, uint32_t wh[2]);    /* L8 */
const uint8_t *smh_oracle_scales(smh_oracle *o, uint32_t wh[2]);       /* L8 */
/* raw (pre-dilation) marker mask of the last mask_marker_lines call, L8 0/255 */
const uint8_t *smh_oracle_raw_mask(smh_oracle *o, uint32_t wh[2]);

/* work counters of the last find_marker_lines / find_lines_in call on this thread */
void smh_oracle_lsd_stats(uint64_t *seeds, uint64_t *ray_steps);

/* calc_meters_to_px_ratio — src/vision/mpx_ratio.rs:3-133 on a caller-supplied `scales_preprocessed`
 * image.  scales[k] = {meters, x, y} (<= 3).  Returns 1 (Some: *ratio set) or 0 (None).
 * debug[k] = {meters, left, right, y} of each scale bar found (the scales_debug lines), may be NULL.
 * D16: where the reference reads outside the image (unchecked in release, panic in debug) the
 * pixel counts as non-black. */
int smh_oracle_calc_meters_to_px_ratio(const uint8_t *img, uint32_t w, uint32_t h, const uint32_t (*scales)[3],
                                       uint32_t n_scales, double *ratio, uint32_t debug[3][4], uint32_t *n_debug);
/* find_minimap — src/vision/find_minimap.rs:51-146 on the view (vx, vy, w, h) of a BGRA image with
 * `pitch_px` pixels per row.  rect = {left, right, top, bottom}.  Returns 1 (Some) or 0 (None). */
int smh_oracle_find_minimap_in(const uint8_t *bgra, uint32_t pitch_px, uint32_t vx, uint32_t vy, uint32_t w, uint32_t h,
                               uint32_t rect[4]);
/* same on the loaded frame's map ROI (src/vision/mod.rs:85) */
int smh_oracle_find_minimap(smh_oracle *o, uint32_t rect[4]);

/* Whole per-frame pipeline in the orchestrator's order (src/vision/mod.rs:66-223):
 * load_frame, crop_to_map, isolate, mask, find_marker_lines(max_gap), ocr_preprocess,
 * find_scales_preprocess(start_y).  Used by the CPU-baseline timing. Returns map_open. */
int smh_oracle_process_frame(smh_oracle *o, const uint8_t *bgra, uint32_t W, uint32_t H,
                             int grayscale, uint32_t start_y, uint32_t max_gap,
                             float lines[32][4], uint32_t *n_lines);


/* Synthetic captures of the benchmarks (the generator of SURVEY.md 8(d), squad-mortar-helper_b200/csrc/smh_synth.h, shared
 * with the product's rasterisers and byte-identical to them): frames [first, first + n) -> host BGRA, tightly packed.
 * params = {seed(u64), W, H, n_segments, n_blips, n_labels, map_open, blip_band_pct, reserved} (u32 each). */
int smh_oracle_synth_frames(const void *params, uint32_t first, uint32_t n, uint8_t *out);

/* ---- `vision-gpu` flavour: the reference's CUDA kernels restated on the CPU ------------------------
 * vision-cpu and vision-gpu are two implementations of one specification by one author that differ in
 * documented details.  vision-cpu cannot be built here (no Rust toolchain) but vision-gpu's
 * kernels can: these functions restate THEIR arithmetic, so that tests/test_gpu_refcuda.py can demand BIT
 * equality with the reference's own object code (oracle/_ref/libsmh_refgpu.so) — that pins the march loop,
 * the end-point rule, the saturating casts and the dilation model against the reference itself.  The
 * vision-cpu functions above share their code with these wherever the two backends agree (the same `march`
 * body with the D5 rule switched), so what remains reader-verified is listed explicitly in DESIGN.md §2. */

/* K7/K8's marker predicate: RGB::to_hsv + is_any_map_marker_color of vision-gpu (cuda.cu:191-239, 277-309) — differs
 * from vision-cpu's in how h and s are formed (D1).  Pinned bit for bit over all 2^24 colours against a table the
 * reference's own K8 object code produced on a B200 (tests/golden/k8_marker_table.npz, tools/k8_table.py). */
int smh_oracle_gpuflavour_is_marker(uint8_t r, uint8_t g, uint8_t b);
void smh_oracle_gpuflavour_table_marker(uint8_t *out /* 2^24 */);

/* One ray of K10 `find_longest_line` (vision-gpu/cuda/cuda.cu:637-722): directions are INPUTS (the kernel
 * computes them with CUDA's cosf/sinf on a double-precision theta, D11; the harness exports that table).
 * D5: besides `gap >= max_gap` the kernel stops on `gap_count >= max_gap` (cuda.cu:664,685,697).
 * end = the kernel's (x_end, y_end).  Returns bit 0: the ray stopped on the gap_count rule while the
 * vision-cpu rule would have marched on (D5 decided this ray). */
uint32_t smh_oracle_gpuflavour_ray(const uint8_t *img, uint32_t w, uint32_t h, float px, float py, float max_gap,
                                   float dx, float dy, float end[2]);
/* all 3600 rays of one seed: dxdy[7200] in, ends[7200] out, d5[3600] out (may be NULL) */
int smh_oracle_gpuflavour_rays_in(const uint8_t *img, uint32_t w, uint32_t h, float px, float py, float max_gap,
                                  const float *dxdy, float *ends, uint8_t *d5);
/* the vision-cpu march (vision-cpu/src/lib.rs:388-432 — the SAME code find_longest_line_in runs) over a
 * caller-supplied direction table: with the kernel's own table it must equal the kernel wherever d5 = 0 */
int smh_oracle_rays_in_dirs(const uint8_t *img, uint32_t w, uint32_t h, float px, float py, float max_gap,
                            const float *dxdy, float *ends);
/* K9: nppiDilate_8u_C1R as vision-gpu calls it (vision-gpu/cuda/dilate.cpp:3-25, vision-gpu/src/lib.rs:465-472):
 * 2x2 all-ones mask anchored (1,1) over the ROI (1,1)..(w-2,h-2); the outer 1-px frame of `out` is NOT
 * written (D4) — the caller pre-fills it. */
void smh_oracle_gpuflavour_dilate(const uint8_t *raw, uint32_t w, uint32_t h, uint8_t *out);
/* K11 `filter_map_marker_icons` for ONE (template, candidate) pair (cuda.cu:548-585): u16 accumulator, the
 * alpha scale evaluated in double (`(float)ad * ((float)a / 255.0)`), every template pixel visited.
 * map: RGB8 with `stride` pixels per row; (x, y): the candidate (window origin). */
uint16_t smh_oracle_gpuflavour_icon_sad(const uint8_t *map_rgb, uint32_t stride, uint32_t x, uint32_t y,
                                        const uint8_t *tpl_rgba, uint32_t size);
/* K12 `filter_map_marker_icons_clear` (cuda.cu:587-614) launched as vision-gpu would
 * (grid ceil(size/8)^2 x (8,8), vision-gpu/src/lib.rs:413-424): blacks the whole launched square (a multiple
 * of 8, not `size`), 5x5 green square at the tip. */
void smh_oracle_gpuflavour_icon_clear(uint8_t *map_rgb, uint32_t w, uint32_t h, uint32_t x, uint32_t y, uint32_t size);

#ifdef __cplusplus
}
#endif
#endif
