/* smh_b200.h — C ABI of the B200 (sm_100a) backend for squad-mortar-helper's per-frame
 * map-analysis vision path.
 *
 * This is the drop-in boundary: a Rust struct implementing `vision_common::Vision`
 * (reference: vision-common/src/lib.rs:30-61) forwards each trait method to the entry
 * point of the same name below (binding shown in INTEGRATION.md).  Plain pointers and
 * sizes only; every function returns 0 on success, a positive cudaError_t, or a negative
 * SMH_E_* code; smh_b200_last_error() gives the text.  "Map closed" is NOT an error
 * (reference: Ok(None), vision-cpu/src/lib.rs:131-134): it is reported through *map_open.
 *
 * Results are bit-identical to the reference's `vision-cpu` backend (the deterministic
 * one) — see DESIGN.md for the handful of defined deviations (D8, D9).
 *
 * Threading (reference: src/vision/mod.rs:219-223): after crop_to_map, the marker branch
 * {isolate_map_markers, mask_marker_lines, find_marker_lines, find_longest_line} and the
 * scales branch {ocr_preprocess, find_scales_preprocess} may be called concurrently from
 * two threads on one handle; they use separate streams and disjoint buffers.
 * load_frame / crop_to_map / destroy are exclusive (`&mut self` / ordered in the reference).
 */
#ifndef SMH_B200_H
#define SMH_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct smh_b200 smh_b200;

enum {
    SMH_OK = 0,
    SMH_E_ARG = -1,       /* null pointer / bad argument */
    SMH_E_GEOMETRY = -2,  /* frame size for which the reference panics (crop out of bounds, D12) */
    SMH_E_STATE = -3,     /* call order violated (e.g. crop_to_map before load_frame) */
    SMH_E_ALIGN = -4,     /* device-resident input not 16-byte aligned (base, pitch, frame stride) */
    SMH_E_NOGPU = -5,     /* no CUDA device / driver: the backend has NO CPU fallback */
    SMH_E_RANGE = -6      /* scales_start_y > height (reference: debug-build panic, D9) */
};

/* DebugView discriminants — vision-common/src/debug.rs:31-40 */
enum { SMH_VIEW_NONE = 0, SMH_VIEW_OCR_INPUT, SMH_VIEW_FIND_SCALES_INPUT, SMH_VIEW_LSD_PREPROCESS,
       SMH_VIEW_LSD_INPUT, SMH_VIEW_CROPPED_BRQ };

/* ---- lifecycle: Vision::init / Drop — vision-gpu/src/cuda.rs:31-93, dylib.rs:79-97 ---------- */
int32_t smh_b200_create(int32_t device, smh_b200 **out);
void smh_b200_destroy(smh_b200 *h);
const char *smh_b200_last_error(smh_b200 *h); /* h may be NULL: error of the last failed create */
/* Vision::thread_ctx — vision-common/src/lib.rs:37, vision-gpu/src/lib.rs:150-165 */
int32_t smh_b200_bind_thread(smh_b200 *h);

/* ROI rectangles the reference derives from the frame size (screen.rs:55-66,
 * consts/mod.rs:7-19, vision-cpu/src/lib.rs:137-145). roi = {button xywh, map xywh, brq xywh},
 * all in frame coordinates.  Pure host arithmetic, needs no GPU. */
int32_t smh_b200_roi(uint32_t W, uint32_t H, uint32_t roi[12]);

/* ---- per-frame trait methods ------------------------------------------------------------- */

/* Vision::load_frame — vision-common/src/lib.rs:44; CPU vision-cpu/src/lib.rs:69-93;
 * GPU vision-gpu/src/lib.rs:167-193.  bgra: HOST pointer, B,G,R,A bytes, pitch_bytes per row.
 * Uploads only the rows/columns the path reads (map ROI + button ROI). The buffer may be
 * freed when the call returns. */
int32_t smh_b200_load_frame(smh_b200 *h, const uint8_t *bgra, uint32_t W, uint32_t H, uint32_t pitch_bytes);

/* Vision::crop_to_map — lib.rs:46; CPU :110-171; GPU :221-297.  Runs the fused pixel pass
 * (button test, ui_map, marker mask + dilation, OCR/scales preprocessing) once.
 * *map_open = 0 mirrors Ok(None).  ui_map_rgba: borrowed pinned host image, w*h*4 bytes,
 * tightly packed, valid until the next load_frame. xywh = map ROI in frame coordinates. */
int32_t smh_b200_crop_to_map(smh_b200 *h, int32_t grayscale, int32_t *map_open, uint32_t xywh[4],
                             const uint8_t **ui_map_rgba);

/* Vision::ocr_preprocess — lib.rs:49; CPU :173-231; GPU :300-321.  Borrowed pinned host L8
 * image of the bottom-right quadrant (len = bqw*bqh), valid until the next load_frame. */
int32_t smh_b200_ocr_preprocess(smh_b200 *h, const uint8_t **out, size_t *len);

/* Vision::find_scales_preprocess — lib.rs:50; CPU :233-251; GPU :323-346.  Rows >= start_y are
 * refreshed (luma != 0 ? 255 : 0); rows above keep the previous frame's content (D9; zero on a
 * fresh handle / after a resolution change).  Borrowed host L8 image bqw*bqh. */
int32_t smh_b200_find_scales_preprocess(smh_b200 *h, uint32_t scales_start_y, const uint8_t **out,
                                        uint32_t wh[2]);

/* Vision::isolate_map_markers / mask_marker_lines — lib.rs:52,54; CPU :253-280,357-375;
 * GPU :348-371,432-475.  The fused pass already produced the dilated mask; these wait for it. */
int32_t smh_b200_isolate_map_markers(smh_b200 *h);
int32_t smh_b200_mask_marker_lines(smh_b200 *h);

/* Vision::load_map_markers / filter_map_marker_icons — the marker icon template match the reference
 * carries COMMENTED OUT (trait: vision-common/src/lib.rs:46,53; CPU vision-cpu/src/lib.rs:101-108,
 * 270-277,282-355; GPU vision-gpu/src/lib.rs:129-137,202-204,373-430 + cuda.cu:548-614; orchestrator
 * src/vision/mod.rs:48-49,95-107).  Off unless the host calls it.
 *
 * load_map_markers: n_markers (<= 8) templates of size x size RGBA pixels, tightly packed, in the
 * layout vision-gpu uploads (`load_markers::<UnfilteredMarkers>(size)`, vision-gpu/src/gpumarkers.rs:24-52;
 * the TGA decode / Gaussian resize / colour isolation / corner mask of
 * vision-common/src/markers/mod.rs:96-119 stay host-side `image`-crate work).  size <= 32, n_markers*size*size <= 8192.
 *
 * filter_map_marker_icons: call between isolate_map_markers and mask_marker_lines, at most once per frame.
 * Scores every candidate pixel (marker-coloured, size <= x < w-size-1, size <= y < h-size-1) against every
 * template — sum over opaque template pixels of u32(f32(|dr|+|dg|+|db|) * (f32(alpha)/255)) on the isolated
 * crop — takes the global minimum (ties: first in raster order, D17; the reference's order is
 * nondeterministic), blacks the icon's size x size square out of the crop and leaves the 4x4 green square at
 * its tip, so that the mask, the line search and the LSDPreprocess / LSDInput debug views see the edit.
 * *found = 0 when the map holds no candidate (or is smaller than the template). */
int32_t smh_b200_load_map_markers(smh_b200 *h, const uint8_t *rgba, uint32_t n_markers, uint32_t size);
int32_t smh_b200_filter_map_marker_icons(smh_b200 *h, int32_t *found, uint32_t xy[2], uint32_t *sad);

/* Vision::find_longest_line — lib.rs:55; CPU :387-449; GPU :489-538 (+K10).  Searches the
 * handle's current dilated mask (the trait's `image` argument is the backend's own LSDImage).
 * line = {p0.x,p0.y,p1.x,p1.y} — layout of #[repr(C)] Line<f32>, util/src/geometry.rs:168-173. */
int32_t smh_b200_find_longest_line(smh_b200 *h, float x, float y, float max_gap, float line[4],
                                   float *len_sq);

/* Vision::find_marker_lines — lib.rs:56; CPU :377-385 + vision-common/src/lsd.rs:60-105.
 * The whole sequential seed scan runs on the device in one launch. */
int32_t smh_b200_find_marker_lines(smh_b200 *h, uint32_t max_gap, float lines[32][4], uint32_t *n);

/* Vision::get_debug_view — lib.rs:60; CPU :451-460.  Borrowed host RGBA image; *rgba = NULL for
 * SMH_VIEW_NONE. Valid until the next debug_view / load_frame. */
int32_t smh_b200_debug_view(smh_b200 *h, uint8_t which, const uint8_t **rgba, uint32_t wh[2]);

/* Host copy of the dilated marker mask (what the reference GPU backend keeps pinned for the
 * host-side line search, vision-gpu/src/lib.rs:470-474).  L8 w*h, borrowed. */
int32_t smh_b200_lsd_image(smh_b200 *h, const uint8_t **l8, uint32_t wh[2]);

/* red-pixel count of the last crop_to_map's button test (vision-gpu/src/lib.rs:236-243) */
int32_t smh_b200_red_pixels(smh_b200 *h, uint32_t *count);

/* ---- the callers either side of the trait ----------------------- */

/* find_minimap — src/vision/find_minimap.rs:51-146, called by the orchestrator on
 * `vision.get_cpu_frame().view(x, y, w, h)` (src/vision/mod.rs:85).  Here it runs on the device copy of
 * the loaded frame, so the host frame is never needed again.  rect = {left, right, top, bottom} in map
 * ROI coordinates; *found = 0 mirrors None (view smaller than 3x3).  Needs load_frame. */
int32_t smh_b200_find_minimap(smh_b200 *h, int32_t *found, uint32_t rect[4]);

/* One "NNNm" label the OCR found: (meters, (x, y)) with x = (left + right) / 2, y = bottom of the text,
 * in bottom-right-quadrant coordinates — src/vision/mod.rs:183 */
typedef struct { uint32_t meters, x, y; } smh_b200_scale;

/* calc_meters_to_px_ratio — src/vision/mpx_ratio.rs:3-133, on the device copy of the image
 * find_scales_preprocess last returned (including its stale rows, D9).  n_scales = 0..3.
 * *found = 0 mirrors None.  debug_lines[k] = {meters, left, right, y} of each bar found (the reference's
 * scales_debug, before the brq offset is added), may be NULL.  D16: where the reference reads outside
 * the image (unchecked in release builds, a panic in debug builds) the pixel counts as non-black. */
int32_t smh_b200_calc_meters_to_px_ratio(smh_b200 *h, const smh_b200_scale *scales, uint32_t n_scales, int32_t *found,
                                         double *ratio, uint32_t debug_lines[3][4], uint32_t *n_debug);

/* VisionState::process — src/vision/mod.rs:36-240 — as one call: load_frame, crop_to_map, find_minimap,
 * then the marker branch {isolate, mask, find_marker_lines(max_gap)} overlapped with the scales branch
 * {ocr_preprocess, `ocr` callback (Tesseract's place: src/vision/mod.rs:127-190), find_scales_preprocess,
 * calc_meters_to_px_ratio}.  The callback receives the preprocessed L8 quadrant and reports up to 3
 * scales and scales_start_y (the smallest `bottom` of the labels); returning 0 scales mirrors
 * `return Ok(None)`.  ocr = NULL skips the scales branch (a heightmap is set, mod.rs:120-122);
 * detect_markers = 0 skips the marker branch (mod.rs:88,116-118). */
typedef int32_t (*smh_b200_ocr_fn)(void *user, const uint8_t *ocr_l8, uint32_t w, uint32_t h, smh_b200_scale scales[3],
                                   uint32_t *n_scales, uint32_t *scales_start_y);
typedef struct {
    int32_t map_open;               /* 0: process() returned Ok(None); nothing else is valid */
    uint32_t xywh[4];               /* map ROI */
    const uint8_t *ui_map_rgba;     /* VisionResults::map (borrowed, w*h*4) */
    float markers[32][4];           /* VisionResults::markers */
    uint32_t n_markers;
    int32_t has_ratio;              /* VisionResults::meters_to_px_ratio: Option<f64> */
    double meters_to_px_ratio;
    int32_t has_minimap;            /* VisionResults::minimap_bounds: Option<Rect<u32>> */
    uint32_t minimap[4];            /* left, right, top, bottom */
    uint32_t scales_debug[3][4];    /* meters, left, right, y */
    uint32_t n_scales_debug;
    uint32_t reserved;
} smh_b200_vision_results;
int32_t smh_b200_process(smh_b200 *h, const uint8_t *bgra, uint32_t W, uint32_t H, uint32_t pitch_bytes, int32_t grayscale,
                         int32_t detect_markers, uint32_t max_gap, smh_b200_ocr_fn ocr, void *user,
                         smh_b200_vision_results *out);

/* ---- batch extension (not in the trait): N independent frames per launch ------------------ */

/* A device image plane: pixel (i,j) of frame f lives at
 *   ptr + f*frame_stride + j*pitch + (i + pad)*bytes_per_pixel.
 * Rows are 128-byte aligned and `pad` keeps the plane's 16-byte vector grid in phase with the captured frame's
 * (a TMA box and a 128-bit vector must start on a 16-byte boundary of the capture); the quadrant planes are padded
 * further so that a 128-column strip of the pixel pass starts on a 128-byte boundary in them too. */
typedef struct {
    uint8_t *ptr;
    uint64_t frame_stride;
    uint32_t pitch;
    uint32_t pad;             /* left pad in pixels: 0..3 for the map planes, up to 127 for the quadrant planes */
    uint32_t width, height, bytes_per_pixel;
    uint32_t reserved;
} smh_b200_plane;

typedef struct {
    uint32_t n_frames, W, H, reserved;
    uint32_t roi[12];          /* button, map, brq (frame coordinates) */
    smh_b200_plane ui_map;     /* RGBA8 w*h */
    smh_b200_plane lsd_image;  /* L8 w*h, dilated marker mask 0/255 */
    smh_b200_plane ocr_out;    /* L8 bqw*bqh */
    smh_b200_plane scales;     /* L8 bqw*bqh; rows < start_y[f] are zero: every frame of a batch starts from a fresh buffer (D9) */
    float *lines;              /* device, [n][32][4] */
    uint32_t *n_lines;         /* device, [n] */
    uint32_t *red_pixels;      /* device, [n] */
    uint8_t *map_open;         /* device, [n]: 0 = map closed (n_lines = 0, images still written) */
    uint64_t *lsd_stats;       /* device, [n][3] = {seeds, ray_steps, pruned seeds} when stats were requested, else NULL */
    uint32_t *minimap;         /* device, [n][4] = left, right, top, bottom (SMH_BATCH_MINIMAP), else NULL */
    uint8_t *minimap_found;    /* device, [n] */
    uint8_t *icon_found;       /* device, [n]: template match found (SMH_BATCH_FILTER_ICONS), else NULL */
    uint32_t *icon_xy;         /* device, [n][2] = x, y of the matched icon's window in the map crop */
    uint32_t *icon_sad;        /* device, [n] */
} smh_b200_batch;

#define SMH_BATCH_GRAYSCALE 1u /* ui_map = (L,L,L,255) — settings default, src/settings.rs:135 */
#define SMH_BATCH_LSD_STATS 2u /* count seeds / ray steps per frame (slower; for reporting) */
#define SMH_BATCH_SKIP_LSD 4u  /* pixel pass only */
#define SMH_BATCH_SKIP_PIXELS 8u /* line search only, on the planes of the previous call */
#define SMH_BATCH_SERIAL 64u   /* one launch per kernel on the caller's stream (profiling: no chunk overlap) */
#define SMH_BATCH_FILTER_ICONS 128u /* run the (reference-disabled) icon template match between the mask and the
                                       line search; needs smh_b200_load_map_markers */
#define SMH_BATCH_MINIMAP 32u   /* also run find_minimap on every open frame (src/vision/find_minimap.rs) */
#define SMH_BATCH_UI_L8 256u     /* smh_b200_process_host_batch with GRAYSCALE: res->ui_map receives the L plane (n * w*h bytes) —
                                   every pixel of a grayscale ui_map is (L, L, L, 255), so a caller that can show an L8 image
                                   saves the widening pass and three quarters of the host writes */
#define SMH_BATCH_LSD_EXHAUSTIVE 16u /* march all 3600 rays of every seed (no isolated-blob pruning): makes
                                        the ray-step counter comparable with the CPU oracle's */

/* Process n device-resident frames (tightly stacked: frame f at d_frames + f*frame_stride,
 * rows pitch_bytes apart; base, pitch and stride multiples of 16).  start_y: HOST array of n
 * values or NULL (= 0).  All work is enqueued on `stream` (a cudaStream_t, may be NULL) and NOT
 * synchronised.  The output planes are owned by the handle and reused by the next call with
 * the same n/W/H. */
int32_t smh_b200_process_batch(smh_b200 *h, const uint8_t *d_frames, uint32_t n, uint32_t W, uint32_t H,
                               uint32_t pitch_bytes, uint64_t frame_stride, const uint32_t *start_y,
                               uint32_t max_gap, uint32_t flags, void *stream, smh_b200_batch *out);

/* Host-resident results of a batch, tightly packed (what the trait hands back per frame). */
typedef struct {
    uint8_t *ui_map;      /* n * w*h*4   (may be NULL to skip); n * w*h with SMH_BATCH_UI_L8 */
    uint8_t *ocr_out;     /* n * bqw*bqh (may be NULL) */
    uint8_t *scales;      /* n * bqw*bqh (may be NULL) */
    uint8_t *lsd_image;   /* n * w*h     (may be NULL) */
    float *lines;         /* n * 32*4 */
    uint32_t *n_lines;    /* n */
    uint32_t *red_pixels; /* n */
    uint8_t *map_open;    /* n */
    uint32_t *minimap;    /* n * 4, with SMH_BATCH_MINIMAP (may be NULL) */
    uint8_t *minimap_found; /* n (may be NULL) */
    uint8_t *icon_found;  /* n, with SMH_BATCH_FILTER_ICONS (may be NULL) */
    uint32_t *icon_xy;    /* n * 2 (may be NULL) */
    uint32_t *icon_sad;   /* n (may be NULL) */
} smh_b200_host_results;

/* End-to-end batch from HOST frames (pinned memory recommended): uploads each frame's ROI band,
 * runs the path and copies the results back, pipelined over chunks of `chunk` frames on
 * internal streams; returns after everything has landed in `res`.  With SMH_BATCH_GRAYSCALE the
 * ui_map crosses the bus as its L plane and is widened to (L, L, L, 255) in `res->ui_map` by host
 * threads of the library (or handed over as it is: SMH_BATCH_UI_L8).  Flags honoured: GRAYSCALE, UI_L8, SKIP_LSD,
 * MINIMAP, FILTER_ICONS. */
int32_t smh_b200_process_host_batch(smh_b200 *h, const uint8_t *frames, uint32_t n, uint32_t W, uint32_t H,
                                    uint32_t pitch_bytes, uint64_t frame_stride, const uint32_t *start_y,
                                    uint32_t max_gap, uint32_t flags, uint32_t chunk,
                                    const smh_b200_host_results *res);

/* The same pipeline over a bounded RING of host buffers — how a capture loop with a fixed pool of frame buffers feeds
 * it: `frames` holds ring_n frames, frame f of the n processed is frames[f % ring_n] and its results land at index
 * f % ring_n of every array of `res` (sized for ring_n frames; start_y still has n entries).  The caller owns the
 * reuse discipline: a slot's results are overwritten ring_n frames later.  ring_n >= 4 * chunk. */
int32_t smh_b200_process_host_stream(smh_b200 *h, const uint8_t *frames, uint32_t ring_n, uint32_t n, uint32_t W, uint32_t H,
                                     uint32_t pitch_bytes, uint64_t frame_stride, const uint32_t *start_y,
                                     uint32_t max_gap, uint32_t flags, uint32_t chunk,
                                     const smh_b200_host_results *res);

/* ---- several GPUs of one box -------------------------------------------------------------
 * The reference is single-device (vision-gpu/src/cuda.rs:34, `Device::get_device(0)`); frames are independent, so a
 * batch is split into contiguous frame ranges — GPU g of G takes frames [g*n/G, (g+1)*n/G) — each run by its own host
 * thread on its own handle (context + streams) through smh_b200_process_host_batch.  No inter-GPU traffic, no
 * collective; results land in `res` in the caller's frame order. */
typedef struct smh_b200_multi smh_b200_multi;
/* devices: n_dev CUDA device ordinals (a device may appear once). */
int32_t smh_b200_multi_create(const int32_t *devices, uint32_t n_dev, smh_b200_multi **out);
void smh_b200_multi_destroy(smh_b200_multi *m);
const char *smh_b200_multi_last_error(smh_b200_multi *m);
uint32_t smh_b200_multi_device_count(smh_b200_multi *m);
/* the frame range GPU g of n_dev takes of a batch of n frames: [*first, *first + *count) */
void smh_b200_shard_range(uint32_t n, uint32_t n_dev, uint32_t g, uint32_t *first, uint32_t *count);
/* Same arguments and results as smh_b200_process_host_batch.  load_map_markers_rgba/n_markers/size: templates for
 * SMH_BATCH_FILTER_ICONS (may be NULL / 0).  Returns the first non-zero code of any shard. */
int32_t smh_b200_multi_process_host_batch(smh_b200_multi *m, const uint8_t *frames, uint32_t n, uint32_t W, uint32_t H,
                                          uint32_t pitch_bytes, uint64_t frame_stride, const uint32_t *start_y,
                                          uint32_t max_gap, uint32_t flags, uint32_t chunk,
                                          const smh_b200_host_results *res);
/* smh_b200_load_map_markers on every device of the set */
int32_t smh_b200_multi_load_map_markers(smh_b200_multi *m, const uint8_t *rgba, uint32_t n_markers, uint32_t size);
/* per-device wall-clock seconds of the last smh_b200_multi_process_host_batch (host threads run concurrently) */
int32_t smh_b200_multi_last_times(smh_b200_multi *m, double *seconds, uint32_t n);

/* calc_meters_to_px_ratio for every frame of the last smh_b200_process_batch on this handle, reading its
 * device-resident `scales` planes.  scales: HOST [n][3], n_scales: HOST [n] (0..3 each); results land in
 * the HOST arrays ratio[n], found[n] after the call (it synchronises `stream`). */
int32_t smh_b200_batch_meters_to_px_ratio(smh_b200 *h, const smh_b200_scale *scales, const uint8_t *n_scales, uint32_t n,
                                          void *stream, double *ratio, uint8_t *found);

/* Copy plane `p` of frames [first, first+count) to tightly packed host memory. */
int32_t smh_b200_download_plane(smh_b200 *h, const smh_b200_plane *p, uint32_t first, uint32_t count,
                                uint8_t *host_out, void *stream);

/* cudaStreamSynchronize(stream) on the handle's device (stream may be NULL). */
int32_t smh_b200_stream_synchronize(smh_b200 *h, void *stream);

/* number of kernel launches this handle has issued (bench.py's gpu_launches claim) */
uint64_t smh_b200_launch_count(smh_b200 *h);

/* Per-stage timing of the last frame through the trait calls — the reference's `Timeshares`
 * (vision-common/src/debug.rs:3-29, filled by debug_waterfall! around every call, src/vision/mod.rs:54-64).
 * Same eleven stages, same order.  host_ms[i]: wall-clock time of the stage's last call (what the reference records);
 * device_ms[i]: CUDA-event time of the device work that call enqueued (kernels + copies; 0 where the stage has none —
 * the fused pixel pass runs inside crop_to_map, so isolate_map_markers / mask_marker_lines cost nothing here).
 * SMH_STAGE_OCR is the callback of smh_b200_process (Tesseract's place).  Either array may be NULL. */
enum { SMH_STAGE_LOAD_FRAME = 0, SMH_STAGE_CROP_TO_MAP, SMH_STAGE_FIND_MINIMAP, SMH_STAGE_OCR_PREPROCESS, SMH_STAGE_OCR,
       SMH_STAGE_FIND_SCALES_PREPROCESS, SMH_STAGE_CALC_METERS_TO_PX_RATIO, SMH_STAGE_ISOLATE_MAP_MARKERS,
       SMH_STAGE_FILTER_MAP_MARKER_ICONS, SMH_STAGE_MASK_MARKER_LINES, SMH_STAGE_FIND_MARKER_LINES, SMH_STAGE_COUNT };
int32_t smh_b200_timeshares(smh_b200 *h, float host_ms[SMH_STAGE_COUNT], float device_ms[SMH_STAGE_COUNT]);
/* the stage's name as the reference spells it ("load_frame", ...); NULL when out of range */
const char *smh_b200_stage_name(uint32_t stage);

/* ---- synthetic captures; replaces src/capture.rs in the benchmarks ------- */
typedef struct {
    uint64_t seed;        /* frame k uses seed + k */
    uint32_t W, H;
    uint32_t n_segments;  /* marker lines (3 px wide, 60..600 px) with an icon disc at the start */
    uint32_t n_blips;     /* isolated 2..4 px marker-coloured squares */
    uint32_t n_labels;    /* "NNNm"-like OCR boxes + scale bars in the bottom-right quadrant (<=4) */
    uint32_t map_open;    /* paint the red close-deployment button */
    uint32_t blip_band_pct; /* blips confined to the top pct % of the map (100 = anywhere) */
    uint32_t reserved;
} smh_b200_synth_params;

/* CPU rasteriser: frames [first, first+n) -> host BGRA, tightly packed. Needs no GPU. */
int32_t smh_b200_synth_frames_host(const smh_b200_synth_params *p, uint32_t first, uint32_t n, uint8_t *out);
/* CUDA rasteriser: byte-identical frames written to device memory (frame stride W*H*4). */
int32_t smh_b200_synth_frames_device(smh_b200 *h, const smh_b200_synth_params *p, uint32_t first, uint32_t n,
                                     uint8_t *d_out, void *stream);

/* ---- exhaustive colour tables computed ON THE DEVICE (parity tests) ---------------------- */
/* out[r<<16|g<<8|b] for all 2^24 colours; host output buffer. which: 0 marker predicate and
 * 1 luma8 through the exact code paths the pixel kernel uses (integer fast-reject + bit table;
 * I2F byte-select conversions); 2 button-red; 3 ocr class (bit0 core, bit1 edge); 4 marker
 * predicate and 5 luma8 through the plain float formulas; 6 luma8 through the biased-word
 * (PRMT) conversions the quadrant vectors use. */
int32_t smh_b200_colour_table(smh_b200 *h, uint32_t which, uint8_t *host_out);

#ifdef __cplusplus
}
#endif
#endif /* SMH_B200_H */
