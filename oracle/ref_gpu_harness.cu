// ref_gpu_harness.cu — runs the REFERENCE's own CUDA kernels (vision-gpu/cuda/cuda.cu, compiled
// unmodified from where they lie under /root/reference, plus its NPP dilate wrapper dilate.cpp) in the
// launch sequence of the reference's GPU backend.  TEST INFRASTRUCTURE ONLY (lives under oracle/,
// builds into oracle/_ref/libsmh_refgpu.so, never linked or loaded by the product).
//
// Two uses:
//   * pin the CPU oracle: outputs of the reference itself, run on the B200, are compared with the
//     oracle wherever the two reference backends agree by construction (SURVEY.md 8.A lists where
//     they do not: D1 HSV rounding, D2 FMA luma, D4 dilation shape, D5 gap rule, D6 ties, D10 OCR clip);
//   * the "GPU kernel to beat" of SURVEY.md 8(d): the reference's launch shapes, per-stage
//     synchronisation, per-frame allocation and per-seed launch + copy + sync loop, timed on the same B200.
//
// The host sequence below restates vision-gpu/src/lib.rs (cited per step); the kernels are the
// reference's own object code.  lsd::find_lines (host code shared by both reference backends,
// vision-common/src/lsd.rs:60-105) is the oracle's restatement, driven through a callback exactly as
// vision-gpu plugs its kernel in (vision-gpu/src/lib.rs:477-487).
//
// build: make -C oracle refgpu   (only where /root/reference is mounted)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <chrono>
#include <vector>

#include "cuda.cu"     // the reference's kernels (-I /root/reference/vision-gpu/cuda)
#include "dilate.cpp"  // the reference's NPP wrapper

#include "smh_oracle.h"

#define RCHECK(expr)                                                                                  \
    do {                                                                                              \
        cudaError_t e__ = (expr);                                                                     \
        if (e__ != cudaSuccess) {                                                                     \
            snprintf(g_err, sizeof g_err, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return (int)e__;                                                                          \
        }                                                                                             \
    } while (0)

static char g_err[512];

struct refgpu {
    uint32_t W = 0, H = 0, roi[12] = {0};
    // GpuMemory — vision-gpu/src/lib.rs:33-104
    cudaStream_t stream_a = nullptr, stream_b = nullptr, markers_stream = nullptr, scales_stream = nullptr;
    BGRA* frame = nullptr;
    uint32_t* red_pixels = nullptr;
    RGBA* ui_map = nullptr;
    RGB* cropped_map = nullptr;
    RGB* cropped_brq = nullptr;
    uint8_t* ocr_out = nullptr; uint8_t* ocr_out_host = nullptr;           // PinnedGpuImage
    uint8_t* scales = nullptr; std::vector<uint8_t> scales_host;           // pageable GrayImage
    uint8_t* lsd_image = nullptr; uint8_t* lsd_image_host = nullptr;       // PinnedGpuImage
    uint8_t* lsd_image_dilate = nullptr;
    uint8_t* lsd_kernel_dilate = nullptr;
    Line* longest_lines = nullptr;
    uint64_t seeds = 0, launches = 0;
    // per-kernel device time (CUDA events around every launch) when `timing` is on: K1..K10 of SURVEY.md 2.2 + the upload
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double kernel_us[12] = {0};
    uint64_t kernel_launches[12] = {0};
};

enum { RK_UPLOAD = 0, RK_K1 = 1, RK_K2 = 2, RK_K3 = 3, RK_K4 = 4, RK_K5 = 5, RK_K6 = 6, RK_K7 = 7, RK_K8 = 8, RK_K9 = 9, RK_K10 = 10 };
// bracket one launch with events on its stream and add the elapsed time to the kernel's account
#define TIMED(r, k, stream, launch)                                                              \
    do {                                                                                         \
        if ((r)->timing) cudaEventRecord((r)->ev0, stream);                                      \
        launch;                                                                                  \
        if ((r)->timing) {                                                                       \
            cudaEventRecord((r)->ev1, stream);                                                   \
            cudaEventSynchronize((r)->ev1);                                                      \
            float ms__ = 0.0f;                                                                   \
            cudaEventElapsedTime(&ms__, (r)->ev0, (r)->ev1);                                     \
            (r)->kernel_us[k] += 1e3 * ms__; (r)->kernel_launches[k]++;                          \
        }                                                                                        \
    } while (0)

static void release(refgpu* r) {
    cudaFree(r->frame); cudaFree(r->red_pixels); cudaFree(r->ui_map); cudaFree(r->cropped_map); cudaFree(r->cropped_brq);
    cudaFree(r->ocr_out); cudaFreeHost(r->ocr_out_host); cudaFree(r->scales); cudaFree(r->lsd_image); cudaFreeHost(r->lsd_image_host);
    cudaFree(r->lsd_image_dilate); cudaFree(r->lsd_kernel_dilate); cudaFree(r->longest_lines);
    r->frame = nullptr; r->red_pixels = nullptr; r->ui_map = nullptr; r->cropped_map = nullptr; r->cropped_brq = nullptr;
    r->ocr_out = nullptr; r->ocr_out_host = nullptr; r->scales = nullptr; r->lsd_image = nullptr; r->lsd_image_host = nullptr;
    r->lsd_image_dilate = nullptr; r->lsd_kernel_dilate = nullptr; r->longest_lines = nullptr;
    r->W = r->H = 0;
}

// GpuMemory::new — vision-gpu/src/lib.rs:60-104 (buffers are `uninitialized` there; zeroed here so that
// the 1-px frame NPP never writes (D4) is deterministic)
static int ensure(refgpu* r, uint32_t W, uint32_t H) {
    if (r->W == W && r->H == H) return 0;
    release(r);
    if (smh_oracle_roi(W, H, r->roi)) { snprintf(g_err, sizeof g_err, "frame size the reference panics on"); return -2; }
    const size_t w = r->roi[6], h = r->roi[7], qw = r->roi[10], qh = r->roi[11];
    RCHECK(cudaMalloc(&r->red_pixels, 4));
    RCHECK(cudaMalloc(&r->ui_map, w * h * 4));
    RCHECK(cudaMalloc(&r->cropped_map, w * h * 3));
    RCHECK(cudaMalloc(&r->cropped_brq, qw * qh * 3));
    RCHECK(cudaMalloc(&r->ocr_out, qw * qh)); RCHECK(cudaMallocHost(&r->ocr_out_host, qw * qh));
    RCHECK(cudaMalloc(&r->scales, qw * qh)); RCHECK(cudaMemset(r->scales, 0, qw * qh));
    r->scales_host.assign(qw * qh, 0);
    RCHECK(cudaMalloc(&r->lsd_image, w * h)); RCHECK(cudaMemset(r->lsd_image, 0, w * h));
    RCHECK(cudaMallocHost(&r->lsd_image_host, w * h));
    RCHECK(cudaMalloc(&r->lsd_image_dilate, w * h));
    RCHECK(cudaMalloc(&r->lsd_kernel_dilate, 4)); RCHECK(cudaMemset(r->lsd_kernel_dilate, 255, 4));
    RCHECK(cudaMalloc(&r->longest_lines, 8 * sizeof(Line)));
    r->W = W; r->H = H;
    return 0;
}

static inline dim3 grid2d(uint32_t w, uint32_t h, uint32_t bx, uint32_t by) { return dim3((w + bx - 1) / bx, (h + by - 1) / by); }

// CudaInstance::find_longest_line — vision-gpu/src/lib.rs:489-538: <<<8, 450>>>, 8 block winners copied
// back, stream sync, host fold keeping the FIRST strictly greater length
static int ref_longest(refgpu* r, const uint8_t* d_img, uint32_t w, uint32_t h, float x, float y, float max_gap,
                       float line[4], float* len_sq) {
    TIMED(r, RK_K10, r->markers_stream, (find_longest_line<<<8, (360 * 10) / 8, 0, r->markers_stream>>>(d_img, w, h, Point{x, y}, max_gap, r->longest_lines)));
    r->launches++;
    Line block[8];
    RCHECK(cudaMemcpyAsync(block, r->longest_lines, sizeof block, cudaMemcpyDeviceToHost, r->markers_stream));
    RCHECK(cudaStreamSynchronize(r->markers_stream));
    Line best{{0.f, 0.f}, {0.f, 0.f}};
    float best_len = 0.0f;
    for (int i = 0; i < 8; i++) {
        const float ddx = block[i].p0.x - block[i].p1.x, ddy = block[i].p0.y - block[i].p1.y;
        const float l = ddx * ddx + ddy * ddy;  // powi(2) + powi(2)
        if (l > best_len) { best = block[i]; best_len = l; }
    }
    line[0] = best.p0.x; line[1] = best.p0.y; line[2] = best.p1.x; line[3] = best.p1.y;
    *len_sq = best_len;
    return 0;
}

static int longest_cb(void* user, const uint8_t* img, uint32_t w, uint32_t h, float x, float y, float max_gap, float line[4],
                      float* len_sq) {
    (void)img;
    refgpu* r = (refgpu*)user;
    r->seeds++;
    return ref_longest(r, r->lsd_image, w, h, x, y, max_gap, line, len_sq);
}

extern "C" {

const char* refgpu_last_error(void) { return g_err; }

int refgpu_create(refgpu** out) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { snprintf(g_err, sizeof g_err, "no CUDA device"); return -5; }
    refgpu* r = new refgpu();
    // four non-blocking streams — vision-gpu/src/lib.rs:76-78, cuda.rs `stream!`
    RCHECK(cudaStreamCreateWithFlags(&r->stream_a, cudaStreamNonBlocking));
    RCHECK(cudaStreamCreateWithFlags(&r->stream_b, cudaStreamNonBlocking));
    RCHECK(cudaStreamCreateWithFlags(&r->markers_stream, cudaStreamNonBlocking));
    RCHECK(cudaStreamCreateWithFlags(&r->scales_stream, cudaStreamNonBlocking));
    RCHECK(cudaEventCreate(&r->ev0)); RCHECK(cudaEventCreate(&r->ev1));
    *out = r;
    return 0;
}

void refgpu_destroy(refgpu* r) {
    if (!r) return;
    cudaDeviceSynchronize();
    release(r);
    cudaStreamDestroy(r->stream_a); cudaStreamDestroy(r->stream_b); cudaStreamDestroy(r->markers_stream); cudaStreamDestroy(r->scales_stream);
    cudaEventDestroy(r->ev0); cudaEventDestroy(r->ev1);
    delete r;
}

/* One captured frame through the reference's GPU backend in the orchestrator's order
 * (src/vision/mod.rs:66-223).  bgra: host, tightly packed.  Every output may be NULL.
 * Returns 0 / a cudaError_t / an NPP status; *map_open = 0 mirrors Ok(None). */
int refgpu_process_frame(refgpu* r, const uint8_t* bgra, uint32_t W, uint32_t H, int grayscale, uint32_t start_y,
                         uint32_t max_gap, int* map_open, uint32_t* red_out, uint8_t* ui_map, uint8_t* cropped_map_isolated,
                         uint8_t* cropped_brq, uint8_t* raw_mask, uint8_t* lsd_image, uint8_t* ocr_out, uint8_t* scales,
                         float lines[32][4], uint32_t* n_lines, uint64_t* seeds) {
    int rc = ensure(r, W, H);
    if (rc) return rc;
    const uint32_t* roi = r->roi;
    const uint32_t w = roi[6], h = roi[7], qw = roi[10], qh = roi[11];
    r->seeds = 0;
    if (n_lines) *n_lines = 0;

    // ---- load_frame — vision-gpu/src/lib.rs:167-193: a fresh device image per frame
    // (GpuImage::async_try_from allocates; the old one is dropped), H2D from pageable memory,
    // red_pixels <- 0, stream sync
    cudaFree(r->frame); r->frame = nullptr;
    RCHECK(cudaMalloc(&r->frame, (size_t)W * H * 4));
    TIMED(r, RK_UPLOAD, r->stream_a, RCHECK(cudaMemcpyAsync(r->frame, bgra, (size_t)W * H * 4, cudaMemcpyHostToDevice, r->stream_a)));
    const uint32_t zero = 0;
    RCHECK(cudaMemcpyAsync(r->red_pixels, &zero, 4, cudaMemcpyHostToDevice, r->stream_a));
    RCHECK(cudaStreamSynchronize(r->stream_a));

    // ---- crop_to_map — :206-297
    {
        const uint32_t bx = roi[0], by = roi[1], bw = roi[2], bh = roi[3];
        TIMED(r, RK_K1, r->stream_a, (count_close_deployment_button_red_pixels<<<grid2d(bw, bh, 16, 16), dim3(16, 16), 4, r->stream_a>>>(
            r->frame, W, bx, by, bw, bh, r->red_pixels)));
        r->launches++;
        uint32_t red = 0;
        RCHECK(cudaMemcpyAsync(&red, r->red_pixels, 4, cudaMemcpyDeviceToHost, r->stream_a));
        RCHECK(cudaStreamSynchronize(r->stream_a));
        if (red_out) *red_out = red;
        const float ratio = (float)red / (float)(bw * bh);
        if (ratio < CLOSE_DEPLOYMENT_BUTTON_RED_PIXEL_THRESHOLD) { if (map_open) *map_open = 0; return 0; }
        if (map_open) *map_open = 1;
    }
    {
        const uint32_t x = roi[4], y = roi[5];
        if (grayscale) TIMED(r, RK_K3, r->stream_a, (crop_to_map_grayscale<<<grid2d(w, h, 8, 8), dim3(8, 8), 0, r->stream_a>>>(r->frame, W, x, y, w, h, r->cropped_map, r->ui_map)));
        else TIMED(r, RK_K2, r->stream_a, (crop_to_map<<<grid2d(w, h, 8, 8), dim3(8, 8), 0, r->stream_a>>>(r->frame, W, x, y, w, h, r->cropped_map, r->ui_map)));
        TIMED(r, RK_K4, r->stream_b, (crop_to_map_brq<<<grid2d(qw, qh, 8, 8), dim3(8, 8), 0, r->stream_b>>>(r->frame, W, x + qw, y + qh, qw, qh, r->cropped_brq)));
        r->launches += 2;
        std::vector<uint8_t> ui_host((size_t)w * h * 4);  // `vec![0u8; ..]` per frame, pageable (:289-290)
        RCHECK(cudaMemcpyAsync(ui_host.data(), r->ui_map, ui_host.size(), cudaMemcpyDeviceToHost, r->stream_a));
        RCHECK(cudaStreamSynchronize(r->stream_a));
        RCHECK(cudaStreamSynchronize(r->stream_b));
        if (ui_map) memcpy(ui_map, ui_host.data(), ui_host.size());
    }
    if (cropped_brq) RCHECK(cudaMemcpy(cropped_brq, r->cropped_brq, (size_t)qw * qh * 3, cudaMemcpyDeviceToHost));

    // ---- markers branch: isolate_map_markers (:348-371), mask_marker_lines (:432-475), find_marker_lines (:477-487)
    TIMED(r, RK_K7, r->markers_stream, (isolate_map_markers<<<grid2d(w, h, 8, 8), dim3(8, 8), 0, r->markers_stream>>>(r->cropped_map, w, h)));
    r->launches++;
    RCHECK(cudaStreamSynchronize(r->markers_stream));
    if (cropped_map_isolated) RCHECK(cudaMemcpy(cropped_map_isolated, r->cropped_map, (size_t)w * h * 3, cudaMemcpyDeviceToHost));
    TIMED(r, RK_K8, r->markers_stream, (mask_marker_lines<<<grid2d(w, h, 8, 8), dim3(8, 8), 0, r->markers_stream>>>(r->cropped_map, w, h, r->lsd_image_dilate)));
    r->launches++;
    RCHECK(cudaStreamSynchronize(r->markers_stream));
    if (raw_mask) RCHECK(cudaMemcpy(raw_mask, r->lsd_image_dilate, (size_t)w * h, cudaMemcpyDeviceToHost));
    {
        int res = 0;
        TIMED(r, RK_K9, (cudaStream_t)0, (res = gpu_dilate(r->lsd_image_dilate, r->lsd_image, w, h, r->lsd_kernel_dilate, 2, 2)));  // NPP, default stream
        r->launches++;
        if (res != 0) { snprintf(g_err, sizeof g_err, "nppiDilate_8u_C1R error %d", res); return res; }
        RCHECK(cudaMemcpy(r->lsd_image_host, r->lsd_image, (size_t)w * h, cudaMemcpyDeviceToHost));  // copy_from_gpu (blocking)
        if (lsd_image) memcpy(lsd_image, r->lsd_image_host, (size_t)w * h);
    }
    {
        float l[32][4];
        uint32_t n = 0;
        rc = smh_oracle_find_lines_cb(r->lsd_image_host, w, h, max_gap, longest_cb, r, l, &n);
        if (rc) return rc;
        if (lines) memcpy(lines, l, sizeof l);
        if (n_lines) *n_lines = n;
        if (seeds) *seeds = r->seeds;
    }

    // ---- scales branch: ocr_preprocess (:300-321), [Tesseract], find_scales_preprocess (:323-346)
    TIMED(r, RK_K5, r->scales_stream, (ocr_preprocess<<<grid2d(qw, qh, 8, 8), dim3(8, 8), 0, r->scales_stream>>>(r->cropped_brq, qw, qh, r->ocr_out)));
    r->launches++;
    RCHECK(cudaMemcpyAsync(r->ocr_out_host, r->ocr_out, (size_t)qw * qh, cudaMemcpyDeviceToHost, r->scales_stream));
    RCHECK(cudaStreamSynchronize(r->scales_stream));
    if (ocr_out) memcpy(ocr_out, r->ocr_out_host, (size_t)qw * qh);
    TIMED(r, RK_K6, r->scales_stream, (find_scales_preprocess<<<grid2d(qw, qh, 8, 8), dim3(8, 8), 0, r->scales_stream>>>(r->cropped_brq, qw, qh, start_y, r->scales)));
    r->launches++;
    RCHECK(cudaMemcpyAsync(r->scales_host.data(), r->scales, (size_t)qw * qh, cudaMemcpyDeviceToHost, r->scales_stream));
    RCHECK(cudaStreamSynchronize(r->scales_stream));
    if (scales) memcpy(scales, r->scales_host.data(), (size_t)qw * qh);
    return 0;
}

/* find_longest_line of the reference GPU backend on a caller-supplied host L8 image */
int refgpu_find_longest_line_in(refgpu* r, const uint8_t* img, uint32_t w, uint32_t h, float x, float y, float max_gap,
                                float line[4], float* len_sq) {
    uint8_t* d = nullptr;
    RCHECK(cudaMalloc(&d, (size_t)w * h));
    RCHECK(cudaMemcpy(d, img, (size_t)w * h, cudaMemcpyHostToDevice));
    const int rc = ref_longest(r, d, w, h, x, y, max_gap, line, len_sq);
    cudaFree(d);
    return rc;
}

/* Wall-clock frames/s of refgpu_process_frame over `n` host frames (tightly stacked): `warm` untimed frames, then
 * `passes` timed passes over all n frames; frames_per_s[passes] (take the median). */
int refgpu_time_frames(refgpu* r, const uint8_t* frames, uint32_t n, uint32_t W, uint32_t H, int grayscale, uint32_t max_gap,
                       uint32_t warm, uint32_t passes, double* frames_per_s, uint64_t* seeds_total, uint64_t* launches_total) {
    uint64_t seeds = 0;
    int open = 0, rc = 0;
    r->timing = false;
    for (uint32_t i = 0; i < warm; i++) {  // allocations, module load, clocks
        rc = refgpu_process_frame(r, frames + (size_t)(i % n) * W * H * 4, W, H, grayscale, 0, max_gap, &open, nullptr, nullptr, nullptr, nullptr,
                                  nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        if (rc) return rc;
    }
    RCHECK(cudaDeviceSynchronize());
    const uint64_t l0 = r->launches;
    for (uint32_t k = 0; k < passes; k++) {
        const auto t0 = std::chrono::steady_clock::now();
        for (uint32_t i = 0; i < n; i++) {
            uint64_t s = 0;
            rc = refgpu_process_frame(r, frames + (size_t)i * W * H * 4, W, H, grayscale, 0, max_gap, &open, nullptr, nullptr, nullptr,
                                      nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &s);
            if (rc) return rc;
            seeds += s;
        }
        RCHECK(cudaDeviceSynchronize());
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (frames_per_s) frames_per_s[k] = (double)n / dt;
    }
    if (seeds_total) *seeds_total = seeds;
    if (launches_total) *launches_total = r->launches - l0;
    return 0;
}

/* Device time of every reference kernel (CUDA events around each launch, reference launch shapes) summed over `n`
 * frames: us[11] = {upload, K1, K2, K3, K4, K5, K6, K7, K8, K9 (NPP), K10 (all seeds)}, launches[11] likewise. */
int refgpu_kernel_times(refgpu* r, const uint8_t* frames, uint32_t n, uint32_t W, uint32_t H, int grayscale, uint32_t max_gap,
                        double* us, uint64_t* launches) {
    int open = 0, rc = 0;
    r->timing = false;
    for (uint32_t i = 0; i < 4; i++) {
        rc = refgpu_process_frame(r, frames + (size_t)(i % n) * W * H * 4, W, H, grayscale, 0, max_gap, &open, nullptr, nullptr, nullptr, nullptr,
                                  nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        if (rc) return rc;
    }
    for (int k = 0; k < 12; k++) { r->kernel_us[k] = 0.0; r->kernel_launches[k] = 0; }
    r->timing = true;
    for (uint32_t i = 0; i < n && !rc; i++)
        rc = refgpu_process_frame(r, frames + (size_t)i * W * H * 4, W, H, grayscale, 0, max_gap, &open, nullptr, nullptr, nullptr, nullptr,
                                  nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    r->timing = false;
    for (int k = 0; k < 11; k++) { us[k] = r->kernel_us[k]; launches[k] = r->kernel_launches[k]; }
    return rc;
}

}  // extern "C"

/* ---- per-ray / per-kernel access to the reference's object code (pins oracle/smh_oracle.c's vision-gpu
 *      flavour bit for bit; tests/test_gpu_refcuda.py) ------------------------------------------------ */

// The 3600 ray directions exactly as K10 computes them (cuda.cu:654,669-670): the same expression compiled by
// the same nvcc with the same flags in this translation unit, so CUDA's cosf/sinf expand to the same code.
// (K10 itself only reveals them for the first quadrant — a seed at (1e-30, 1e-30) of an all-black image ends
// every ray at exactly (-dx, -dy) — which the test uses to check this table against the kernel.)
__global__ void harness_ray_dirs(float* out) {
    const unsigned k = threadIdx.x + blockIdx.x * blockDim.x;
    if (k >= 3600) return;
    const float theta = ((float)k / 10.0) * CUDART_PI_F / 180.0;
    out[2 * k] = cosf(theta);
    out[2 * k + 1] = sinf(theta);
}

extern "C" {

int refgpu_ray_dirs(refgpu* r, float* dxdy /* 7200 */) {
    float* d = nullptr;
    RCHECK(cudaMalloc(&d, 7200 * sizeof(float)));
    harness_ray_dirs<<<8, 450, 0, r->markers_stream>>>(d);
    RCHECK(cudaMemcpyAsync(dxdy, d, 7200 * sizeof(float), cudaMemcpyDeviceToHost, r->markers_stream));
    RCHECK(cudaStreamSynchronize(r->markers_stream));
    cudaFree(d);
    return 0;
}

/* K10 launched as <<<3600, 1>>>: the kernel indexes its ray by threadIdx.x + blockIdx.x * blockDim.x and writes
 * one Line per BLOCK (cuda.cu:654,730-738), so with one thread per block the unmodified object code hands back
 * EVERY ray's line.  img: host L8; seeds: n x (x, y); lines: n x 3600 x 4 floats. */
int refgpu_rays_per_seed(refgpu* r, const uint8_t* img, uint32_t w, uint32_t h, const float* seeds, uint32_t n, float max_gap,
                         float* lines) {
    uint8_t* d = nullptr; Line* dl = nullptr;
    RCHECK(cudaMalloc(&d, (size_t)w * h));
    RCHECK(cudaMalloc(&dl, (size_t)n * 3600 * sizeof(Line)));
    RCHECK(cudaMemcpy(d, img, (size_t)w * h, cudaMemcpyHostToDevice)); RCHECK(cudaDeviceSynchronize());
    for (uint32_t i = 0; i < n; i++) {
        find_longest_line<<<3600, 1, 0, r->markers_stream>>>(d, w, h, Point{seeds[2 * i], seeds[2 * i + 1]}, max_gap, dl + (size_t)i * 3600);
        r->launches++;
    }
    RCHECK(cudaGetLastError());
    RCHECK(cudaMemcpyAsync(lines, dl, (size_t)n * 3600 * sizeof(Line), cudaMemcpyDeviceToHost, r->markers_stream));
    RCHECK(cudaStreamSynchronize(r->markers_stream));
    cudaFree(d); cudaFree(dl);
    return 0;
}

/* K10 in vision-gpu's own launch shape <<<8, 450>>> (vision-gpu/src/lib.rs:494-510): the 8 block winners of
 * every seed, before the host fold.  blocks: n x 8 x 4 floats. */
int refgpu_block_winners(refgpu* r, const uint8_t* img, uint32_t w, uint32_t h, const float* seeds, uint32_t n, float max_gap,
                         float* blocks) {
    uint8_t* d = nullptr; Line* dl = nullptr;
    RCHECK(cudaMalloc(&d, (size_t)w * h));
    RCHECK(cudaMalloc(&dl, (size_t)n * 8 * sizeof(Line)));
    RCHECK(cudaMemcpy(d, img, (size_t)w * h, cudaMemcpyHostToDevice)); RCHECK(cudaDeviceSynchronize());
    for (uint32_t i = 0; i < n; i++)
        find_longest_line<<<8, 450, 0, r->markers_stream>>>(d, w, h, Point{seeds[2 * i], seeds[2 * i + 1]}, max_gap, dl + (size_t)i * 8);
    RCHECK(cudaGetLastError());
    RCHECK(cudaMemcpyAsync(blocks, dl, (size_t)n * 8 * sizeof(Line), cudaMemcpyDeviceToHost, r->markers_stream));
    RCHECK(cudaStreamSynchronize(r->markers_stream));
    cudaFree(d); cudaFree(dl);
    return 0;
}

/* K8 + K9 on a caller-supplied RGB crop: mask_marker_lines then NPP dilate exactly as
 * vision-gpu/src/lib.rs:454-472 (lsd_image pre-zeroed so the frame NPP never writes is defined). */
int refgpu_mask_and_dilate(refgpu* r, const uint8_t* rgb, uint32_t w, uint32_t h, uint8_t* raw_mask, uint8_t* lsd_image) {
    RGB* d = nullptr; uint8_t *m = nullptr, *o = nullptr, *k = nullptr;
    RCHECK(cudaMalloc(&d, (size_t)w * h * 3)); RCHECK(cudaMalloc(&m, (size_t)w * h)); RCHECK(cudaMalloc(&o, (size_t)w * h));
    RCHECK(cudaMalloc(&k, 4)); RCHECK(cudaMemset(k, 255, 4)); RCHECK(cudaMemset(o, 0, (size_t)w * h));
    RCHECK(cudaMemcpy(d, rgb, (size_t)w * h * 3, cudaMemcpyHostToDevice)); RCHECK(cudaDeviceSynchronize());
    mask_marker_lines<<<grid2d(w, h, 8, 8), dim3(8, 8), 0, r->markers_stream>>>(d, w, h, m);
    RCHECK(cudaStreamSynchronize(r->markers_stream));
    const int res = gpu_dilate(m, o, w, h, k, 2, 2);
    if (res != 0) { snprintf(g_err, sizeof g_err, "nppiDilate_8u_C1R error %d", res); return res; }
    RCHECK(cudaDeviceSynchronize());
    if (raw_mask) RCHECK(cudaMemcpy(raw_mask, m, (size_t)w * h, cudaMemcpyDeviceToHost));
    if (lsd_image) RCHECK(cudaMemcpy(lsd_image, o, (size_t)w * h, cudaMemcpyDeviceToHost));
    cudaFree(d); cudaFree(m); cudaFree(o); cudaFree(k);
    return 0;
}

/* The template-match kernels vision-gpu carries but never launches (K11/K12, cuda.cu:548-614), driven as the
 * commented host code would (vision-gpu/src/lib.rs:373-430): K11 over `n_cand` candidates (xy = y * w + x) with
 * the templates ONE AT A TIME — with several templates the kernel's threads add into the same u16 without
 * atomics (cuda.cu:583), a data race; one template per launch makes its arithmetic observable.
 * sads: n_markers x n_cand (u16).  Then, when clear_xy != UINT32_MAX, K12 on that position; map_out: the crop after it. */
int refgpu_icons(refgpu* r, const uint8_t* map_rgb, uint32_t w, uint32_t h, const uint8_t* tpl_rgba, uint32_t n_markers, uint32_t size,
                 const uint32_t* cand_xy, uint32_t n_cand, uint16_t* sads, uint32_t clear_xy, uint8_t* map_out) {
    RGB* d = nullptr; RGBA* t = nullptr; const RGBA** tp = nullptr; markers::TemplateMatch* c = nullptr;
    const size_t np = (size_t)size * size;
    RCHECK(cudaMalloc(&d, (size_t)w * h * 3)); RCHECK(cudaMalloc(&t, np * 4 * n_markers)); RCHECK(cudaMalloc(&tp, sizeof(RGBA*)));
    RCHECK(cudaMalloc(&c, sizeof(markers::TemplateMatch) * (n_cand ? n_cand : 1)));
    RCHECK(cudaMemcpy(d, map_rgb, (size_t)w * h * 3, cudaMemcpyHostToDevice)); RCHECK(cudaDeviceSynchronize());
    RCHECK(cudaMemcpy(t, tpl_rgba, np * 4 * n_markers, cudaMemcpyHostToDevice)); RCHECK(cudaDeviceSynchronize());
    std::vector<markers::TemplateMatch> hc(n_cand);
    for (uint32_t m = 0; m < n_markers && n_cand; m++) {
        for (uint32_t i = 0; i < n_cand; i++) hc[i] = markers::TemplateMatch{cand_xy[i], 0};
        RCHECK(cudaMemcpy(c, hc.data(), sizeof(markers::TemplateMatch) * n_cand, cudaMemcpyHostToDevice)); RCHECK(cudaDeviceSynchronize());
        const RGBA* one = t + np * m;
        RCHECK(cudaMemcpy(tp, &one, sizeof one, cudaMemcpyHostToDevice)); RCHECK(cudaDeviceSynchronize());
        filter_map_marker_icons<<<grid2d(1, n_cand, 8, 8), dim3(8, 8), 0, r->markers_stream>>>(d, w, c, tp, size, 1, n_cand);
        RCHECK(cudaStreamSynchronize(r->markers_stream));
        RCHECK(cudaMemcpy(hc.data(), c, sizeof(markers::TemplateMatch) * n_cand, cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < n_cand; i++) sads[(size_t)m * n_cand + i] = hc[i].sad;
    }
    if (clear_xy != 0xFFFFFFFFu) {
        filter_map_marker_icons_clear<<<grid2d(size, size, 8, 8), dim3(8, 8), 0, r->markers_stream>>>(d, w, h, clear_xy, size);
        RCHECK(cudaStreamSynchronize(r->markers_stream));
    }
    if (map_out) RCHECK(cudaMemcpy(map_out, d, (size_t)w * h * 3, cudaMemcpyDeviceToHost));
    cudaFree(d); cudaFree(t); cudaFree(tp); cudaFree(c);
    return 0;
}

}  // extern "C"
