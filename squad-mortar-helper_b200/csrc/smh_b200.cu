// smh_b200.cu — host side of the C ABI (include/smh_b200.h): device context, buffers,
// streams and kernel launches.  No CPU fallback: every compute entry point fails with
// SMH_E_NOGPU / a cudaError_t when there is no usable CUDA device.
//
// Reference host code this replaces: vision-gpu/src/lib.rs:33-559 (GpuMemory, launches,
// copies), vision-gpu/src/cuda.rs:31-93 (context + PTX selection), vision-gpu/src/gpuimage.rs.
#include "../../include/smh_b200.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "kernels_lsd.cuh"
#include "kernels_px.cuh"
#include "kernels_next.cuh"
#include "kernels_icons.cuh"

extern "C" void smh_host_widen_l8_to_rgba(const uint8_t* src, uint32_t* dst, size_t n);  // smh_host_widen.cpp

namespace {

thread_local std::string g_create_error;

// chunks in flight in smh_b200_process_host_batch: upload of chunk c+2, kernels of c+1, download of c
#ifndef SMH_PIPE_DEPTH
#define SMH_PIPE_DEPTH 3
#endif
// warps per frame of the line search (kernels_lsd.cuh)
#ifndef SMH_LSD2_WARPS
#define SMH_LSD2_WARPS 4
#endif
// smh_b200_process_batch splits a large batch into chunks that run on up to this many internal streams
#define SMH_BATCH_LANES_MAX 8

struct Geometry {
    uint32_t W = 0, H = 0;
    uint32_t btn[4] = {0, 0, 0, 0}, map[4] = {0, 0, 0, 0}, brq[4] = {0, 0, 0, 0};
};

// RelativeBound::ScreenH(f).into_absolute — vision-common/src/screen.rs:61-66 (f64 multiply,
// round half away from zero, `as u32`)
inline uint32_t rel_h(double frac, uint32_t H) { return (uint32_t)llround(frac * (double)H); }

// vision-common/src/consts/mod.rs:7-19 + screen.rs:8-46 + vision-cpu/src/lib.rs:116,137-145.
// Fails where the reference panics (util/src/image.rs:9-16,67-73; D12).
int compute_geometry(uint32_t W, uint32_t H, Geometry* g) {
    if (W == 0 || H == 0 || W > 16384 || H > 16384) return SMH_E_GEOMETRY;
    const int64_t bw = rel_h(SMH_BTN_W_FRAC, H), bh = rel_h(SMH_BTN_H_FRAC, H);
    const int64_t bx = (int64_t)W - rel_h(SMH_BTN_X_RIGHT_FRAC, H) - bw;
    const int64_t by = (int64_t)H - rel_h(SMH_BTN_Y_BOTTOM_FRAC, H) - bh;
    const int64_t w0 = rel_h(SMH_MAP_W_FRAC, H), h = rel_h(SMH_MAP_H_FRAC, H), x0 = rel_h(SMH_MAP_X_FRAC, H);
    const int64_t y = (int64_t)H - rel_h(SMH_MAP_Y_BOTTOM_FRAC, H) - h;
    const int64_t w = (int64_t)W - w0;   // "Map fills remaining space"
    const int64_t x = (int64_t)W - x0 - w;
    if (bx < 0 || by < 0 || bw <= 0 || bh <= 0 || x < 0 || y < 0 || w <= 0 || h <= 0) return SMH_E_GEOMETRY;
    if (bx + bw > W || by + bh > H) return SMH_E_GEOMETRY;
    if (x + w >= W || y + h >= H) return SMH_E_GEOMETRY;
    const int64_t bqw = w / 2, bqh = h / 2;
    if (bqw < SMH_OCR_RADIUS || bqh < SMH_OCR_RADIUS) return SMH_E_GEOMETRY;
    g->W = W; g->H = H;
    g->btn[0] = (uint32_t)bx; g->btn[1] = (uint32_t)by; g->btn[2] = (uint32_t)bw; g->btn[3] = (uint32_t)bh;
    g->map[0] = (uint32_t)x; g->map[1] = (uint32_t)y; g->map[2] = (uint32_t)w; g->map[3] = (uint32_t)h;
    g->brq[0] = (uint32_t)(x + bqw); g->brq[1] = (uint32_t)(y + bqh); g->brq[2] = (uint32_t)bqw; g->brq[3] = (uint32_t)bqh;
    return SMH_OK;
}

inline uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

// Output planes of one batch (or of the single-frame path, n = 1).
struct Planes {
    uint32_t n = 0, W = 0, H = 0;
    Geometry geo;
    uint32_t pad = 0, qpad = 0, nvec = 0;
    smh_b200_plane ui{}, lsd{}, ocr{}, scales{};
    float* lines = nullptr;
    uint32_t* n_lines = nullptr;
    uint32_t* red = nullptr;
    uint8_t* open = nullptr;
    unsigned long long* stats = nullptr;
    uint32_t* start_y = nullptr;
    uint32_t* counter = nullptr;  // frame dispenser of the persistent line-search CTAs
    uint32_t* minimap = nullptr;  // [n][4] find_minimap rectangles
    uint8_t* minimap_found = nullptr;
    uint32_t* mpx_in = nullptr;   // [n][3][3] scales of calc_meters_to_px_ratio
    uint8_t* mpx_n = nullptr; double* mpx_ratio = nullptr; uint8_t* mpx_found = nullptr;
    uint32_t* mpx_debug = nullptr; uint8_t* mpx_ndebug = nullptr;
    unsigned long long* icon_best = nullptr;  // [n] template-match keys
    uint8_t* icon_found = nullptr; uint32_t* icon_xy = nullptr; uint32_t* icon_sad = nullptr;
    uint32_t* bits = nullptr;     // tile-major 1-bit copies of the dilated masks (line-search input)
    uint64_t bits_stride = 0;     // words per frame
    smh::BitsGeom bg{};
    uint8_t* edge = nullptr;      // what the strips of the pixel pass tell their neighbours (Pixel2Args::edge)
    uint64_t edge_stride = 0; uint32_t edge_hp = 0;
    void release() {
        cudaFree(bits); cudaFree(edge); cudaFree(minimap); cudaFree(minimap_found);
        cudaFree(icon_best); cudaFree(icon_found); cudaFree(icon_xy); cudaFree(icon_sad);
        cudaFree(mpx_in); cudaFree(mpx_n); cudaFree(mpx_ratio); cudaFree(mpx_found); cudaFree(mpx_debug); cudaFree(mpx_ndebug);
        cudaFree(ui.ptr); cudaFree(lsd.ptr); cudaFree(ocr.ptr); cudaFree(scales.ptr);
        cudaFree(lines); cudaFree(n_lines); cudaFree(red); cudaFree(open); cudaFree(stats); cudaFree(start_y); cudaFree(counter);
        *this = Planes();
    }
};

// Host-side workers of smh_b200_process_host_batch: with a grayscale ui_map every pixel is (L, L, L, 255), so
// only L crosses the bus (a quarter of the bytes of the largest output) and these threads widen it into the
// caller's RGBA buffer while later chunks are still in flight.  A task waits for its chunk's download event.
struct HostPool {
    struct Task { cudaEvent_t ev; const uint8_t* src; uint32_t* dst; size_t n; std::atomic<int>* pending; };
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv, cv_done;
    std::deque<Task> q;
    bool stop = false;
    int device = 0;
    std::atomic<int> failed{0};

    void run() {
        cudaSetDevice(device);
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || !q.empty(); });
                if (q.empty()) return;
                t = q.front(); q.pop_front();
            }
            if (cudaEventSynchronize(t.ev) != cudaSuccess) failed = 1;
            else smh_host_widen_l8_to_rgba(t.src, t.dst, t.n);
            if (t.pending->fetch_sub(1) == 1) { std::lock_guard<std::mutex> lk(mu); cv_done.notify_all(); }
        }
    }
    void start(int dev, unsigned n_threads) {
        if (!th.empty()) return;
        device = dev;
        for (unsigned i = 0; i < n_threads; i++) th.emplace_back([this] { run(); });
    }
    void push(const Task& t) { { std::lock_guard<std::mutex> lk(mu); q.push_back(t); } cv.notify_one(); }
    void wait_zero(std::atomic<int>& pending) {
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending.load() == 0; });
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        for (auto& t : th) t.join();
    }
};

// first byte of every RGBA pixel of a ui_map plane (its L when the map is grayscale), rows tightly packed
__global__ void k_pack_gray(const uint8_t* __restrict__ ui, uint32_t pitch, uint32_t pad, uint32_t w, uint32_t rows, uint8_t* __restrict__ out) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    for (uint32_t r = blockIdx.y; r < rows; r += gridDim.y)
        out[(size_t)r * w + x] = (uint8_t)__ldg(reinterpret_cast<const uint32_t*>(ui + (size_t)r * pitch + (size_t)(pad + x) * 4u));
}

}  // namespace

struct smh_b200 {
    int device = 0;
    int sm_count = 0;
    std::mutex err_mu;
    std::string err;
    std::atomic<uint64_t> launches{0};
    // Timeshares (vision-common/src/debug.rs:3-29): per stage the wall-clock of its last call and the device time of what it enqueued
    float stage_host_ms[SMH_STAGE_COUNT] = {0};
    float stage_dev_ms[SMH_STAGE_COUNT] = {0};
    cudaEvent_t stage_ev[SMH_STAGE_COUNT][2] = {};
    float2* d_rays = nullptr;  // 3600 (cos, sin)
    float2* d_ckpt = nullptr;  // accumulated ray offsets every LSD_CKPT steps, [3600][n_ckpt]
    uint32_t n_ckpt = 0;
    uint32_t* d_lut = nullptr; // marker predicate bit table, 2^24 bits
    // marker icon templates (smh_b200_load_map_markers): one allocation, see smh::IconTable
    uint32_t* d_icons = nullptr;
    smh::IconTable icons{};
    bool icons_applied = false; int icon_found = 0; uint32_t icon_xy[2] = {0, 0};

    // ---- single-frame (trait) state ----
    Geometry geo;                 // of the loaded frame
    bool have_frame = false, cropped = false, map_open = false;
    uint32_t red_pixels = 0;
    int grayscale = 0;
    // device copy of the region the path reads: frame columns [ux0, W), rows [uy0, uy1)
    uint8_t* d_frame = nullptr;
    uint32_t ux0 = 0, uy0 = 0, uy1 = 0, upitch = 0;
    Planes one;                   // n = 1 output planes
    // pinned host mirrors
    uint8_t* h_ui = nullptr; uint8_t* h_ocr = nullptr;
    uint8_t* h_scales = nullptr;  // persistent (stale rows kept, D9)
    uint8_t* h_lsd = nullptr; uint8_t* h_view = nullptr; uint8_t* d_view = nullptr;
    size_t view_cap = 0;
    float* h_small = nullptr;     // pinned scratch: lines[32][4], n, longest[5]
    float* d_longest = nullptr;
    uint8_t* d_scales_keep = nullptr;  // device twin of h_scales (tight rows): what calc_meters_to_px_ratio reads
    bool ocr_fetched = false, lsd_fetched = false;
    cudaStream_t s_main = nullptr, s_markers = nullptr, s_scales = nullptr;
    cudaEvent_t ev_pixels = nullptr;
    std::mutex markers_mu, scales_mu;

    // ---- batch state ----
    Planes batch;
    std::vector<cudaStream_t> pipe_streams;   // host-batch pipeline: [0] upload, [1],[2] compute, [3] download
    std::vector<cudaEvent_t> pipe_events;     // [3 * SMH_PIPE_DEPTH]: uploaded / computed / downloaded per slot
    uint8_t* d_stage = nullptr; size_t stage_cap = 0;   // host-batch upload staging (device)
    uint8_t* d_pack = nullptr; size_t pack_cap = 0;     // host-batch download staging: planes without pitch padding
    Planes pipe[SMH_PIPE_DEPTH];
    HostPool pool;                            // host-batch: widens grayscale ui_map rows (L8 over the bus)
    uint8_t* h_gray = nullptr; size_t gray_cap = 0;   // pinned landing area of the L8 rows, [SMH_PIPE_DEPTH][chunk * w*h]
    cudaEvent_t ev_gray[SMH_PIPE_DEPTH] = {};
    std::atomic<int> gray_pending[SMH_PIPE_DEPTH];
    std::vector<cudaStream_t> lane_streams;   // process_batch: chunks of a large batch, taken in turn
    std::vector<cudaEvent_t> lane_events;     // [0] fork, [1 + lane] join
};

namespace {

int fail(smh_b200* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (h) { std::lock_guard<std::mutex> g(h->err_mu); h->err = buf; }
    else g_create_error = buf;
    return code;
}
#define CU(h, expr)                                                                         \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess)                                                             \
            return fail(h, (int)e__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

// One trait call = one stage of the reference's Timeshares: wall-clock over the scope, CUDA events around the
// device work (begin() / end() on the stream it is enqueued on; every trait call synchronises that stream itself).
struct StageTimer {
    smh_b200* h; int stage; std::chrono::steady_clock::time_point t0; cudaStream_t st = nullptr; bool open = false;
    StageTimer(smh_b200* h_, int stage_) : h(h_), stage(stage_), t0(std::chrono::steady_clock::now()) { if (h) h->stage_dev_ms[stage] = 0.0f; }
    void begin(cudaStream_t s) { if (h && h->stage_ev[stage][0]) { st = s; open = cudaEventRecord(h->stage_ev[stage][0], s) == cudaSuccess; } }
    void end() {  // after the caller synchronised `st`
        if (open && cudaEventRecord(h->stage_ev[stage][1], st) == cudaSuccess && cudaEventSynchronize(h->stage_ev[stage][1]) == cudaSuccess) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, h->stage_ev[stage][0], h->stage_ev[stage][1]) == cudaSuccess) h->stage_dev_ms[stage] = ms;
        }
        open = false;
    }
    ~StageTimer() {
        if (open) end();
        if (h) h->stage_host_ms[stage] = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
};

int alloc_plane(smh_b200* h, smh_b200_plane* p, uint32_t n, uint32_t width, uint32_t height, uint32_t bpp, uint32_t pad) {
    p->width = width; p->height = height; p->bytes_per_pixel = bpp; p->pad = pad;
    p->pitch = round_up((width + pad + 3u) * bpp, 128u);
    p->frame_stride = (uint64_t)p->pitch * height;
    CU(h, cudaMalloc(&p->ptr, p->frame_stride * n));
    CU(h, cudaMemset(p->ptr, 0, p->frame_stride * n));
    return SMH_OK;
}

int ensure_planes(smh_b200* h, Planes* P, uint32_t n, const Geometry& g, bool stats) {
    if (P->n == n && P->W == g.W && P->H == g.H) {
        if (stats && !P->stats) CU(h, cudaMalloc(&P->stats, (size_t)n * 3 * sizeof(unsigned long long)));
        return SMH_OK;
    }
    P->release();
    P->geo = g;
    // every plane carries the capture's 16-byte phase as a left pad (a TMA box / 128-bit vector must start on a 16-byte
    // boundary of the capture); k_pixels2's quadrant planes are padded further so that a 128-column strip's bytes start
    // on a 128-byte boundary
    P->pad = g.map[0] & 3u;
    P->qpad = (P->pad + g.brq[2]) & 127u;
    P->nvec = (P->pad + g.map[2] + 3u) / 4u;
    int rc;
    if ((rc = alloc_plane(h, &P->ui, n, g.map[2], g.map[3], 4, P->pad))) return rc;
    if ((rc = alloc_plane(h, &P->lsd, n, g.map[2], g.map[3], 1, P->pad))) return rc;
    if ((rc = alloc_plane(h, &P->ocr, n, g.brq[2], g.brq[3], 1, P->qpad))) return rc;
    if ((rc = alloc_plane(h, &P->scales, n, g.brq[2], g.brq[3], 1, P->qpad))) return rc;
    CU(h, cudaMalloc(&P->lines, (size_t)n * 32 * 4 * sizeof(float)));
    CU(h, cudaMemset(P->lines, 0, (size_t)n * 32 * 4 * sizeof(float)));
    CU(h, cudaMalloc(&P->n_lines, (size_t)n * sizeof(uint32_t)));
    CU(h, cudaMemset(P->n_lines, 0, (size_t)n * sizeof(uint32_t)));
    CU(h, cudaMalloc(&P->red, (size_t)n * sizeof(uint32_t)));
    CU(h, cudaMalloc(&P->open, (size_t)n));
    CU(h, cudaMalloc(&P->start_y, (size_t)n * sizeof(uint32_t)));
    CU(h, cudaMemset(P->start_y, 0, (size_t)n * sizeof(uint32_t)));
    CU(h, cudaMalloc(&P->counter, SMH_BATCH_LANES_MAX * sizeof(uint32_t)));
    CU(h, cudaMalloc(&P->minimap, (size_t)n * 4 * sizeof(uint32_t)));
    CU(h, cudaMemset(P->minimap, 0, (size_t)n * 4 * sizeof(uint32_t)));
    CU(h, cudaMalloc(&P->minimap_found, (size_t)n));
    CU(h, cudaMemset(P->minimap_found, 0, (size_t)n));
    CU(h, cudaMalloc(&P->icon_best, (size_t)n * sizeof(unsigned long long)));
    CU(h, cudaMalloc(&P->icon_found, (size_t)n));
    CU(h, cudaMemset(P->icon_found, 0, (size_t)n));
    CU(h, cudaMalloc(&P->icon_xy, (size_t)n * 2 * sizeof(uint32_t)));
    CU(h, cudaMemset(P->icon_xy, 0, (size_t)n * 2 * sizeof(uint32_t)));
    CU(h, cudaMalloc(&P->icon_sad, (size_t)n * sizeof(uint32_t)));
    CU(h, cudaMemset(P->icon_sad, 0, (size_t)n * sizeof(uint32_t)));
    CU(h, cudaMalloc(&P->mpx_in, (size_t)n * 9 * sizeof(uint32_t)));
    CU(h, cudaMalloc(&P->mpx_n, (size_t)n));
    CU(h, cudaMalloc(&P->mpx_ratio, (size_t)n * sizeof(double)));
    CU(h, cudaMalloc(&P->mpx_found, (size_t)n));
    CU(h, cudaMalloc(&P->mpx_debug, (size_t)n * 12 * sizeof(uint32_t)));
    CU(h, cudaMalloc(&P->mpx_ndebug, (size_t)n));
    P->bg.pad = P->pad; P->bg.w = g.map[2]; P->bg.h = g.map[3];
    P->bg.tx = (P->pad + g.map[2] + 31u) / 32u; P->bg.ty = (g.map[3] + 31u) / 32u;
    P->bits_stride = (uint64_t)P->bg.tx * P->bg.ty * 32u;
    CU(h, cudaMalloc(&P->bits, P->bits_stride * n * sizeof(uint32_t)));
    CU(h, cudaMemset(P->bits, 0, P->bits_stride * n * sizeof(uint32_t)));  // the pixel pass never writes rows >= h / columns >= nvec
    P->edge_hp = round_up(g.map[3], 4u);
    P->edge_stride = (uint64_t)((P->pad + g.map[2] + 127u) / 128u) * 2u * P->edge_hp;
    CU(h, cudaMalloc(&P->edge, P->edge_stride * n));
    if (stats) CU(h, cudaMalloc(&P->stats, (size_t)n * 3 * sizeof(unsigned long long)));
    P->n = n; P->W = g.W; P->H = g.H;
    return SMH_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// The fused pixel pass over n frames (k_pixels2, kernels_px.cuh).  `frames` may be a cropped view of the capture:
// (ox, oy) is the frame coordinate of its first byte, rows_avail the rows it holds.
int launch_pixels(smh_b200* h, const Planes& P, const uint8_t* frames, uint64_t frame_stride, uint32_t pitch, uint32_t rows_avail,
                   uint32_t ox, uint32_t oy, uint32_t n, bool per_frame_start_y, uint32_t start_y_uniform, int grayscale, cudaStream_t st) {
    const Geometry& g = P.geo;
    if ((pitch & 15u) || (frame_stride & 15u) || ((uintptr_t)frames & 15u))
        return fail(h, SMH_E_ALIGN, "captured frames must be 16-byte aligned (base %p, pitch %u, stride %llu)",
                    (const void*)frames, pitch, (unsigned long long)frame_stride);
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(h, SMH_E_NOGPU, "cuTensorMapEncodeTiled is not available from this driver");
    CUtensorMap tmap;  // the strips' boxes: 128 pixels x PX2_ROWS rows
    {
        const uint64_t fs = frame_stride ? frame_stride : (uint64_t)pitch * rows_avail;
        // the tensor ends with the map (columns rounded up to a 32-byte sector): what a box holds beyond it — the right
        // part of the last strip, the rows a band looks at below the map — is zero-filled without being fetched
        const uint32_t vx = g.map[0] - ox + g.map[2], vy = g.map[1] - oy + g.map[3];
        const cuuint64_t dims[3] = {std::min<uint32_t>(pitch / 4u, round_up(vx, 8u)), std::min<uint32_t>(rows_avail, vy), n};
        const cuuint64_t strides[2] = {pitch, fs};
        const cuuint32_t box[3] = {smh::PX2_BOX_W, smh::PX2_ROWS, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void*)frames, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(h, SMH_E_ARG, "cuTensorMapEncodeTiled failed with %d (pitch %u, rows %u, frames %u)", (int)r, pitch, rows_avail, n);
    }
    smh::Pixel2Args a{};
    a.frames = frames; a.frame_stride = frame_stride; a.pitch = pitch;
    a.map_x = g.map[0] - ox; a.map_y = g.map[1] - oy; a.w = g.map[2]; a.h = g.map[3];
    a.bqw = g.brq[2]; a.bqh = g.brq[3]; a.pad = P.pad;
    if ((a.map_x - a.pad) & 3u) return fail(h, SMH_E_ALIGN, "the capture view must start on a multiple of 4 pixels (x0 %u)", ox);
    a.btn_x = g.btn[0] - ox; a.btn_y = g.btn[1] - oy; a.btn_w = g.btn[2]; a.btn_h = g.btn[3];
    a.ui = P.ui.ptr; a.ui_stride = P.ui.frame_stride; a.ui_pitch = P.ui.pitch;
    a.lsd = P.lsd.ptr; a.lsd_stride = P.lsd.frame_stride; a.lsd_pitch = P.lsd.pitch;
    a.ocr = P.ocr.ptr; a.scales = P.scales.ptr; a.q_stride = P.ocr.frame_stride; a.q_pitch = P.ocr.pitch; a.q_pad = P.qpad;
    a.start_y = per_frame_start_y ? P.start_y : nullptr; a.start_y_uniform = start_y_uniform;
    a.red = P.red; a.open = P.open;
    a.n_frames = n;
    a.lut = h->d_lut;
    a.bits = P.bits; a.bits_stride = P.bits_stride; a.bits_tx = P.bg.tx;
    a.edge = P.edge; a.edge_stride = P.edge_stride; a.edge_hp = P.edge_hp;
    // rows per item: 64; a batch too small to fill the SMs with those (the trait's single frame) takes 16-row bands
    // (four times the warps, a quarter of the rows each: the pass of one frame is bound by a warp's chain of rows)
    const uint32_t nstrips = (a.pad + a.w + 127u) / 128u;
    const uint32_t band = (uint64_t)n * nstrips * ((a.h + 63u) / 64u) >= 8ull * (uint64_t)h->sm_count ? 64u : 16u;
    smh_px_plan_make(&a.plan, a.pad, a.w, a.h, a.bqw, a.bqh, band);
    // one warp per (frame, band, strip) item, plus one per frame for the button test (see k_pixels2)
    const uint64_t items_p = (uint64_t)n * ((uint64_t)smh_px_plan_plain_items(&a.plan) + 1u);
    const uint64_t items_q = (uint64_t)n * smh_px_plan_quad_items(&a.plan);
    const uint64_t blocks_p = (items_p + smh::PX2_WARPS - 1) / smh::PX2_WARPS, blocks_q = (items_q + smh::PX2_WARPS - 1) / smh::PX2_WARPS;
    if (blocks_p > 0x7FFFFFFFull || blocks_q > 0x7FFFFFFFull) return fail(h, SMH_E_ARG, "batch too large for one launch");
    CU(h, cudaMemsetAsync(P.edge, 0, P.edge_stride * n, st));  // the strips' notes to their neighbours: stored only where nonzero
    static const bool merged = [] { const char* e = getenv("SMH_B200_PX_MERGED"); return e ? atoi(e) != 0 : true; }();
    if (merged) {  // one launch, the two kinds of CTAs interleaved (k_pixels2)
        a.blocks_q = (uint32_t)blocks_q;
        if (blocks_p + blocks_q > 0x7FFFFFFFull) return fail(h, SMH_E_ARG, "batch too large for one launch");
        if (grayscale) smh::k_pixels2<true, 2><<<(unsigned)(blocks_p + blocks_q), smh::PX2_THREADS, smh::PX2_SMEM_BYTES, st>>>(tmap, a);
        else smh::k_pixels2<false, 2><<<(unsigned)(blocks_p + blocks_q), smh::PX2_THREADS, smh::PX2_SMEM_BYTES, st>>>(tmap, a);
        h->launches++;
        CU(h, cudaGetLastError());
    } else {
        if (grayscale) smh::k_pixels2<true, 0><<<(unsigned)blocks_p, smh::PX2_THREADS, smh::PX2_SMEM_BYTES, st>>>(tmap, a);
        else smh::k_pixels2<false, 0><<<(unsigned)blocks_p, smh::PX2_THREADS, smh::PX2_SMEM_BYTES, st>>>(tmap, a);
        h->launches++;
        CU(h, cudaGetLastError());
        if (blocks_q) {
            if (grayscale) smh::k_pixels2<true, 1><<<(unsigned)blocks_q, smh::PX2_THREADS, smh::PX2_SMEM_BYTES, st>>>(tmap, a);
            else smh::k_pixels2<false, 1><<<(unsigned)blocks_q, smh::PX2_THREADS, smh::PX2_SMEM_BYTES, st>>>(tmap, a);
            h->launches++;
            CU(h, cudaGetLastError());
        }
    }
    if (nstrips > 1u) {  // the pixels either side of the seams between strips
        const uint64_t threads = (uint64_t)n * (nstrips - 1u) * (a.edge_hp >> 2);
        smh::k_px_seams<<<(unsigned)((threads + 255u) / 256u), 256, 0, st>>>(a);
        h->launches++;
        CU(h, cudaGetLastError());
    }
    return SMH_OK;
}

// offsets table of the warp-per-ray march: long enough for any ray inside this mask plus two blocks
int ensure_ray_ckpt(smh_b200* h, const Planes& P, cudaStream_t st) {
    const double diag = sqrt((double)P.bg.w * P.bg.w + (double)P.bg.h * P.bg.h);
    const uint32_t need = ((uint32_t)diag / 32u + 5u) * (32u / smh::LSD_CKPT);
    if (need > h->n_ckpt) {
        CU(h, cudaDeviceSynchronize());
        cudaFree(h->d_ckpt); h->d_ckpt = nullptr; h->n_ckpt = 0;
        CU(h, cudaMalloc(&h->d_ckpt, (size_t)SMH_LSD_RAYS * need * sizeof(float2)));
        smh::k_build_ray_ckpt<<<(SMH_LSD_RAYS + 127) / 128, 128, 0, st>>>(h->d_rays, need, h->d_ckpt);
        h->launches++;
        CU(h, cudaGetLastError());
        CU(h, cudaStreamSynchronize(st));
        h->n_ckpt = need;
    }
    return SMH_OK;
}

// Frames [f0, f0 + cn) of a batch's planes as a batch of their own (shallow: never release()d);
// `lane` picks the frame dispenser the line search of this view uses.
Planes planes_view(const Planes& P, uint32_t f0, uint32_t cn, uint32_t lane) {
    Planes V = P;
    V.n = cn;
    V.ui.ptr += (uint64_t)f0 * P.ui.frame_stride; V.lsd.ptr += (uint64_t)f0 * P.lsd.frame_stride;
    V.ocr.ptr += (uint64_t)f0 * P.ocr.frame_stride; V.scales.ptr += (uint64_t)f0 * P.scales.frame_stride;
    V.lines += (size_t)f0 * 128; V.n_lines += f0; V.red += f0; V.open += f0; V.start_y += f0;
    if (V.stats) V.stats += (size_t)f0 * 3;
    V.counter += lane;
    V.minimap += (size_t)f0 * 4; V.minimap_found += f0;
    V.icon_best += f0; V.icon_found += f0; V.icon_xy += (size_t)f0 * 2; V.icon_sad += f0;
    V.bits += (uint64_t)f0 * P.bits_stride;
    V.edge += (uint64_t)f0 * P.edge_stride;
    return V;
}

// Marker icon template match over n frames (after the pixel pass): scores, then edits mask + bitmap.
int launch_icons(smh_b200* h, const Planes& P, const uint8_t* frames, uint64_t frame_stride, uint32_t pitch, uint32_t ox, uint32_t oy,
                 uint32_t n, bool use_open, cudaStream_t st) {
    if (!h->d_icons) return fail(h, SMH_E_STATE, "filter_map_marker_icons before load_map_markers");
    const Geometry& g = P.geo;
    const uint32_t size = h->icons.size;
    smh::k_fill_u64<<<(n + 255u) / 256u, 256, 0, st>>>(P.icon_best, n, smh::ICON_NONE);
    h->launches++;
    CU(h, cudaGetLastError());
    // `if w >= map_marker_size && h >= map_marker_size` (src/vision/mod.rs:104); the candidate bounds
    // w - size - 1 of vision-cpu/src/lib.rs:272-273 must not wrap
    if (g.map[2] >= size + 2u && g.map[3] >= size + 2u) {
        smh::IconArgs a{};
        a.frames = frames; a.frame_stride = frame_stride; a.pitch = pitch;
        a.map_x = g.map[0] - ox; a.map_y = g.map[1] - oy; a.w = g.map[2]; a.h = g.map[3];
        a.lut = h->d_lut; a.bits = P.bits; a.bits_stride = P.bits_stride; a.g = P.bg;
        a.open = use_open ? P.open : nullptr;
        a.t = h->icons; a.n_frames = n; a.best = P.icon_best;
        a.counter = P.counter;  // this lane's dispenser; the line search zeroes it again before its own launch
        CU(h, cudaMemsetAsync(P.counter, 0, sizeof(uint32_t), st));
        const uint32_t WIN = smh::ICON_TILE + size - 1u, WP = WIN | 1u;
        const size_t npx = (size_t)size * size;
        const size_t dyn = (128u + (size_t)WIN * WP + smh::ICON_TILE * smh::ICON_TILE + 3u * npx * h->icons.n + (size_t)h->icons.n * size + P.bg.tx) * 4u;
        {
            static std::mutex mu;
            static size_t raised[64] = {0};
            std::lock_guard<std::mutex> lk(mu);
            size_t& cur = raised[h->device & 63];
            if (dyn > 32768 && dyn > cur) {
                CU(h, cudaFuncSetAttribute(smh::k_icons, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
                cur = dyn;
            }
        }
        int nb = 0;
        CU(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, smh::k_icons, smh::ICON_THREADS, dyn));
        const uint64_t nitems = (uint64_t)P.bg.ty * n;
        uint64_t grid = (uint64_t)h->sm_count * (nb > 0 ? nb : 1);
        if (grid > nitems) grid = nitems;
        smh::k_icons<<<(unsigned)grid, smh::ICON_THREADS, dyn, st>>>(a);
        h->launches++;
        CU(h, cudaGetLastError());
    }
    smh::IconApplyArgs b{};
    b.frames = frames; b.frame_stride = frame_stride; b.pitch = pitch;
    b.map_x = g.map[0] - ox; b.map_y = g.map[1] - oy; b.w = g.map[2]; b.h = g.map[3];
    b.lut = h->d_lut;
    b.lsd = P.lsd.ptr; b.lsd_stride = P.lsd.frame_stride; b.lsd_pitch = P.lsd.pitch; b.pad = P.pad;
    b.bits = P.bits; b.bits_stride = P.bits_stride; b.g = P.bg;
    b.size = size; b.best = P.icon_best; b.found = P.icon_found; b.xy = P.icon_xy; b.sad = P.icon_sad;
    smh::k_icon_apply<<<n, smh::ICON_THREADS, 0, st>>>(b);
    h->launches++;
    CU(h, cudaGetLastError());
    return SMH_OK;
}

int launch_lsd(smh_b200* h, const Planes& P, uint32_t n, float max_gap, bool use_open, bool stats, bool prune, cudaStream_t st) {
    { int rc = ensure_ray_ckpt(h, P, st); if (rc) return rc; }
    smh::LsdArgs a{};
    a.bits = P.bits; a.bits_stride = P.bits_stride; a.g = P.bg;
    a.rays = h->d_rays; a.ckpt = h->d_ckpt; a.n_ckpt = h->n_ckpt; a.max_gap = max_gap;
    // `gap.0 >= max_gap` on an integer-valued counter <=> counter >= ceil(max_gap)
    a.gap_limit = (max_gap != max_gap || max_gap >= 2147483648.0f) ? 2147483647 : (max_gap <= -2147483648.0f ? -2147483647 : (int)ceilf(max_gap));
    a.n_frames = n; a.counter = P.counter;
    a.open = use_open ? P.open : nullptr;
    a.lines = P.lines; a.n_lines = P.n_lines;
    a.stats = stats ? P.stats : nullptr;
    CU(h, cudaMemsetAsync(P.counter, 0, sizeof(uint32_t), st));
    if (stats) CU(h, cudaMemsetAsync(P.stats, 0, (size_t)n * 3 * sizeof(unsigned long long), st));
    // the limit on dynamic shared memory is a property of the kernel on this device, shared by every handle: only ever raise it
    auto raise_smem = [&](const void* fn, size_t dyn, size_t* cur) -> int {
        if (dyn > 32768 && dyn > *cur) {
            CU(h, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            *cur = dyn;
        }
        return SMH_OK;
    };
    static std::mutex mu;
    static size_t raised[64][6] = {{0}};
    if (prune && a.gap_limit >= 1 && a.gap_limit <= 32) {
        // k_lsd2: a small CTA per frame in flight, warp 0 its master (kernels_lsd.cuh); persistent CTAs.  A batch that cannot
        // fill the SMs anyway (the trait's single frame) takes wider CTAs: the bulk phases of a seed finish sooner
        auto launch = [&](auto nw_tag, int slot) -> int {
            constexpr int NW = decltype(nw_tag)::value;
            const size_t dyn = (sizeof(smh::Lsd2Shared<NW>) + 15) / 16 * 16 + (size_t)P.bg.tx * 33u * sizeof(uint32_t);
            {
                std::lock_guard<std::mutex> g(mu);
                int rc;
                if ((rc = raise_smem((const void*)smh::k_lsd2<NW, false>, dyn, &raised[h->device & 63][slot]))) return rc;
                if ((rc = raise_smem((const void*)smh::k_lsd2<NW, true>, dyn, &raised[h->device & 63][slot + 1]))) return rc;
            }
            int nb = 0;
            CU(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, smh::k_lsd2<NW, false>, NW * 32, dyn));
            // persistent CTAs per SM: five finish a 1024-frame batch sooner than the seven that fit (3.20 vs 3.33 ms: fewer
            // CTAs than frames = the faster CTAs take more frames, and less contention) and leave registers for the next
            // batch's pixel pass; SMH_B200_LSD_CTAS_PER_SM overrides (0 = all that fit)
            static const int cap = [] { const char* e = getenv("SMH_B200_LSD_CTAS_PER_SM"); return e ? atoi(e) : 5; }();
            if (cap >= 1 && nb > cap) nb = cap;
            const uint32_t resident = (uint32_t)(h->sm_count * (nb > 0 ? nb : 1));
            const uint32_t grid = n < resident ? n : resident;
            if (stats) smh::k_lsd2<NW, true><<<grid, NW * 32, dyn, st>>>(a);
            else smh::k_lsd2<NW, false><<<grid, NW * 32, dyn, st>>>(a);
            h->launches++;
            CU(h, cudaGetLastError());
            return SMH_OK;
        };
        static const int lat_nw = [] { const char* e = getenv("SMH_B200_LSD_LATENCY_WARPS"); return e ? atoi(e) : 8; }();
        if (n * 2u <= (uint32_t)h->sm_count && lat_nw == 8) return launch(std::integral_constant<int, 8>(), 4);
        return launch(std::integral_constant<int, SMH_LSD2_WARPS>(), 0);
    }
    // the literal search (every ray of every seed; any max_gap): one CTA per frame in flight
    // one tile-row of the bitmap + its fringe map
    const size_t dyn = (size_t)P.bg.tx * 32u * 2u * sizeof(uint32_t);
    {
        std::lock_guard<std::mutex> g(mu);
        int rc;
        if ((rc = raise_smem((const void*)smh::k_lsd<false>, dyn, &raised[h->device & 63][2]))) return rc;
        if ((rc = raise_smem((const void*)smh::k_lsd<true>, dyn, &raised[h->device & 63][3]))) return rc;
    }
    int nb = 0;
    CU(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, smh::k_lsd<false>, smh::LSD_THREADS, dyn));
    const uint32_t resident = (uint32_t)(h->sm_count * (nb > 0 ? nb : 1));
    const uint32_t grid = n < resident ? n : resident;
    if (stats) smh::k_lsd<true><<<grid, smh::LSD_THREADS, dyn, st>>>(a);
    else smh::k_lsd<false><<<grid, smh::LSD_THREADS, dyn, st>>>(a);
    h->launches++;
    CU(h, cudaGetLastError());
    return SMH_OK;
}

// find_minimap over n frames: one CTA per (direction, frame)
int launch_minimap(smh_b200* h, const Planes& P, const uint8_t* frames, uint64_t frame_stride, uint32_t pitch, uint32_t ox, uint32_t oy,
                   uint32_t n, bool use_open, cudaStream_t st) {
    const Geometry& g = P.geo;
    smh::MinimapArgs a{};
    a.frames = frames; a.frame_stride = frame_stride; a.pitch = pitch;
    a.map_x = g.map[0] - ox; a.map_y = g.map[1] - oy; a.w = g.map[2]; a.h = g.map[3];
    a.n_frames = n; a.open = use_open ? P.open : nullptr;
    a.rect = P.minimap; a.found = P.minimap_found;
    if (n > 65535u) return fail(h, SMH_E_ARG, "batch too large for one launch");
    smh::k_minimap<<<dim3(4, n), smh::MM_THREADS, 0, st>>>(a);
    h->launches++;
    CU(h, cudaGetLastError());
    return SMH_OK;
}

// calc_meters_to_px_ratio over n frames; the scales were already copied to P.mpx_in / P.mpx_n
int launch_mpx(smh_b200* h, const Planes& P, const uint8_t* img, uint64_t stride, uint32_t pitch, uint32_t pad, uint32_t n,
               cudaStream_t st) {
    smh::MpxArgs a{};
    a.scales = img; a.stride = stride; a.pitch = pitch; a.pad = pad; a.w = P.geo.brq[2]; a.h = P.geo.brq[3];
    a.n_frames = n; a.in = P.mpx_in; a.n_in = P.mpx_n; a.ratio = P.mpx_ratio; a.found = P.mpx_found;
    a.debug = P.mpx_debug; a.n_debug = P.mpx_ndebug;
    smh::k_mpx_ratio<<<n, 96, 0, st>>>(a);
    h->launches++;
    CU(h, cudaGetLastError());
    return SMH_OK;
}

// chunk size / stream count of the chunked form of smh_b200_process_batch (environment = tuning knobs)
void batch_tuning(const smh_b200* h, uint32_t* chunk, uint32_t* lanes) {
    const char* ec = getenv("SMH_B200_BATCH_CHUNK");
    const char* el = getenv("SMH_B200_BATCH_LANES");
    int c = ec ? atoi(ec) : 0, l = el ? atoi(el) : 0;
    // Measured on B200 (1024 dense 1440p frames, profiles/r1h_sweep_overlap.txt): one launch per kernel
    // 7.54 ms; best chunked form (6 streams x 222 frames) 7.80 ms, 3 x 148 10.8 ms.  The line search is
    // bound by the per-frame latency of its sequential seed chain, so splitting it into more, smaller
    // waves costs more than the overlap returns: the chunked form is opt-in (SMH_B200_BATCH_LANES >= 2).
    if (c <= 0) c = 2 * h->sm_count;
    if (l <= 0) l = 1;
    if (l > SMH_BATCH_LANES_MAX) l = SMH_BATCH_LANES_MAX;
    *chunk = (uint32_t)c; *lanes = (uint32_t)l;
}

void fill_batch_out(const Planes& P, bool stats, bool minimap, bool icons, smh_b200_batch* out) {
    memset(out, 0, sizeof *out);
    out->n_frames = P.n; out->W = P.W; out->H = P.H;
    memcpy(out->roi, P.geo.btn, 16); memcpy(out->roi + 4, P.geo.map, 16); memcpy(out->roi + 8, P.geo.brq, 16);
    out->ui_map = P.ui; out->lsd_image = P.lsd; out->ocr_out = P.ocr; out->scales = P.scales;
    out->lines = P.lines; out->n_lines = P.n_lines; out->red_pixels = P.red; out->map_open = P.open;
    out->lsd_stats = stats ? (uint64_t*)P.stats : nullptr;
    out->minimap = minimap ? P.minimap : nullptr;
    out->minimap_found = minimap ? P.minimap_found : nullptr;
    out->icon_found = icons ? P.icon_found : nullptr; out->icon_xy = icons ? P.icon_xy : nullptr; out->icon_sad = icons ? P.icon_sad : nullptr;
}

int download_plane(smh_b200* h, const smh_b200_plane* p, uint32_t first, uint32_t count, uint8_t* host_out, cudaStream_t st) {
    const size_t row = (size_t)p->width * p->bytes_per_pixel;
    // one strided copy: rows of all frames are `pitch` apart and frame_stride = pitch*height
    CU(h, cudaMemcpy2DAsync(host_out, row, p->ptr + (uint64_t)first * p->frame_stride + (size_t)p->pad * p->bytes_per_pixel,
                            p->pitch, row, (size_t)p->height * count, cudaMemcpyDeviceToHost, st));
    return SMH_OK;
}

void free_single(smh_b200* h) {
    cudaFree(h->d_frame); h->d_frame = nullptr;
    cudaFree(h->d_scales_keep); h->d_scales_keep = nullptr;
    h->one.release();
    cudaFreeHost(h->h_ui); cudaFreeHost(h->h_ocr); cudaFreeHost(h->h_scales);
    cudaFreeHost(h->h_lsd);
    h->h_ui = h->h_ocr = h->h_scales = h->h_lsd = nullptr;
}

}  // namespace

extern "C" {

const char* smh_b200_last_error(smh_b200* h) {
    if (!h) return g_create_error.c_str();
    // a copy per calling thread: the two branches of a frame may fail concurrently on one handle
    thread_local std::string copy;
    std::lock_guard<std::mutex> g(h->err_mu);
    copy = h->err;
    return copy.c_str();
}

const char* smh_b200_stage_name(uint32_t stage) {
    static const char* const names[SMH_STAGE_COUNT] = {"load_frame", "crop_to_map", "find_minimap", "ocr_preprocess", "ocr",
        "find_scales_preprocess", "calc_meters_to_px_ratio", "isolate_map_markers", "filter_map_marker_icons", "mask_marker_lines",
        "find_marker_lines"};
    return stage < SMH_STAGE_COUNT ? names[stage] : nullptr;
}

int32_t smh_b200_timeshares(smh_b200* h, float host_ms[SMH_STAGE_COUNT], float device_ms[SMH_STAGE_COUNT]) {
    if (!h) return SMH_E_ARG;
    if (host_ms) memcpy(host_ms, h->stage_host_ms, sizeof h->stage_host_ms);
    if (device_ms) memcpy(device_ms, h->stage_dev_ms, sizeof h->stage_dev_ms);
    return SMH_OK;
}

int32_t smh_b200_roi(uint32_t W, uint32_t H, uint32_t roi[12]) {
    if (!roi) return SMH_E_ARG;
    Geometry g;
    int rc = compute_geometry(W, H, &g);
    if (rc) return rc;
    memcpy(roi, g.btn, 16); memcpy(roi + 4, g.map, 16); memcpy(roi + 8, g.brq, 16);
    return SMH_OK;
}

int32_t smh_b200_create(int32_t device, smh_b200** out) {
    if (!out) return SMH_E_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, SMH_E_NOGPU, "no CUDA device (%s); this backend has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= count) return fail(nullptr, SMH_E_ARG, "device %d out of range (0..%d)", device, count - 1);
    smh_b200* h = new smh_b200();
    h->device = device;
#define CUC(expr)                                                                         \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            fail(nullptr, (int)e__, "%s failed: %s", #expr, cudaGetErrorString(e__));     \
            delete h;                                                                     \
            return (int)e__;                                                              \
        }                                                                                 \
    } while (0)
    CUC(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUC(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    if (prop.major != 10 || prop.minor != 0) {
        fail(nullptr, SMH_E_NOGPU, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
        delete h;
        return SMH_E_NOGPU;
    }
    CUC(cudaStreamCreateWithFlags(&h->s_main, cudaStreamNonBlocking));
    CUC(cudaStreamCreateWithFlags(&h->s_markers, cudaStreamNonBlocking));
    CUC(cudaStreamCreateWithFlags(&h->s_scales, cudaStreamNonBlocking));
    CUC(cudaEventCreateWithFlags(&h->ev_pixels, cudaEventDisableTiming));
    for (int s = 0; s < SMH_STAGE_COUNT; s++)
        for (int k = 0; k < 2; k++) CUC(cudaEventCreate(&h->stage_ev[s][k]));
    // ray directions: theta_k = (k as f32 / 10.0).to_radians(); (cos, sin) from the host libm so
    // that they are the bits vision-cpu computes (vision-cpu/src/lib.rs:398-399,437; D11)
    std::vector<float2> rays(SMH_LSD_RAYS);
    const float rads_per_deg = 3.14159265358979323846f / 180.0f;
    for (int k = 0; k < SMH_LSD_RAYS; k++) {
        const float theta = ((float)k / 10.0f) * rads_per_deg;
        rays[k] = make_float2(cosf(theta), sinf(theta));
    }
    CUC(cudaMalloc(&h->d_rays, sizeof(float2) * SMH_LSD_RAYS));
    CUC(cudaMemcpy(h->d_rays, rays.data(), sizeof(float2) * SMH_LSD_RAYS, cudaMemcpyHostToDevice));
    // exact marker predicate memoised over all 2^24 colours (2 MB, stays in L2)
    CUC(cudaMalloc(&h->d_lut, (1u << 24) / 8));
    smh::k_build_marker_lut<<<(1u << 19) / 256, 256, 0, h->s_main>>>(h->d_lut);
    h->launches++;
    CUC(cudaGetLastError());
    CUC(cudaStreamSynchronize(h->s_main));
    CUC(cudaMallocHost(&h->h_small, 256 * sizeof(float)));
    CUC(cudaMalloc(&h->d_longest, 8 * sizeof(float)));
#undef CUC
    *out = h;
    return SMH_OK;
}

void smh_b200_destroy(smh_b200* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    free_single(h);
    h->batch.release(); for (auto& P : h->pipe) P.release();
    cudaFree(h->d_stage); cudaFree(h->d_pack); cudaFree(h->d_rays); cudaFree(h->d_ckpt); cudaFree(h->d_lut); cudaFree(h->d_icons); cudaFree(h->d_longest); cudaFree(h->d_view);
    cudaFreeHost(h->h_small); cudaFreeHost(h->h_view); cudaFreeHost(h->h_gray);
    for (auto e : h->ev_gray) if (e) cudaEventDestroy(e);
    for (auto s : h->pipe_streams) cudaStreamDestroy(s);
    for (auto e : h->pipe_events) cudaEventDestroy(e);
    for (auto s : h->lane_streams) cudaStreamDestroy(s);
    for (auto e : h->lane_events) cudaEventDestroy(e);
    if (h->s_main) cudaStreamDestroy(h->s_main);
    if (h->s_markers) cudaStreamDestroy(h->s_markers);
    if (h->s_scales) cudaStreamDestroy(h->s_scales);
    if (h->ev_pixels) cudaEventDestroy(h->ev_pixels);
    for (auto& pair : h->stage_ev) for (auto e : pair) if (e) cudaEventDestroy(e);
    delete h;
}

int32_t smh_b200_bind_thread(smh_b200* h) {
    if (!h) return SMH_E_ARG;
    CU(h, cudaSetDevice(h->device));
    return SMH_OK;
}

uint64_t smh_b200_launch_count(smh_b200* h) { return h ? h->launches.load() : 0; }

// ------------------------------------------------------------------------------------------
// per-frame trait methods
// ------------------------------------------------------------------------------------------

int32_t smh_b200_load_frame(smh_b200* h, const uint8_t* bgra, uint32_t W, uint32_t H, uint32_t pitch_bytes) {
    if (!h || !bgra || pitch_bytes < W * 4u) return fail(h, SMH_E_ARG, "load_frame: bad argument");
    StageTimer tm(h, SMH_STAGE_LOAD_FRAME);
    CU(h, cudaSetDevice(h->device));
    if (!h->have_frame || h->geo.W != W || h->geo.H != H) {
        // resolution change: rebuild every per-resolution buffer (vision-cpu/src/lib.rs:70-88,
        // vision-gpu/src/lib.rs:118-125); scales_preprocessed starts from zeros again
        Geometry g;
        int rc = compute_geometry(W, H, &g);
        if (rc) return fail(h, rc, "load_frame: %ux%u is a frame size the reference panics on (crop out of bounds)", W, H);
        CU(h, cudaDeviceSynchronize());
        free_single(h);
        h->have_frame = false;
        h->geo = g;
        if ((rc = ensure_planes(h, &h->one, 1, g, true))) return rc;
        // uploaded region: columns [min(map_x, btn_x) & ~3, W), rows [min y, max(map bottom, button bottom))
        h->ux0 = (g.map[0] < g.btn[0] ? g.map[0] : g.btn[0]) & ~3u;
        h->uy0 = g.map[1] < g.btn[1] ? g.map[1] : g.btn[1];
        const uint32_t yb = g.btn[1] + g.btn[3], ym = g.map[1] + g.map[3];
        h->uy1 = yb > ym ? yb : ym;
        h->upitch = round_up((W - h->ux0) * 4u, 128u);
        CU(h, cudaMalloc(&h->d_frame, (size_t)h->upitch * (h->uy1 - h->uy0)));
        const size_t npx = (size_t)g.map[2] * g.map[3], nq = (size_t)g.brq[2] * g.brq[3];
        CU(h, cudaMallocHost(&h->h_ui, npx * 4));
        CU(h, cudaMallocHost(&h->h_lsd, npx));
        CU(h, cudaMallocHost(&h->h_ocr, nq));
        CU(h, cudaMallocHost(&h->h_scales, nq));
        memset(h->h_scales, 0, nq);
        CU(h, cudaMalloc(&h->d_scales_keep, nq));
        CU(h, cudaMemset(h->d_scales_keep, 0, nq));
    }
    const Geometry& g = h->geo;
    tm.begin(h->s_main);
    CU(h, cudaMemcpy2DAsync(h->d_frame, h->upitch, bgra + (size_t)h->uy0 * pitch_bytes + (size_t)h->ux0 * 4u, pitch_bytes,
                            (size_t)(g.W - h->ux0) * 4u, h->uy1 - h->uy0, cudaMemcpyHostToDevice, h->s_main));
    CU(h, cudaStreamSynchronize(h->s_main));  // the caller may drop the frame now
    tm.end();
    h->have_frame = true; h->cropped = false; h->map_open = false;
    h->ocr_fetched = h->lsd_fetched = false;
    return SMH_OK;
}

int32_t smh_b200_crop_to_map(smh_b200* h, int32_t grayscale, int32_t* map_open, uint32_t xywh[4], const uint8_t** ui_map_rgba) {
    if (!h || !map_open) return fail(h, SMH_E_ARG, "crop_to_map: bad argument");
    if (!h->have_frame) return fail(h, SMH_E_STATE, "crop_to_map before load_frame");
    StageTimer tm(h, SMH_STAGE_CROP_TO_MAP);
    CU(h, cudaSetDevice(h->device));
    const Geometry& g = h->geo;
    tm.begin(h->s_main);
    // the fresh thresholded scales image goes to a scratch plane; find_scales_preprocess merges
    // rows >= start_y into the persistent host image, so pass start_y = 0 here
    int rc = launch_pixels(h, h->one, h->d_frame, 0, h->upitch, h->uy1 - h->uy0, h->ux0, h->uy0, 1, false, 0, grayscale, h->s_main);
    if (rc) return rc;
    CU(h, cudaEventRecord(h->ev_pixels, h->s_main));
    uint32_t* h_red = reinterpret_cast<uint32_t*>(h->h_small + 200);
    CU(h, cudaMemcpyAsync(h_red, h->one.red, 4, cudaMemcpyDeviceToHost, h->s_main));
    CU(h, cudaMemcpyAsync(h_red + 1, h->one.open, 1, cudaMemcpyDeviceToHost, h->s_main));
    if ((rc = download_plane(h, &h->one.ui, 0, 1, h->h_ui, h->s_main))) return rc;
    CU(h, cudaStreamSynchronize(h->s_main));
    tm.end();
    h->red_pixels = h_red[0];
    h->map_open = (reinterpret_cast<uint8_t*>(h_red + 1))[0] != 0;
    h->grayscale = grayscale;
    h->cropped = true; h->ocr_fetched = h->lsd_fetched = false; h->icons_applied = false;
    *map_open = h->map_open ? 1 : 0;
    if (xywh) memcpy(xywh, g.map, 16);
    if (ui_map_rgba) *ui_map_rgba = h->map_open ? h->h_ui : nullptr;
    return SMH_OK;
}

int32_t smh_b200_red_pixels(smh_b200* h, uint32_t* count) {
    if (!h || !count) return SMH_E_ARG;
    if (!h->cropped) return fail(h, SMH_E_STATE, "red_pixels before crop_to_map");
    *count = h->red_pixels;
    return SMH_OK;
}

static int need_open_map(smh_b200* h, const char* who) {
    if (!h) return SMH_E_ARG;
    if (!h->cropped) return fail(h, SMH_E_STATE, "%s before crop_to_map", who);
    if (!h->map_open) return fail(h, SMH_E_STATE, "%s: map is closed (crop_to_map returned None)", who);
    return SMH_OK;
}

int32_t smh_b200_ocr_preprocess(smh_b200* h, const uint8_t** out, size_t* len) {
    int rc = need_open_map(h, "ocr_preprocess");
    if (rc) return rc;
    std::lock_guard<std::mutex> g(h->scales_mu);
    StageTimer tm(h, SMH_STAGE_OCR_PREPROCESS);
    CU(h, cudaSetDevice(h->device));
    if (!h->ocr_fetched) {
        CU(h, cudaStreamWaitEvent(h->s_scales, h->ev_pixels, 0));
        tm.begin(h->s_scales);
        if ((rc = download_plane(h, &h->one.ocr, 0, 1, h->h_ocr, h->s_scales))) return rc;
        CU(h, cudaStreamSynchronize(h->s_scales));
        tm.end();
        h->ocr_fetched = true;
    }
    if (out) *out = h->h_ocr;
    if (len) *len = (size_t)h->geo.brq[2] * h->geo.brq[3];
    return SMH_OK;
}

int32_t smh_b200_find_scales_preprocess(smh_b200* h, uint32_t start_y, const uint8_t** out, uint32_t wh[2]) {
    int rc = need_open_map(h, "find_scales_preprocess");
    if (rc) return rc;
    const uint32_t qw = h->geo.brq[2], qh = h->geo.brq[3];
    if (start_y > qh) return fail(h, SMH_E_RANGE, "find_scales_preprocess: scales_start_y %u > height %u", start_y, qh);
    std::lock_guard<std::mutex> g(h->scales_mu);
    StageTimer tm(h, SMH_STAGE_FIND_SCALES_PREPROCESS);
    CU(h, cudaSetDevice(h->device));
    if (start_y < qh) {
        // rows >= start_y come from this frame; rows above keep the previous content (D9)
        const smh_b200_plane& p = h->one.scales;
        CU(h, cudaStreamWaitEvent(h->s_scales, h->ev_pixels, 0));
        tm.begin(h->s_scales);
        CU(h, cudaMemcpy2DAsync(h->h_scales + (size_t)start_y * qw, qw, p.ptr + (size_t)start_y * p.pitch + p.pad, p.pitch,
                                qw, qh - start_y, cudaMemcpyDeviceToHost, h->s_scales));
        CU(h, cudaMemcpy2DAsync(h->d_scales_keep + (size_t)start_y * qw, qw, p.ptr + (size_t)start_y * p.pitch + p.pad, p.pitch,
                                qw, qh - start_y, cudaMemcpyDeviceToDevice, h->s_scales));
        CU(h, cudaStreamSynchronize(h->s_scales));
        tm.end();
    }
    if (out) *out = h->h_scales;
    if (wh) { wh[0] = qw; wh[1] = qh; }
    return SMH_OK;
}

int32_t smh_b200_isolate_map_markers(smh_b200* h) {
    int rc = need_open_map(h, "isolate_map_markers");
    if (rc) return rc;
    StageTimer tm(h, SMH_STAGE_ISOLATE_MAP_MARKERS);
    // the marker predicate was evaluated inside the fused pixel pass; nothing left to launch
    return SMH_OK;
}

int32_t smh_b200_load_map_markers(smh_b200* h, const uint8_t* rgba, uint32_t n_markers, uint32_t size) {
    if (!h || !rgba) return fail(h, SMH_E_ARG, "load_map_markers: bad argument");
    if (n_markers == 0 || n_markers > (uint32_t)smh::ICON_MAX_TEMPLATES || size < 4 || size > 32 ||
        (uint64_t)n_markers * size * size > 8192u)
        return fail(h, SMH_E_RANGE, "load_map_markers: %u templates of %ux%u (at most %d templates, 4 <= size <= 32, n*size*size <= 8192)",
                    n_markers, size, size, smh::ICON_MAX_TEMPLATES);
    std::lock_guard<std::mutex> g(h->markers_mu);
    CU(h, cudaSetDevice(h->device));
    // per (template pixel p = my*size + mx, template t): colour, alpha/255 and the score against a black pixel —
    // ((ad as f32) * (alpha as f32 / 255.0)) as u32, vision-cpu/src/lib.rs:323; per (t, my) the opaque columns
    const size_t np = (size_t)size * size, ne = np * n_markers;
    std::vector<uint32_t> blob(3 * ne + n_markers + (size_t)n_markers * size, 0u);
    uint32_t* rgb = blob.data(); float* af = reinterpret_cast<float*>(rgb + ne);
    uint32_t* c0 = reinterpret_cast<uint32_t*>(af + ne); uint32_t* base = c0 + ne; uint32_t* opq = base + n_markers;
    for (size_t p = 0; p < np; p++)
        for (uint32_t t = 0; t < n_markers; t++) {
            const uint8_t* px = rgba + ((size_t)t * np + p) * 4;
            const volatile float a = (float)px[3] / 255.0f;
            const volatile float v = (float)((uint32_t)px[0] + px[1] + px[2]) * a;
            rgb[p * n_markers + t] = (uint32_t)px[2] | ((uint32_t)px[1] << 8) | ((uint32_t)px[0] << 16);  // b | g<<8 | r<<16
            af[p * n_markers + t] = a;
            c0[p * n_markers + t] = (uint32_t)v;
            base[t] += (uint32_t)v;
            if (px[3] != 0) opq[(size_t)t * size + p / size] |= 1u << (p % size);
        }
    CU(h, cudaDeviceSynchronize());  // no kernel may still read the previous table
    cudaFree(h->d_icons); h->d_icons = nullptr;
    CU(h, cudaMalloc(&h->d_icons, blob.size() * 4));
    CU(h, cudaMemcpy(h->d_icons, blob.data(), blob.size() * 4, cudaMemcpyHostToDevice));
    h->icons.rgb = h->d_icons; h->icons.af = reinterpret_cast<const float*>(h->d_icons + ne);
    h->icons.c0 = h->d_icons + 2 * ne; h->icons.base = h->d_icons + 3 * ne; h->icons.opq = h->d_icons + 3 * ne + n_markers;
    h->icons.n = n_markers; h->icons.size = size;
    return SMH_OK;
}

int32_t smh_b200_filter_map_marker_icons(smh_b200* h, int32_t* found, uint32_t xy[2], uint32_t* sad) {
    int rc = need_open_map(h, "filter_map_marker_icons");
    if (rc) return rc;
    if (!found) return fail(h, SMH_E_ARG, "filter_map_marker_icons: null output");
    std::lock_guard<std::mutex> g(h->markers_mu);
    if (!h->d_icons) return fail(h, SMH_E_STATE, "filter_map_marker_icons before load_map_markers");
    if (h->icons_applied) return fail(h, SMH_E_STATE, "filter_map_marker_icons: already applied to this frame");
    StageTimer tm(h, SMH_STAGE_FILTER_MAP_MARKER_ICONS);
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamWaitEvent(h->s_markers, h->ev_pixels, 0));
    tm.begin(h->s_markers);
    if ((rc = launch_icons(h, h->one, h->d_frame, 0, h->upitch, h->ux0, h->uy0, 1, false, h->s_markers))) return rc;
    uint32_t* hs = reinterpret_cast<uint32_t*>(h->h_small + 248);  // (the scales branch owns 216..243)
    CU(h, cudaMemcpyAsync(hs, h->one.icon_xy, 8, cudaMemcpyDeviceToHost, h->s_markers));
    CU(h, cudaMemcpyAsync(hs + 2, h->one.icon_sad, 4, cudaMemcpyDeviceToHost, h->s_markers));
    CU(h, cudaMemcpyAsync(hs + 3, h->one.icon_found, 1, cudaMemcpyDeviceToHost, h->s_markers));
    CU(h, cudaStreamSynchronize(h->s_markers));
    tm.end();
    h->icons_applied = true; h->lsd_fetched = false;
    h->icon_found = reinterpret_cast<uint8_t*>(hs + 3)[0] ? 1 : 0;
    h->icon_xy[0] = hs[0]; h->icon_xy[1] = hs[1];
    *found = h->icon_found;
    if (xy) { xy[0] = hs[0]; xy[1] = hs[1]; }
    if (sad) *sad = hs[2];
    return SMH_OK;
}

int32_t smh_b200_mask_marker_lines(smh_b200* h) {
    int rc = need_open_map(h, "mask_marker_lines");
    if (rc) return rc;
    std::lock_guard<std::mutex> g(h->markers_mu);
    StageTimer tm(h, SMH_STAGE_MASK_MARKER_LINES);
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamWaitEvent(h->s_markers, h->ev_pixels, 0));  // the dilated mask is ready after this
    return SMH_OK;
}

int32_t smh_b200_lsd_image(smh_b200* h, const uint8_t** l8, uint32_t wh[2]) {
    int rc = need_open_map(h, "lsd_image");
    if (rc) return rc;
    std::lock_guard<std::mutex> g(h->markers_mu);
    CU(h, cudaSetDevice(h->device));
    if (!h->lsd_fetched) {
        CU(h, cudaStreamWaitEvent(h->s_markers, h->ev_pixels, 0));
        if ((rc = download_plane(h, &h->one.lsd, 0, 1, h->h_lsd, h->s_markers))) return rc;
        CU(h, cudaStreamSynchronize(h->s_markers));
        h->lsd_fetched = true;
    }
    if (l8) *l8 = h->h_lsd;
    if (wh) { wh[0] = h->geo.map[2]; wh[1] = h->geo.map[3]; }
    return SMH_OK;
}

int32_t smh_b200_find_longest_line(smh_b200* h, float x, float y, float max_gap, float line[4], float* len_sq) {
    int rc = need_open_map(h, "find_longest_line");
    if (rc) return rc;
    if (!line || !len_sq) return fail(h, SMH_E_ARG, "find_longest_line: null output");
    std::lock_guard<std::mutex> g(h->markers_mu);
    CU(h, cudaSetDevice(h->device));
    const smh_b200_plane& p = h->one.lsd;
    CU(h, cudaStreamWaitEvent(h->s_markers, h->ev_pixels, 0));
    smh::k_longest<<<1, smh::LSD_THREADS, 0, h->s_markers>>>(p.ptr, p.pitch, p.pad, p.width, p.height, h->d_rays, x, y,
                                                            max_gap, h->d_longest);
    h->launches++;
    CU(h, cudaGetLastError());
    float* hs = h->h_small + 160;
    CU(h, cudaMemcpyAsync(hs, h->d_longest, 5 * sizeof(float), cudaMemcpyDeviceToHost, h->s_markers));
    CU(h, cudaStreamSynchronize(h->s_markers));
    memcpy(line, hs, 16);
    *len_sq = hs[4];
    return SMH_OK;
}

// the line search of the single-frame path in two halves, so that the orchestrator can overlap the
// scales branch with it: enqueue on s_markers (no wait), then collect
static int marker_lines_enqueue(smh_b200* h, uint32_t max_gap) {
    int rc;
    CU(h, cudaStreamWaitEvent(h->s_markers, h->ev_pixels, 0));
    CU(h, cudaEventRecord(h->stage_ev[SMH_STAGE_FIND_MARKER_LINES][0], h->s_markers));
    if ((rc = launch_lsd(h, h->one, 1, (float)max_gap, false, false, true, h->s_markers))) return rc;
    CU(h, cudaMemcpyAsync(h->h_small, h->one.lines, 128 * sizeof(float), cudaMemcpyDeviceToHost, h->s_markers));
    CU(h, cudaMemcpyAsync(h->h_small + 128, h->one.n_lines, 4, cudaMemcpyDeviceToHost, h->s_markers));
    CU(h, cudaEventRecord(h->stage_ev[SMH_STAGE_FIND_MARKER_LINES][1], h->s_markers));
    return SMH_OK;
}
static int marker_lines_collect(smh_b200* h, float lines[32][4], uint32_t* n) {
    CU(h, cudaStreamSynchronize(h->s_markers));
    {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, h->stage_ev[SMH_STAGE_FIND_MARKER_LINES][0], h->stage_ev[SMH_STAGE_FIND_MARKER_LINES][1]) == cudaSuccess)
            h->stage_dev_ms[SMH_STAGE_FIND_MARKER_LINES] = ms;
    }
    const uint32_t cnt = *reinterpret_cast<uint32_t*>(h->h_small + 128);
    *n = cnt;
    memcpy(lines, h->h_small, (size_t)cnt * 16);
    return SMH_OK;
}

int32_t smh_b200_find_marker_lines(smh_b200* h, uint32_t max_gap, float lines[32][4], uint32_t* n) {
    int rc = need_open_map(h, "find_marker_lines");
    if (rc) return rc;
    if (!lines || !n) return fail(h, SMH_E_ARG, "find_marker_lines: null output");
    std::lock_guard<std::mutex> g(h->markers_mu);
    const auto t0 = std::chrono::steady_clock::now();
    CU(h, cudaSetDevice(h->device));
    if ((rc = marker_lines_enqueue(h, max_gap))) return rc;
    rc = marker_lines_collect(h, lines, n);
    h->stage_host_ms[SMH_STAGE_FIND_MARKER_LINES] = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

int32_t smh_b200_find_minimap(smh_b200* h, int32_t* found, uint32_t rect[4]) {
    if (!h || !found || !rect) return fail(h, SMH_E_ARG, "find_minimap: bad argument");
    if (!h->have_frame) return fail(h, SMH_E_STATE, "find_minimap before load_frame");
    StageTimer tm(h, SMH_STAGE_FIND_MINIMAP);
    CU(h, cudaSetDevice(h->device));
    tm.begin(h->s_main);
    int rc;
    if ((rc = launch_minimap(h, h->one, h->d_frame, 0, h->upitch, h->ux0, h->uy0, 1, false, h->s_main))) return rc;
    uint32_t* hs = reinterpret_cast<uint32_t*>(h->h_small + 208);
    CU(h, cudaMemcpyAsync(hs, h->one.minimap, 16, cudaMemcpyDeviceToHost, h->s_main));
    CU(h, cudaMemcpyAsync(hs + 4, h->one.minimap_found, 1, cudaMemcpyDeviceToHost, h->s_main));
    CU(h, cudaStreamSynchronize(h->s_main));
    tm.end();
    *found = reinterpret_cast<uint8_t*>(hs + 4)[0] ? 1 : 0;
    memcpy(rect, hs, 16);
    return SMH_OK;
}

int32_t smh_b200_calc_meters_to_px_ratio(smh_b200* h, const smh_b200_scale* scales, uint32_t n_scales, int32_t* found,
                                         double* ratio, uint32_t debug_lines[3][4], uint32_t* n_debug) {
    int rc = need_open_map(h, "calc_meters_to_px_ratio");
    if (rc) return rc;
    if (!found || !ratio || (n_scales && !scales) || n_scales > 3) return fail(h, SMH_E_ARG, "calc_meters_to_px_ratio: bad argument");
    std::lock_guard<std::mutex> g(h->scales_mu);
    StageTimer tm(h, SMH_STAGE_CALC_METERS_TO_PX_RATIO);
    CU(h, cudaSetDevice(h->device));
    *found = 0;
    if (n_debug) *n_debug = 0;
    if (n_scales == 0) return SMH_OK;  // `if scales.is_empty() { return None }`
    tm.begin(h->s_scales);
    // pinned scratch: [216..225) scales in, then the results
    uint32_t* in = reinterpret_cast<uint32_t*>(h->h_small + 216);
    for (uint32_t k = 0; k < n_scales; k++) { in[3 * k] = scales[k].meters; in[3 * k + 1] = scales[k].x; in[3 * k + 2] = scales[k].y; }
    uint8_t* nin = reinterpret_cast<uint8_t*>(h->h_small + 226);
    nin[0] = (uint8_t)n_scales;
    CU(h, cudaMemcpyAsync(h->one.mpx_in, in, 36, cudaMemcpyHostToDevice, h->s_scales));
    CU(h, cudaMemcpyAsync(h->one.mpx_n, nin, 1, cudaMemcpyHostToDevice, h->s_scales));
    const uint32_t qw = h->geo.brq[2];
    if ((rc = launch_mpx(h, h->one, h->d_scales_keep, 0, qw, 0, 1, h->s_scales))) return rc;
    double* hr = reinterpret_cast<double*>(h->h_small + 228);
    uint32_t* hd = reinterpret_cast<uint32_t*>(h->h_small + 230);
    uint8_t* hf = reinterpret_cast<uint8_t*>(h->h_small + 242);
    CU(h, cudaMemcpyAsync(hr, h->one.mpx_ratio, 8, cudaMemcpyDeviceToHost, h->s_scales));
    CU(h, cudaMemcpyAsync(hd, h->one.mpx_debug, 48, cudaMemcpyDeviceToHost, h->s_scales));
    CU(h, cudaMemcpyAsync(hf, h->one.mpx_found, 1, cudaMemcpyDeviceToHost, h->s_scales));
    CU(h, cudaMemcpyAsync(hf + 1, h->one.mpx_ndebug, 1, cudaMemcpyDeviceToHost, h->s_scales));
    CU(h, cudaStreamSynchronize(h->s_scales));
    tm.end();
    *found = hf[0] ? 1 : 0;
    *ratio = hr[0];
    if (debug_lines) memcpy(debug_lines, hd, (size_t)hf[1] * 16);
    if (n_debug) *n_debug = hf[1];
    return SMH_OK;
}

int32_t smh_b200_process(smh_b200* h, const uint8_t* bgra, uint32_t W, uint32_t H, uint32_t pitch_bytes, int32_t grayscale,
                         int32_t detect_markers, uint32_t max_gap, smh_b200_ocr_fn ocr, void* user, smh_b200_vision_results* out) {
    if (!h || !out) return fail(h, SMH_E_ARG, "process: bad argument");
    memset(out, 0, sizeof *out);
    int rc;
    if ((rc = smh_b200_load_frame(h, bgra, W, H, pitch_bytes))) return rc;                        // mod.rs:66
    if ((rc = smh_b200_crop_to_map(h, grayscale, &out->map_open, out->xywh, &out->ui_map_rgba))) return rc;  // mod.rs:69-75
    if (!out->map_open) return SMH_OK;                                                             // Ok(None)
    if ((rc = smh_b200_find_minimap(h, &out->has_minimap, out->minimap))) return rc;               // mod.rs:85
    // marker branch (mod.rs:87-118): the search is enqueued now and collected after the scales branch,
    // which is how the two rayon workers of the reference overlap (mod.rs:219-223)
    if (detect_markers) {
        if ((rc = smh_b200_isolate_map_markers(h))) return rc;
        if ((rc = smh_b200_mask_marker_lines(h))) return rc;
        std::lock_guard<std::mutex> g(h->markers_mu);
        const auto t0 = std::chrono::steady_clock::now();
        if ((rc = marker_lines_enqueue(h, max_gap))) return rc;
        h->stage_host_ms[SMH_STAGE_FIND_MARKER_LINES] = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    int scales_rc = SMH_OK;
    if (ocr) {                                                                                     // mod.rs:124-216
        const uint8_t* ocr_img = nullptr; size_t len = 0;
        scales_rc = smh_b200_ocr_preprocess(h, &ocr_img, &len);
        if (!scales_rc) {
            smh_b200_scale sc[3]; uint32_t ns = 0, start_y = 0xFFFFFFFFu;
            memset(sc, 0, sizeof sc);
            int32_t orc;
            { StageTimer tmo(h, SMH_STAGE_OCR); orc = ocr(user, ocr_img, h->geo.brq[2], h->geo.brq[3], sc, &ns, &start_y); }
            if (orc) scales_rc = fail(h, orc, "process: the ocr callback failed with %d", orc);
            else if (ns > 0 && start_y != 0xFFFFFFFFu) {
                if (ns > 3) ns = 3;
                const uint8_t* simg = nullptr; uint32_t wh[2];
                scales_rc = smh_b200_find_scales_preprocess(h, start_y, &simg, wh);
                if (!scales_rc)
                    scales_rc = smh_b200_calc_meters_to_px_ratio(h, sc, ns, &out->has_ratio, &out->meters_to_px_ratio,
                                                                 out->scales_debug, &out->n_scales_debug);
            }
        }
    }
    if (detect_markers) {
        std::lock_guard<std::mutex> g(h->markers_mu);
        const auto t0 = std::chrono::steady_clock::now();
        if ((rc = marker_lines_collect(h, out->markers, &out->n_markers))) return rc;
        // (enqueue + collect: the search itself ran under the scales branch)
        h->stage_host_ms[SMH_STAGE_FIND_MARKER_LINES] += std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return scales_rc;
}

int32_t smh_b200_debug_view(smh_b200* h, uint8_t which, const uint8_t** rgba, uint32_t wh[2]) {
    if (!h || !rgba) return fail(h, SMH_E_ARG, "debug_view: bad argument");
    *rgba = nullptr;
    if (which == SMH_VIEW_NONE) return SMH_OK;
    int rc = need_open_map(h, "debug_view");
    if (rc) return rc;
    if (which > SMH_VIEW_CROPPED_BRQ) return fail(h, SMH_E_ARG, "debug_view: unknown view %u", which);
    const Geometry& g = h->geo;
    const bool full = which == SMH_VIEW_LSD_PREPROCESS || which == SMH_VIEW_LSD_INPUT;
    const uint32_t w = full ? g.map[2] : g.brq[2], hh = full ? g.map[3] : g.brq[3];
    const size_t need = (size_t)w * hh * 4;
    std::lock_guard<std::mutex> gm(h->markers_mu);
    std::lock_guard<std::mutex> gs(h->scales_mu);
    CU(h, cudaSetDevice(h->device));
    if (h->view_cap < need) {
        cudaFreeHost(h->h_view); cudaFree(h->d_view); h->h_view = nullptr; h->d_view = nullptr; h->view_cap = 0;
        CU(h, cudaMallocHost(&h->h_view, need));
        CU(h, cudaMalloc(&h->d_view, need));
        h->view_cap = need;
    }
    CU(h, cudaStreamSynchronize(h->s_main));
    if (which == SMH_VIEW_LSD_PREPROCESS || which == SMH_VIEW_CROPPED_BRQ) {
        // colour views are re-derived from the resident capture (rare path)
        const uint32_t x0 = (full ? g.map[0] : g.brq[0]) - h->ux0, y0 = (full ? g.map[1] : g.brq[1]) - h->uy0;
        dim3 grid((w + 255) / 256, hh);
        smh::k_debug_rgb<<<grid, 256, 0, h->s_main>>>(h->d_frame, h->upitch, x0, y0, w, hh, full ? 1 : 0, h->d_view);
        h->launches++;
        CU(h, cudaGetLastError());
        if (full && h->icons_applied && h->icon_found) {  // the crop after filter_map_marker_icons (vision-cpu/src/lib.rs:333-352)
            smh::k_icon_overlay<<<1, 256, 0, h->s_main>>>(reinterpret_cast<uint32_t*>(h->d_view), w, hh, h->icon_xy[0], h->icon_xy[1], h->icons.size);
            h->launches++;
            CU(h, cudaGetLastError());
        }
        CU(h, cudaMemcpyAsync(h->h_view, h->d_view, need, cudaMemcpyDeviceToHost, h->s_main));
        CU(h, cudaStreamSynchronize(h->s_main));
    } else {
        // grey views: expand an L8 host image to (l,l,l,255) like image's ConvertBuffer
        const uint8_t* src;
        std::vector<uint8_t> tmp;
        if (which == SMH_VIEW_FIND_SCALES_INPUT) src = h->h_scales;
        else {
            const smh_b200_plane& p = which == SMH_VIEW_OCR_INPUT ? h->one.ocr : h->one.lsd;
            tmp.resize((size_t)w * hh);
            CU(h, cudaMemcpy2D(tmp.data(), w, p.ptr + p.pad, p.pitch, w, hh, cudaMemcpyDeviceToHost));
            src = tmp.data();
        }
        for (size_t i = 0; i < (size_t)w * hh; i++) {
            h->h_view[4 * i] = h->h_view[4 * i + 1] = h->h_view[4 * i + 2] = src[i];
            h->h_view[4 * i + 3] = 255;
        }
    }
    *rgba = h->h_view;
    if (wh) { wh[0] = w; wh[1] = hh; }
    return SMH_OK;
}

// ------------------------------------------------------------------------------------------
// batch extension
// ------------------------------------------------------------------------------------------

int32_t smh_b200_process_batch(smh_b200* h, const uint8_t* d_frames, uint32_t n, uint32_t W, uint32_t H,
                               uint32_t pitch_bytes, uint64_t frame_stride, const uint32_t* start_y,
                               uint32_t max_gap, uint32_t flags, void* stream, smh_b200_batch* out) {
    if (!h || !out || n == 0 || (!d_frames && !(flags & SMH_BATCH_SKIP_PIXELS)))
        return fail(h, SMH_E_ARG, "process_batch: bad argument");
    CU(h, cudaSetDevice(h->device));
    Geometry g;
    int rc = compute_geometry(W, H, &g);
    if (rc) return fail(h, rc, "process_batch: %ux%u is a frame size the reference panics on", W, H);
    cudaStream_t st = (cudaStream_t)stream;
    const bool stats = flags & SMH_BATCH_LSD_STATS;
    const bool icons = flags & SMH_BATCH_FILTER_ICONS;
    if (icons && (flags & SMH_BATCH_SKIP_PIXELS))
        return fail(h, SMH_E_ARG, "process_batch: FILTER_ICONS edits the mask of this call's pixel pass; not with SKIP_PIXELS");
    if (icons && !h->d_icons) return fail(h, SMH_E_STATE, "process_batch: FILTER_ICONS before load_map_markers");
    if ((flags & SMH_BATCH_SKIP_PIXELS) && (h->batch.n != n || h->batch.W != W || h->batch.H != H))
        return fail(h, SMH_E_STATE, "process_batch: SKIP_PIXELS needs a previous batch of the same shape");
    if ((rc = ensure_planes(h, &h->batch, n, g, stats))) return rc;
    if (!(flags & SMH_BATCH_SKIP_PIXELS) && start_y) {
        for (uint32_t i = 0; i < n; i++)
            if (start_y[i] > g.brq[3]) return fail(h, SMH_E_RANGE, "process_batch: start_y[%u]=%u > %u", i, start_y[i], g.brq[3]);
        CU(h, cudaMemcpyAsync(h->batch.start_y, start_y, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    }
    // Large batches: frames are independent, so the batch runs as chunks taken in turn by a few internal
    // streams (pixel pass -> bitmap -> line search per chunk).  The line search is latency-bound and its
    // persistent CTAs leave issue slots and, in the tail of every wave, whole SMs idle; the HBM-bound
    // pixel pass of the following chunks fills them.  Same kernels, same results; the caller's stream
    // is forked and joined with events, so the call stays asynchronous and stream-ordered.
    uint32_t chunk = 0, lanes = 0;
    batch_tuning(h, &chunk, &lanes);
    const bool whole = (flags & (SMH_BATCH_SKIP_PIXELS | SMH_BATCH_SKIP_LSD | SMH_BATCH_SERIAL)) || n < 2u * chunk || lanes < 2;
    if (whole) {
        if (!(flags & SMH_BATCH_SKIP_PIXELS)) {
            if ((rc = launch_pixels(h, h->batch, d_frames, frame_stride, pitch_bytes, H, 0, 0, n, start_y != nullptr, 0,
                                    flags & SMH_BATCH_GRAYSCALE, st)))
                return rc;
            if (flags & SMH_BATCH_MINIMAP)
                if ((rc = launch_minimap(h, h->batch, d_frames, frame_stride, pitch_bytes, 0, 0, n, true, st))) return rc;
        }
        if (icons) {
            if ((rc = launch_icons(h, h->batch, d_frames, frame_stride, pitch_bytes, 0, 0, n, true, st))) return rc;
        }
        if (!(flags & SMH_BATCH_SKIP_LSD))
            if ((rc = launch_lsd(h, h->batch, n, (float)max_gap, true, stats, !(flags & SMH_BATCH_LSD_EXHAUSTIVE), st))) return rc;
    } else {
        if ((rc = ensure_ray_ckpt(h, h->batch, st))) return rc;
        if (stats) CU(h, cudaMemsetAsync(h->batch.stats, 0, (size_t)n * 3 * sizeof(unsigned long long), st));
        while (h->lane_streams.size() < lanes) {
            cudaStream_t s;
            CU(h, cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
            h->lane_streams.push_back(s);
        }
        while (h->lane_events.size() < 1u + lanes) {
            cudaEvent_t e;
            CU(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            h->lane_events.push_back(e);
        }
        CU(h, cudaEventRecord(h->lane_events[0], st));
        for (uint32_t l = 0; l < lanes; l++) CU(h, cudaStreamWaitEvent(h->lane_streams[l], h->lane_events[0], 0));
        uint32_t c = 0;
        for (uint32_t f0 = 0; f0 < n; f0 += chunk, c++) {
            const uint32_t cn = n - f0 < chunk ? n - f0 : chunk, lane = c % lanes;
            cudaStream_t ls = h->lane_streams[lane];
            const Planes V = planes_view(h->batch, f0, cn, lane);
            const uint8_t* fr = d_frames + (uint64_t)f0 * frame_stride;
            if ((rc = launch_pixels(h, V, fr, frame_stride, pitch_bytes, H, 0, 0, cn, start_y != nullptr, 0,
                                    flags & SMH_BATCH_GRAYSCALE, ls)))
                return rc;
            if (flags & SMH_BATCH_MINIMAP)
                if ((rc = launch_minimap(h, V, fr, frame_stride, pitch_bytes, 0, 0, cn, true, ls))) return rc;
            if (icons) {
                if ((rc = launch_icons(h, V, fr, frame_stride, pitch_bytes, 0, 0, cn, true, ls))) return rc;
            }
            if ((rc = launch_lsd(h, V, cn, (float)max_gap, true, stats, !(flags & SMH_BATCH_LSD_EXHAUSTIVE), ls))) return rc;
        }
        for (uint32_t l = 0; l < lanes; l++) {
            CU(h, cudaEventRecord(h->lane_events[1 + l], h->lane_streams[l]));
            CU(h, cudaStreamWaitEvent(st, h->lane_events[1 + l], 0));
        }
    }
    fill_batch_out(h->batch, stats, (flags & SMH_BATCH_MINIMAP) != 0, icons, out);
    return SMH_OK;
}

int32_t smh_b200_batch_meters_to_px_ratio(smh_b200* h, const smh_b200_scale* scales, const uint8_t* n_scales, uint32_t n,
                                          void* stream, double* ratio, uint8_t* found) {
    if (!h || !scales || !n_scales || !ratio || !found || n == 0) return fail(h, SMH_E_ARG, "batch_meters_to_px_ratio: bad argument");
    if (h->batch.n != n) return fail(h, SMH_E_STATE, "batch_meters_to_px_ratio: needs a previous process_batch of %u frames", n);
    CU(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const Planes& P = h->batch;
    CU(h, cudaMemcpyAsync(P.mpx_in, scales, (size_t)n * 36, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(P.mpx_n, n_scales, (size_t)n, cudaMemcpyHostToDevice, st));
    int rc;
    if ((rc = launch_mpx(h, P, P.scales.ptr, P.scales.frame_stride, P.scales.pitch, P.scales.pad, n, st))) return rc;
    CU(h, cudaMemcpyAsync(ratio, P.mpx_ratio, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    CU(h, cudaMemcpyAsync(found, P.mpx_found, (size_t)n, cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    return SMH_OK;
}

int32_t smh_b200_download_plane(smh_b200* h, const smh_b200_plane* p, uint32_t first, uint32_t count, uint8_t* host_out, void* stream) {
    if (!h || !p || !host_out) return fail(h, SMH_E_ARG, "download_plane: bad argument");
    CU(h, cudaSetDevice(h->device));
    return download_plane(h, p, first, count, host_out, (cudaStream_t)stream);
}

int32_t smh_b200_stream_synchronize(smh_b200* h, void* stream) {
    if (!h) return SMH_E_ARG;
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamSynchronize((cudaStream_t)stream));
    return SMH_OK;
}

// ring_n = 0: frame f of the batch is frames[f] and its results land at index f.  ring_n > 0 (streaming with bounded
// buffers): frame f is frames[f % ring_n] and its results land at index f % ring_n of every result array.
static int32_t host_batch_impl(smh_b200* h, const uint8_t* frames, uint32_t n, uint32_t ring_n, uint32_t W, uint32_t H,
                               uint32_t pitch_bytes, uint64_t frame_stride, const uint32_t* start_y,
                               uint32_t max_gap, uint32_t flags, uint32_t chunk, const smh_b200_host_results* res) {
    if (!h || !frames || !res || n == 0 || !res->lines || !res->n_lines || !res->red_pixels || !res->map_open)
        return fail(h, SMH_E_ARG, "process_host_batch: bad argument");
    CU(h, cudaSetDevice(h->device));
    Geometry g;
    int rc = compute_geometry(W, H, &g);
    if (rc) return fail(h, rc, "process_host_batch: %ux%u is a frame size the reference panics on", W, H);
    if (chunk == 0) chunk = 32;
    if (chunk > n) chunk = n;
    if (ring_n && ring_n < (SMH_PIPE_DEPTH + 1u) * chunk)
        return fail(h, SMH_E_ARG, "process_host_stream: the ring must hold at least %u frames (%u chunks of %u in flight + one being refilled)",
                    (SMH_PIPE_DEPTH + 1u) * chunk, (unsigned)SMH_PIPE_DEPTH, chunk);
    const bool icons = flags & SMH_BATCH_FILTER_ICONS;
    if (icons && !h->d_icons) return fail(h, SMH_E_STATE, "process_host_batch: FILTER_ICONS before load_map_markers");
    constexpr int DEPTH = SMH_PIPE_DEPTH;
    // Three stages on three streams — upload, compute (kernels + row packing), download — chained by
    // events per slot, so the bus is busy in both directions all the time: chunk k+1 goes up and chunk
    // k-1 comes down while chunk k is computed.  A slot's upload staging is free again once its kernels
    // ran, its planes once they were packed, its packed rows once they were downloaded.
    if (h->pipe_streams.size() < 4) {
        const size_t have = h->pipe_streams.size();
        h->pipe_streams.resize(4);
        for (size_t i = have; i < 4; i++) CU(h, cudaStreamCreateWithFlags(&h->pipe_streams[i], cudaStreamNonBlocking));
    }
    if (h->pipe_events.size() < (size_t)(3 * DEPTH)) {
        const size_t have = h->pipe_events.size();
        h->pipe_events.resize(3 * DEPTH);
        for (size_t i = have; i < (size_t)(3 * DEPTH); i++) CU(h, cudaEventCreateWithFlags(&h->pipe_events[i], cudaEventDisableTiming));
    }
    // (two compute streams, taken in turn: the line search of a chunk is latency-bound — one CTA per frame —
    // and leaves most SMs idle, so consecutive chunks' kernels may overlap)
    cudaStream_t s_up = h->pipe_streams[0], s_run[2] = {h->pipe_streams[1], h->pipe_streams[2]}, s_down = h->pipe_streams[3];
    // Device staging keeps the capture's geometry for the region the path reads (columns [x0&~3, W),
    // rows [map_y, y1)), but only the two rectangles that are read cross the bus: the map rows, as wide
    // as the pixel pass's vector grid, and the button.
    const uint32_t ux0 = (g.map[0] < g.btn[0] ? g.map[0] : g.btn[0]) & ~3u;
    const uint32_t uy0 = g.map[1] < g.btn[1] ? g.map[1] : g.btn[1];
    const uint32_t yb = g.btn[1] + g.btn[3], ym = g.map[1] + g.map[3];
    const uint32_t uy1 = yb > ym ? yb : ym, rows = uy1 - uy0;
    const uint32_t upitch = round_up((W - ux0) * 4u, 128u);
    const uint64_t ustride = (uint64_t)upitch * rows;
    const size_t need = (size_t)ustride * chunk * DEPTH;
    if (h->stage_cap < need) {
        CU(h, cudaDeviceSynchronize());
        cudaFree(h->d_stage); h->d_stage = nullptr; h->stage_cap = 0;
        CU(h, cudaMalloc(&h->d_stage, need));
        CU(h, cudaMemset(h->d_stage, 0, need));
        h->stage_cap = need;
    }
    const size_t npx = (size_t)g.map[2] * g.map[3], nq = (size_t)g.brq[2] * g.brq[3];
    // Download staging: the planes' rows carry pitch padding on the device; a strided copy over the bus
    // pays per row (the quadrant planes' rows are a few hundred bytes), so the rows are first packed on the
    // device and each plane of a chunk then crosses the bus as ONE contiguous copy.
    // grayscale ui_map: every pixel is (L, L, L, 255) — only L crosses the bus, host threads widen it (HostPool)
    static const bool l8_ok = [] { const char* e = getenv("SMH_B200_UI_L8"); return !(e && e[0] == '0'); }();
    if ((flags & SMH_BATCH_UI_L8) && !(flags & SMH_BATCH_GRAYSCALE)) return fail(h, SMH_E_ARG, "process_host_batch: UI_L8 needs GRAYSCALE");
    const bool l8_out = res->ui_map && (flags & SMH_BATCH_UI_L8);           // the caller takes the L plane as it is
    const bool ui_l8 = res->ui_map && (flags & SMH_BATCH_GRAYSCALE) && (l8_ok || l8_out);  // the L plane crosses the bus
    if (ui_l8 && !l8_out) {
        if (h->pool.th.empty()) {
            // measured on a 16-core B200 host, 1024 x 1440p frames: 2 threads 6.5 k frames/s, 4 -> 7.8 k, 8 -> 7.5 k, 12 -> 7.4 k
            // (the RGBA rows over the bus instead: 7.25 k)
            unsigned hw = std::thread::hardware_concurrency(), nt = hw >= 8 ? 4 : 2;
            if (const char* e = getenv("SMH_B200_HOST_THREADS")) { const int v = atoi(e); if (v >= 1 && v <= 64) nt = (unsigned)v; }
            h->pool.start(h->device, nt);
            for (int i = 0; i < DEPTH; i++) {
                CU(h, cudaEventCreateWithFlags(&h->ev_gray[i], cudaEventDisableTiming | cudaEventBlockingSync));
                h->gray_pending[i] = 0;
            }
        }
        const size_t gneed = npx * chunk * DEPTH;
        if (h->gray_cap < gneed) {
            for (int i = 0; i < DEPTH; i++) h->pool.wait_zero(h->gray_pending[i]);
            CU(h, cudaDeviceSynchronize());
            cudaFreeHost(h->h_gray); h->h_gray = nullptr; h->gray_cap = 0;
            CU(h, cudaMallocHost(&h->h_gray, gneed));
            h->gray_cap = gneed;
        }
    }
    const size_t pack_frame = (res->ui_map ? (ui_l8 ? npx : npx * 4) : 0) + (res->ocr_out ? nq : 0) + (res->scales ? nq : 0) + (res->lsd_image ? npx : 0);
    const size_t pack_need = pack_frame * chunk * DEPTH;
    if (h->pack_cap < pack_need) {
        CU(h, cudaDeviceSynchronize());
        cudaFree(h->d_pack); h->d_pack = nullptr; h->pack_cap = 0;
        CU(h, cudaMalloc(&h->d_pack, pack_need));
        h->pack_cap = pack_need;
    }
    for (int b = 0; b < DEPTH; b++)
        if ((rc = ensure_planes(h, &h->pipe[b], chunk, g, false))) return rc;
    // map rectangle in frame coordinates: columns of the vector grid [gx0, gx0 + 4*nvec) clipped to the frame
    const uint32_t gx0 = g.map[0] & ~3u;
    const uint32_t mcols = (gx0 + 4u * h->pipe[0].nvec <= W ? 4u * h->pipe[0].nvec : W - gx0);
    // How the rectangles cross the bus: 1 = one 3-D copy per rectangle and chunk (depth = frames; needs
    // tightly stacked frames), 0 = one 2-D copy per frame.  Measured on B200/PCIe 5 with downloads running
    // the other way: 47 GB/s against 31 GB/s (and 46 GB/s for SM zero-copy loads, which however idle the
    // link while they wait for a slot's kernels).
    int h2d_mode = 1;
    if (const char* e = getenv("SMH_B200_H2D_MODE")) h2d_mode = atoi(e);
    // SMH_B200_PIPE_TRACE=1: event-stamp every chunk's upload / kernels / download and print the timeline
    static const bool trace = getenv("SMH_B200_PIPE_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    auto stamp = [&](cudaStream_t st) { if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tev.push_back(e); } };
    uint32_t done = 0;
    int b = 0;
    uint32_t kth = 0;
    while (done < n) {
        // chunk schedule: a short first and second chunk so that kernels and downloads start early, full
        // chunks in between, halving chunks at the end so that little is left to drain after the last upload
        const uint32_t left = n - done, quarter = chunk / 4u > 0u ? chunk / 4u : 1u;
        uint32_t c = kth == 0 ? quarter : (kth == 1 ? (chunk / 2u > 0u ? chunk / 2u : 1u) : chunk);
        if (left < 2u * chunk) { const uint32_t half = (left + 1u) / 2u; c = half > quarter ? half : quarter; }
        if (c > chunk) c = chunk;
        if (c > left) c = left;
        const uint32_t slot = ring_n ? done % ring_n : done;  // where the chunk's frames and results live on the host
        if (ring_n && c > ring_n - slot) c = ring_n - slot;   // a chunk never straddles the end of the ring
        kth++;
        Planes& P = h->pipe[b];
        uint8_t* stage = h->d_stage + (size_t)b * ustride * chunk;
        cudaEvent_t ev_up = h->pipe_events[3 * b], ev_run = h->pipe_events[3 * b + 1], ev_down = h->pipe_events[3 * b + 2];
        const bool reused = kth > (uint32_t)DEPTH;  // the slot carried chunk k - DEPTH
        // ---- upload: the staging slot is free once the previous occupant's kernels have run --------
        cudaStream_t st = s_up;
        if (reused) CU(h, cudaStreamWaitEvent(st, ev_run, 0));
        stamp(st);
        const uint8_t* src0 = frames + (uint64_t)slot * frame_stride;
        if (h2d_mode == 1 && frame_stride == (uint64_t)pitch_bytes * H) {
            // one 3-D copy per rectangle for the whole chunk (depth = frames)
            cudaMemcpy3DParms mp{};
            mp.srcPtr = make_cudaPitchedPtr((void*)src0, pitch_bytes, (size_t)W * 4u, H);
            mp.dstPtr = make_cudaPitchedPtr(stage, upitch, upitch, rows);
            mp.kind = cudaMemcpyHostToDevice;
            mp.srcPos = make_cudaPos((size_t)gx0 * 4u, g.map[1], 0);
            mp.dstPos = make_cudaPos((size_t)(gx0 - ux0) * 4u, g.map[1] - uy0, 0);
            mp.extent = make_cudaExtent((size_t)mcols * 4u, g.map[3], c);
            CU(h, cudaMemcpy3DAsync(&mp, st));
            mp.srcPos = make_cudaPos((size_t)g.btn[0] * 4u, g.btn[1], 0);
            mp.dstPos = make_cudaPos((size_t)(g.btn[0] - ux0) * 4u, g.btn[1] - uy0, 0);
            mp.extent = make_cudaExtent((size_t)g.btn[2] * 4u, g.btn[3], c);
            CU(h, cudaMemcpy3DAsync(&mp, st));
        } else {
            for (uint32_t i = 0; i < c; i++) {
                const uint8_t* src = src0 + (uint64_t)i * frame_stride;
                uint8_t* dst = stage + (uint64_t)i * ustride;
                CU(h, cudaMemcpy2DAsync(dst + (size_t)(g.map[1] - uy0) * upitch + (size_t)(gx0 - ux0) * 4u, upitch,
                                        src + (size_t)g.map[1] * pitch_bytes + (size_t)gx0 * 4u, pitch_bytes,
                                        (size_t)mcols * 4u, g.map[3], cudaMemcpyHostToDevice, st));
                CU(h, cudaMemcpy2DAsync(dst + (size_t)(g.btn[1] - uy0) * upitch + (size_t)(g.btn[0] - ux0) * 4u, upitch,
                                        src + (size_t)g.btn[1] * pitch_bytes + (size_t)g.btn[0] * 4u, pitch_bytes,
                                        (size_t)g.btn[2] * 4u, g.btn[3], cudaMemcpyHostToDevice, st));
            }
        }
        stamp(st);
        CU(h, cudaEventRecord(ev_up, st));
        // ---- compute: needs the upload; the packed rows of the slot must have gone down ---------------
        st = s_run[kth & 1u];
        CU(h, cudaStreamWaitEvent(st, ev_up, 0));
        if (reused) CU(h, cudaStreamWaitEvent(st, ev_down, 0));
        if (start_y) CU(h, cudaMemcpyAsync(P.start_y, start_y + done, (size_t)c * 4, cudaMemcpyHostToDevice, st));
        if ((rc = launch_pixels(h, P, stage, ustride, upitch, rows, ux0, uy0, c, start_y != nullptr, 0, flags & SMH_BATCH_GRAYSCALE, st)))
            return rc;
        if ((flags & SMH_BATCH_MINIMAP) && res->minimap && res->minimap_found)
            if ((rc = launch_minimap(h, P, stage, ustride, upitch, ux0, uy0, c, true, st))) return rc;
        if (icons) {
            if ((rc = launch_icons(h, P, stage, ustride, upitch, ux0, uy0, c, true, st))) return rc;
        }
        if (!(flags & SMH_BATCH_SKIP_LSD))
            if ((rc = launch_lsd(h, P, c, (float)max_gap, true, false, true, st))) return rc;
        uint8_t* const pack0 = h->d_pack + (size_t)b * pack_frame * chunk;
        struct Want { const smh_b200_plane* pl; uint8_t* host; };
        const Want wants[4] = {{res->ui_map ? &P.ui : nullptr, res->ui_map ? res->ui_map + (size_t)slot * npx * 4 : nullptr},
                               {res->ocr_out ? &P.ocr : nullptr, res->ocr_out ? res->ocr_out + (size_t)slot * nq : nullptr},
                               {res->scales ? &P.scales : nullptr, res->scales ? res->scales + (size_t)slot * nq : nullptr},
                               {res->lsd_image ? &P.lsd : nullptr, res->lsd_image ? res->lsd_image + (size_t)slot * npx : nullptr}};
        uint8_t* cursor = pack0;
        for (const Want& wnt : wants) {  // rows without their pitch padding, plane after plane
            if (!wnt.pl) continue;
            if (ui_l8 && wnt.pl == &P.ui) {
                const uint32_t nrows = P.ui.height * c;
                k_pack_gray<<<dim3((P.ui.width + 255u) / 256u, nrows < 4096u ? nrows : 4096u), 256, 0, st>>>(P.ui.ptr, P.ui.pitch, P.ui.pad, P.ui.width, nrows, cursor);
                h->launches++;
                CU(h, cudaGetLastError());
                cursor += npx * c;
                continue;
            }
            const size_t row = (size_t)wnt.pl->width * wnt.pl->bytes_per_pixel;
            CU(h, cudaMemcpy2DAsync(cursor, row, wnt.pl->ptr + (size_t)wnt.pl->pad * wnt.pl->bytes_per_pixel, wnt.pl->pitch, row,
                                    (size_t)wnt.pl->height * c, cudaMemcpyDeviceToDevice, st));
            cursor += row * wnt.pl->height * c;
        }
        stamp(st);
        CU(h, cudaEventRecord(ev_run, st));
        // ---- download: every plane of the chunk as one contiguous copy ------------------------------
        st = s_down;
        CU(h, cudaStreamWaitEvent(st, ev_run, 0));
        cursor = pack0;
        for (const Want& wnt : wants) {
            if (!wnt.pl) continue;
            if (l8_out && wnt.pl == &P.ui) {  // straight into the caller's L8 buffer
                CU(h, cudaMemcpyAsync(res->ui_map + (size_t)slot * npx, cursor, npx * c, cudaMemcpyDeviceToHost, st));
                cursor += npx * c;
                continue;
            }
            if (ui_l8 && wnt.pl == &P.ui) {
                // the slot's landing area is free once the workers widened its previous occupant
                h->pool.wait_zero(h->gray_pending[b]);
                uint8_t* land = h->h_gray + (size_t)b * npx * chunk;
                CU(h, cudaMemcpyAsync(land, cursor, npx * c, cudaMemcpyDeviceToHost, st));
                CU(h, cudaEventRecord(h->ev_gray[b], st));
                const size_t total = npx * c, nt = h->pool.th.size(), per = (total / nt + 63) & ~(size_t)63;
                int tasks = 0;
                for (size_t o = 0; o < total; o += per) tasks++;
                h->gray_pending[b] = tasks;
                for (size_t o = 0; o < total; o += per)
                    h->pool.push({h->ev_gray[b], land + o, reinterpret_cast<uint32_t*>(wnt.host) + o, total - o < per ? total - o : per,
                                  &h->gray_pending[b]});
                cursor += npx * c;
                continue;
            }
            const size_t bytes = (size_t)wnt.pl->width * wnt.pl->bytes_per_pixel * wnt.pl->height * c;
            CU(h, cudaMemcpyAsync(wnt.host, cursor, bytes, cudaMemcpyDeviceToHost, st));
            cursor += bytes;
        }
        CU(h, cudaMemcpyAsync(res->lines + (size_t)slot * 128, P.lines, (size_t)c * 128 * sizeof(float), cudaMemcpyDeviceToHost, st));
        CU(h, cudaMemcpyAsync(res->n_lines + slot, P.n_lines, (size_t)c * 4, cudaMemcpyDeviceToHost, st));
        CU(h, cudaMemcpyAsync(res->red_pixels + slot, P.red, (size_t)c * 4, cudaMemcpyDeviceToHost, st));
        CU(h, cudaMemcpyAsync(res->map_open + slot, P.open, (size_t)c, cudaMemcpyDeviceToHost, st));
        if ((flags & SMH_BATCH_MINIMAP) && res->minimap && res->minimap_found) {
            CU(h, cudaMemcpyAsync(res->minimap + (size_t)slot * 4, P.minimap, (size_t)c * 16, cudaMemcpyDeviceToHost, st));
            CU(h, cudaMemcpyAsync(res->minimap_found + slot, P.minimap_found, (size_t)c, cudaMemcpyDeviceToHost, st));
        }
        if (icons) {
            if (res->icon_found) CU(h, cudaMemcpyAsync(res->icon_found + slot, P.icon_found, (size_t)c, cudaMemcpyDeviceToHost, st));
            if (res->icon_xy) CU(h, cudaMemcpyAsync(res->icon_xy + (size_t)slot * 2, P.icon_xy, (size_t)c * 8, cudaMemcpyDeviceToHost, st));
            if (res->icon_sad) CU(h, cudaMemcpyAsync(res->icon_sad + slot, P.icon_sad, (size_t)c * 4, cudaMemcpyDeviceToHost, st));
        }
        stamp(st);
        CU(h, cudaEventRecord(ev_down, st));
        done += c;
        b = (b + 1) % DEPTH;
    }
    if (trace) {
        for (auto s : h->pipe_streams) cudaStreamSynchronize(s);
        for (size_t k = 0; k + 3 < tev.size(); k += 4) {
            float t[4];
            for (int j = 0; j < 4; j++) cudaEventElapsedTime(&t[j], tev[0], tev[k + j]);
            fprintf(stderr, "[pipe] chunk %2zu  up %7.3f-%7.3f  kernels -%7.3f  down -%7.3f ms\n", k / 4, t[0], t[1], t[2], t[3]);
        }
        for (auto e : tev) cudaEventDestroy(e);
    }
    for (auto s : h->pipe_streams) CU(h, cudaStreamSynchronize(s));
    if (ui_l8 && !l8_out) {
        for (int i = 0; i < DEPTH; i++) h->pool.wait_zero(h->gray_pending[i]);
        if (h->pool.failed.exchange(0)) return fail(h, SMH_E_STATE, "process_host_batch: a host worker failed to wait for its download");
    }
    return SMH_OK;
}

int32_t smh_b200_process_host_batch(smh_b200* h, const uint8_t* frames, uint32_t n, uint32_t W, uint32_t H,
                                    uint32_t pitch_bytes, uint64_t frame_stride, const uint32_t* start_y,
                                    uint32_t max_gap, uint32_t flags, uint32_t chunk, const smh_b200_host_results* res) {
    return host_batch_impl(h, frames, n, 0, W, H, pitch_bytes, frame_stride, start_y, max_gap, flags, chunk, res);
}

int32_t smh_b200_process_host_stream(smh_b200* h, const uint8_t* frames, uint32_t ring_n, uint32_t n, uint32_t W, uint32_t H,
                                     uint32_t pitch_bytes, uint64_t frame_stride, const uint32_t* start_y,
                                     uint32_t max_gap, uint32_t flags, uint32_t chunk, const smh_b200_host_results* res) {
    if (ring_n == 0) return fail(h, SMH_E_ARG, "process_host_stream: empty ring");
    return host_batch_impl(h, frames, n, ring_n, W, H, pitch_bytes, frame_stride, start_y, max_gap, flags, chunk, res);
}

// ------------------------------------------------------------------------------------------
// synthetic captures + colour tables
// ------------------------------------------------------------------------------------------

namespace {
struct Scene {
    std::vector<smh_synth_frame_desc> desc;
    std::vector<smh_synth_segment> segs;
    std::vector<smh_synth_blip> blips;
};
int build_scene(const smh_b200_synth_params* p, uint32_t first, uint32_t n, Scene* sc) {
    Geometry g;
    int rc = compute_geometry(p->W, p->H, &g);
    if (rc) return rc;
    if (p->n_labels > SMH_SYNTH_MAX_LABELS) return SMH_E_ARG;
    static_assert(sizeof(smh_synth_params) == sizeof(smh_b200_synth_params), "the generator's parameters are the ABI's");
    sc->desc.resize(n); sc->segs.resize((size_t)n * p->n_segments); sc->blips.resize((size_t)n * p->n_blips);
    uint32_t roi[12];
    memcpy(roi, g.btn, 16); memcpy(roi + 4, g.map, 16); memcpy(roi + 8, g.brq, 16);
    for (uint32_t f = 0; f < n; f++)
        smh_synth_scene_frame(reinterpret_cast<const smh_synth_params*>(p), roi, first + f, &sc->desc[f],
                              sc->segs.data() + (size_t)f * p->n_segments, sc->blips.data() + (size_t)f * p->n_blips);
    return SMH_OK;
}
}  // namespace

int32_t smh_b200_synth_frames_host(const smh_b200_synth_params* p, uint32_t first, uint32_t n, uint8_t* out) {
    if (!p || !out) return SMH_E_ARG;
    Scene sc;
    int rc = build_scene(p, first, n, &sc);
    if (rc) return rc;
    for (uint32_t f = 0; f < n; f++) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(out + (size_t)f * p->W * p->H * 4);
        const smh_synth_segment* segs = sc.segs.data() + (size_t)f * p->n_segments;
        const smh_synth_blip* blips = sc.blips.data() + (size_t)f * p->n_blips;
        for (uint32_t y = 0; y < p->H; y++)
            for (uint32_t x = 0; x < p->W; x++)
                dst[(size_t)y * p->W + x] = smh_synth_pixel(&sc.desc[f], segs, blips, (int32_t)x, (int32_t)y);
    }
    return SMH_OK;
}

int32_t smh_b200_synth_frames_device(smh_b200* h, const smh_b200_synth_params* p, uint32_t first, uint32_t n,
                                     uint8_t* d_out, void* stream) {
    if (!h || !p || !d_out || n == 0) return fail(h, SMH_E_ARG, "synth_frames_device: bad argument");
    CU(h, cudaSetDevice(h->device));
    Scene sc;
    int rc = build_scene(p, first, n, &sc);
    if (rc) return fail(h, rc, "synth_frames_device: bad parameters");
    cudaStream_t st = (cudaStream_t)stream;
    smh_synth_frame_desc* d_desc; smh_synth_segment* d_segs = nullptr; smh_synth_blip* d_blips = nullptr;
    CU(h, cudaMalloc(&d_desc, sizeof(smh_synth_frame_desc) * n));
    CU(h, cudaMemcpyAsync(d_desc, sc.desc.data(), sizeof(smh_synth_frame_desc) * n, cudaMemcpyHostToDevice, st));
    if (!sc.segs.empty()) {
        CU(h, cudaMalloc(&d_segs, sizeof(smh_synth_segment) * sc.segs.size()));
        CU(h, cudaMemcpyAsync(d_segs, sc.segs.data(), sizeof(smh_synth_segment) * sc.segs.size(), cudaMemcpyHostToDevice, st));
    }
    if (!sc.blips.empty()) {
        CU(h, cudaMalloc(&d_blips, sizeof(smh_synth_blip) * sc.blips.size()));
        CU(h, cudaMemcpyAsync(d_blips, sc.blips.data(), sizeof(smh_synth_blip) * sc.blips.size(), cudaMemcpyHostToDevice, st));
    }
    smh::SynthArgs a{d_desc, d_segs, d_blips, p->n_segments, p->n_blips, p->W, p->H, d_out};
    for (uint32_t f0 = 0; f0 < n; f0 += 32768) {  // gridDim.z limit
        const uint32_t c = n - f0 < 32768 ? n - f0 : 32768;
        smh::SynthArgs b = a;
        b.desc += f0; b.segs += (size_t)f0 * p->n_segments; b.blips += (size_t)f0 * p->n_blips;
        b.out += (size_t)f0 * p->W * p->H * 4;
        smh::k_synth<<<dim3((p->W + 255) / 256, p->H, c), 256, 0, st>>>(b);
        h->launches++;
    }
    CU(h, cudaGetLastError());
    CU(h, cudaStreamSynchronize(st));
    cudaFree(d_desc); cudaFree(d_segs); cudaFree(d_blips);
    return SMH_OK;
}

int32_t smh_b200_colour_table(smh_b200* h, uint32_t which, uint8_t* host_out) {
    if (!h || !host_out || which > 6) return fail(h, SMH_E_ARG, "colour_table: bad argument");
    CU(h, cudaSetDevice(h->device));
    uint8_t* d;
    CU(h, cudaMalloc(&d, 1u << 24));
    if (which <= 1 || which == 6) smh::k_colour_table_fast<<<(1u << 24) / 256, 256, 0, h->s_main>>>(which, h->d_lut, d);
    else smh::k_colour_table<<<(1u << 24) / 256, 256, 0, h->s_main>>>(which >= 4 ? which - 4 : which, d);
    h->launches++;
    CU(h, cudaGetLastError());
    CU(h, cudaMemcpyAsync(host_out, d, 1u << 24, cudaMemcpyDeviceToHost, h->s_main));
    CU(h, cudaStreamSynchronize(h->s_main));
    cudaFree(d);
    return SMH_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// several GPUs of one box: contiguous frame ranges, one host thread + handle per GPU, no collective
// ------------------------------------------------------------------------------------------
struct smh_b200_multi {
    std::vector<smh_b200*> h;
    std::vector<double> seconds;
    std::mutex err_mu;
    std::string err;
};

extern "C" {

void smh_b200_shard_range(uint32_t n, uint32_t n_dev, uint32_t g, uint32_t* first, uint32_t* count) {
    if (n_dev == 0 || g >= n_dev) { if (first) *first = 0; if (count) *count = 0; return; }
    const uint32_t a = (uint32_t)((uint64_t)g * n / n_dev), b = (uint32_t)((uint64_t)(g + 1) * n / n_dev);
    if (first) *first = a;
    if (count) *count = b - a;
}

int32_t smh_b200_multi_create(const int32_t* devices, uint32_t n_dev, smh_b200_multi** out) {
    if (!devices || !out || n_dev == 0 || n_dev > 64) return SMH_E_ARG;
    *out = nullptr;
    for (uint32_t i = 0; i < n_dev; i++)
        for (uint32_t j = 0; j < i; j++)
            if (devices[i] == devices[j]) return fail(nullptr, SMH_E_ARG, "multi_create: device %d listed twice", devices[i]);
    smh_b200_multi* m = new smh_b200_multi();
    for (uint32_t i = 0; i < n_dev; i++) {
        smh_b200* hd = nullptr;
        const int rc = smh_b200_create(devices[i], &hd);
        if (rc) { smh_b200_multi_destroy(m); return rc; }  // (the text is in smh_b200_last_error(NULL))
        m->h.push_back(hd);
    }
    m->seconds.assign(n_dev, 0.0);
    *out = m;
    return SMH_OK;
}

void smh_b200_multi_destroy(smh_b200_multi* m) {
    if (!m) return;
    for (smh_b200* hd : m->h) smh_b200_destroy(hd);
    delete m;
}

const char* smh_b200_multi_last_error(smh_b200_multi* m) {
    if (!m) return "";
    std::lock_guard<std::mutex> g(m->err_mu);
    thread_local std::string copy;
    copy = m->err;
    return copy.c_str();
}

uint32_t smh_b200_multi_device_count(smh_b200_multi* m) { return m ? (uint32_t)m->h.size() : 0u; }

int32_t smh_b200_multi_load_map_markers(smh_b200_multi* m, const uint8_t* rgba, uint32_t n_markers, uint32_t size) {
    if (!m) return SMH_E_ARG;
    for (smh_b200* hd : m->h) {
        const int rc = smh_b200_load_map_markers(hd, rgba, n_markers, size);
        if (rc) { std::lock_guard<std::mutex> g(m->err_mu); m->err = smh_b200_last_error(hd); return rc; }
    }
    return SMH_OK;
}

int32_t smh_b200_multi_last_times(smh_b200_multi* m, double* seconds, uint32_t n) {
    if (!m || !seconds || n < m->h.size()) return SMH_E_ARG;
    for (size_t i = 0; i < m->h.size(); i++) seconds[i] = m->seconds[i];
    return SMH_OK;
}

int32_t smh_b200_multi_process_host_batch(smh_b200_multi* m, const uint8_t* frames, uint32_t n, uint32_t W, uint32_t H,
                                          uint32_t pitch_bytes, uint64_t frame_stride, const uint32_t* start_y, uint32_t max_gap,
                                          uint32_t flags, uint32_t chunk, const smh_b200_host_results* res) {
    if (!m || !frames || !res || n == 0) return SMH_E_ARG;
    Geometry g;
    int rc = compute_geometry(W, H, &g);
    if (rc) { std::lock_guard<std::mutex> lk(m->err_mu); m->err = "multi_process_host_batch: a frame size the reference panics on"; return rc; }
    const size_t npx = (size_t)g.map[2] * g.map[3], nq = (size_t)g.brq[2] * g.brq[3];
    const uint32_t G = (uint32_t)m->h.size();
    std::vector<int> codes(G, 0);
    std::vector<std::thread> th;
    for (uint32_t d = 0; d < G; d++) {
        th.emplace_back([&, d] {
            uint32_t first = 0, cnt = 0;
            smh_b200_shard_range(n, G, d, &first, &cnt);
            const auto t0 = std::chrono::steady_clock::now();
            if (cnt) {
                smh_b200_host_results r = *res;  // the shard's slice of every result array
                if (r.ui_map) r.ui_map += (size_t)first * npx * 4;
                if (r.ocr_out) r.ocr_out += (size_t)first * nq;
                if (r.scales) r.scales += (size_t)first * nq;
                if (r.lsd_image) r.lsd_image += (size_t)first * npx;
                r.lines += (size_t)first * 128; r.n_lines += first; r.red_pixels += first; r.map_open += first;
                if (r.minimap) r.minimap += (size_t)first * 4;
                if (r.minimap_found) r.minimap_found += first;
                if (r.icon_found) r.icon_found += first;
                if (r.icon_xy) r.icon_xy += (size_t)first * 2;
                if (r.icon_sad) r.icon_sad += first;
                codes[d] = smh_b200_process_host_batch(m->h[d], frames + (uint64_t)first * frame_stride, cnt, W, H, pitch_bytes, frame_stride,
                                                       start_y ? start_y + first : nullptr, max_gap, flags, chunk, &r);
            }
            m->seconds[d] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        });
    }
    for (auto& t2 : th) t2.join();
    for (uint32_t d = 0; d < G; d++)
        if (codes[d]) {
            std::lock_guard<std::mutex> lk(m->err_mu);
            m->err = std::string("device shard ") + std::to_string(d) + ": " + smh_b200_last_error(m->h[d]);
            return codes[d];
        }
    return SMH_OK;
}

}  // extern "C"
