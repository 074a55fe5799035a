// kernels_next.cuh — the callers either side of the pixel pass / line search:
//
//   k_minimap    find_minimap (src/vision/find_minimap.rs:8-146): four edge searches from the centre of
//                the map ROI over the resident BGRA capture — the only consumer of the host frame in the
//                reference (get_cpu_frame, src/vision/mod.rs:85).
//   k_mpx_ratio  calc_meters_to_px_ratio (src/vision/mpx_ratio.rs:3-133): scale-bar widths read straight
//                from the device-resident `scales` plane, so that image need not cross the bus.
//
// Both are integer scans whose serial order decides the result; the kernels keep that order and spread
// the independent tests (positions along the search axis, pixels of a run, columns of a bar) over lanes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/smh_consts.h"

namespace smh {

// ------------------------------------------------------------------------------------------
// find_minimap
// ------------------------------------------------------------------------------------------
struct MinimapArgs {
    const uint8_t* frames;  // BGRA captures (or the uploaded region of them)
    uint64_t frame_stride;
    uint32_t pitch;         // bytes per row
    uint32_t map_x, map_y, w, h;  // the view, relative to `frames`
    uint32_t n_frames;
    const uint8_t* open;    // nullptr: every frame
    uint32_t* rect;         // [n][4] = left, right, top, bottom
    uint8_t* found;         // [n]: 1 = Some(Rect)
};

// get_edginess(x, y) <= EDGINESS_THRESHOLD — find_minimap.rs:8-46,66.  The f32 test
// `max as f32 / 765.0 <= 0.01` is `max <= 7` (SMH_MINIMAP_EDGINESS_MAX_SUM); the sum of absolute B,G,R
// differences of two BGRA words is one SAD over the words with alpha masked off.
__device__ __forceinline__ bool mm_low_edginess(const uint8_t* view, uint32_t pitch, uint32_t x, uint32_t y) {
    const uint8_t* p = view + (uint64_t)y * pitch + (uint64_t)x * 4u;
    const uint32_t c = __ldg(reinterpret_cast<const uint32_t*>(p)) & 0x00FFFFFFu;
    uint32_t mx = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
            if (dx == 0 && dy == 0) continue;
            const uint32_t q = __ldg(reinterpret_cast<const uint32_t*>(p + (int64_t)dy * (int64_t)pitch + dx * 4)) & 0x00FFFFFFu;
            mx = max(mx, __vsadu4(c, q));
        }
    return mx <= SMH_MINIMAP_EDGINESS_MAX_SUM;
}

constexpr int MM_THREADS = 128;

// grid = (4 directions, n_frames).  find_edge — find_minimap.rs:65-129: walk from the centre along the
// search axis; the first position whose pixel AND a run of min_line_length pixels along the other axis
// are all below the edginess threshold gives the edge.  The block tests MM_THREADS positions of the walk
// at once, then takes the hits in walk order; a hit's run is tested MM_THREADS pixels at a time.
__global__ void __launch_bounds__(MM_THREADS) k_minimap(const MinimapArgs a) {
    const uint32_t f = blockIdx.y;
    const int dir = (int)blockIdx.x;  // 0 Left, 1 Right, 2 Up, 3 Down
    const int t = (int)threadIdx.x;
    if (a.open && !a.open[f]) {
        if (t == 0) { a.rect[4u * f + dir] = 0u; if (dir == 0) a.found[f] = 0; }
        return;
    }
    if (a.w < 3u || a.h < 3u) {  // `return None`
        if (t == 0) { a.rect[4u * f + dir] = 0u; if (dir == 0) a.found[f] = 0; }
        return;
    }
    const uint8_t* view = a.frames + (uint64_t)f * a.frame_stride + (uint64_t)a.map_y * a.pitch + (uint64_t)a.map_x * 4u;
    const int c_axis = dir < 2 ? 0 : 1;
    const int cod = (dir & 1) ? 1 : -1;
    const uint32_t start[2] = {a.w / 2u, a.h / 2u};
    const uint32_t c_max0 = c_axis == 0 ? a.w : a.h, oc_max0 = c_axis == 0 ? a.h : a.w;
    const uint32_t s_c = start[c_axis], s_oc = start[1 - c_axis];
    const uint32_t min_line_length = (oc_max0 - s_oc) / 2u - 1u;  // abs_diff: oc_max0 > s_oc
    const uint32_t c_max = c_max0 - 3u, oc_max = oc_max0 - 3u;

    __shared__ uint32_t hits[MM_THREADS / 32];
    __shared__ uint32_t ends[MM_THREADS / 32];
    __shared__ int run_failed;
    uint32_t result = 0;
    for (uint32_t base = 0;; base += MM_THREADS) {
        // position base + t + 1 of the walk: (xy[c] as i32 + cod) as u32, step by step
        const uint32_t pos = (uint32_t)((int32_t)s_c + cod * (int32_t)(base + (uint32_t)t + 1u));
        const bool terminal = pos > c_max || pos < 3u;  // `return Some(c_max + 2)` / `return Some(0)`
        bool hit = false;
        if (!terminal) hit = c_axis == 0 ? mm_low_edginess(view, a.pitch, pos, s_oc) : mm_low_edginess(view, a.pitch, s_oc, pos);
        const uint32_t hb = __ballot_sync(0xFFFFFFFFu, hit), eb = __ballot_sync(0xFFFFFFFFu, terminal);
        if ((t & 31) == 0) { hits[t >> 5] = hb; ends[t >> 5] = eb; }
        __syncthreads();
        bool done = false;
        for (int wi = 0; wi < MM_THREADS / 32 && !done; wi++) {
            uint32_t todo = hits[wi] | ends[wi];
            while (todo) {  // block-uniform
                const int b = __ffs(todo) - 1;
                todo &= todo - 1u;
                const uint32_t p = (uint32_t)((int32_t)s_c + cod * (int32_t)(base + (uint32_t)(wi * 32 + b) + 1u));
                if ((ends[wi] >> b) & 1u) { result = p > c_max ? c_max + 2u : 0u; done = true; break; }
                // the run: pixels s_oc - cod*k, k = 1..min_line_length, all inside [3, oc_max] and all low
                bool ok = true;
                for (uint32_t k0 = 0; k0 < min_line_length; k0 += MM_THREADS) {
                    if (t == 0) run_failed = 0;
                    __syncthreads();
                    const uint32_t k = k0 + (uint32_t)t + 1u;
                    if (k <= min_line_length) {
                        const uint32_t o = (uint32_t)((int32_t)s_oc - cod * (int32_t)k);
                        bool good = !(o < 3u || o > oc_max);
                        if (good) good = c_axis == 0 ? mm_low_edginess(view, a.pitch, p, o) : mm_low_edginess(view, a.pitch, o, p);
                        if (!good) run_failed = 1;
                    }
                    __syncthreads();
                    const bool failed = run_failed != 0;
                    __syncthreads();
                    if (failed) { ok = false; break; }
                }
                if (ok) { result = (uint32_t)((int32_t)p - cod); done = true; break; }
            }
        }
        __syncthreads();
        if (done) break;
    }
    if (t == 0) {
        a.rect[4u * f + dir] = result;
        if (dir == 0) a.found[f] = 1;
    }
}

// ------------------------------------------------------------------------------------------
// calc_meters_to_px_ratio
// ------------------------------------------------------------------------------------------
struct MpxArgs {
    const uint8_t* scales;   // L8 plane(s): pixel (x, y) of frame f at scales + f*stride + y*pitch + pad + x
    uint64_t stride;
    uint32_t pitch, pad, w, h;
    uint32_t n_frames;
    const uint32_t* in;      // [n][3][3] = meters, x, y
    const uint8_t* n_in;     // [n] scales per frame (0..3)
    double* ratio;           // [n]
    uint8_t* found;          // [n] 1 = Some(ratio)
    uint32_t* debug;         // [n][3][4] = meters, left, right, y of each bar found (compacted), or nullptr
    uint8_t* n_debug;        // [n]
};

// D16: a pixel outside the image is non-black (the reference reads out of bounds there)
__device__ __forceinline__ bool mpx_black(const uint8_t* img, uint32_t pitch, uint32_t w, uint32_t h, uint32_t x, uint32_t y) {
    return x < w && y < h && __ldg(img + (uint64_t)y * pitch + x) == 0;
}
// the four pixels (x, y) .. (x, y + 3) — mpx_ratio.rs:24-29 (the second half of the chain is an empty range)
__device__ __forceinline__ bool mpx_vbar(const uint8_t* img, uint32_t pitch, uint32_t w, uint32_t h, uint32_t x, uint32_t y) {
    bool ok = true;
#pragma unroll
    for (uint32_t k = 0; k < SMH_SCALE_MIN_VBAR_HEIGHT; k++) ok = ok && mpx_black(img, pitch, w, h, x, y + k);
    return ok;
}

// grid = n_frames, block = 96: warp k runs find_scale_width for scale k (mpx_ratio.rs:4-77) with the
// columns of the right / left walks spread over its lanes; thread 0 then averages the scales that were
// found, in order, in f64 (mpx_ratio.rs:98-124).
__global__ void __launch_bounds__(96) k_mpx_ratio(const MpxArgs a) {
    const uint32_t f = blockIdx.x;
    const int k = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31u);
    __shared__ double s_ratio[3];
    __shared__ uint32_t s_line[3][3];
    __shared__ int s_found[3];
    const uint32_t n_in = a.n_in[f] > 3 ? 3u : a.n_in[f];
    const uint8_t* img = a.scales + (uint64_t)f * a.stride + a.pad;
    const uint32_t w = a.w, h = a.h;
    if (lane == 0) s_found[k] = 0;
    if ((uint32_t)k < n_in) {
        const uint32_t meters = a.in[(f * 3u + k) * 3u], x = a.in[(f * 3u + k) * 3u + 1u], y0 = a.in[(f * 3u + k) * 3u + 2u];
        if (y0 >= SMH_SCALE_MIN_VBAR_HEIGHT && x < w) {
            // max_scale_y_offset = ((20.0 / 640.0) * width as f64).round() as u32
            const uint32_t max_off = (uint32_t)__double2ll_rn(floor(__dmul_rn(20.0 / 640.0, (double)w) + 0.5));
            const unsigned long long y_end64 = (unsigned long long)y0 + max_off;
            const uint32_t y_end = y_end64 < h ? (uint32_t)y_end64 : h;
            for (uint32_t y = y0; y < y_end; y++) {  // warp-uniform
                if (!mpx_black(img, a.pitch, w, h, x, y)) continue;
                // go right: first column >= x that carries a vertical bar
                uint32_t right = 0;
                for (uint32_t x0 = x; x0 < w; x0 += 32u) {
                    const uint32_t xr = x0 + (uint32_t)lane;
                    const uint32_t m = __ballot_sync(0xFFFFFFFFu, xr < w && mpx_vbar(img, a.pitch, w, h, xr, y));
                    if (m) { right = x0 + (uint32_t)(__ffs(m) - 1); break; }
                }
                if (right == 0) continue;
                right -= 1u;
                // go left: first column < x, walking down, that carries a vertical bar
                uint32_t left = 0;
                for (uint32_t done = 0; done < x; done += 32u) {
                    const uint32_t off = done + (uint32_t)lane + 1u;  // column x - off
                    const uint32_t m = __ballot_sync(0xFFFFFFFFu, off <= x && mpx_vbar(img, a.pitch, w, h, x - off, y));
                    if (m) { left = x - (done + (uint32_t)__ffs(m)); break; }
                }
                if (left == 0) continue;
                left += 1u;
                const uint32_t width = right - left;  // wraps like the release build
                if (width < SMH_SCALE_MIN_WIDTH) continue;
                if (lane == 0) {
                    s_found[k] = 1;
                    s_line[k][0] = left; s_line[k][1] = right; s_line[k][2] = y;
                    s_ratio[k] = __ddiv_rn((double)meters, (double)width);
                }
                break;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sum = 0.0;
        uint32_t cnt = 0;
        for (uint32_t j = 0; j < n_in; j++)
            if (s_found[j]) {
                sum = cnt ? __dadd_rn(sum, s_ratio[j]) : s_ratio[j];
                if (a.debug) {
                    uint32_t* d = a.debug + ((uint64_t)f * 3u + cnt) * 4u;
                    d[0] = a.in[(f * 3u + j) * 3u]; d[1] = s_line[j][0]; d[2] = s_line[j][1]; d[3] = s_line[j][2];
                }
                cnt++;
            }
        a.found[f] = cnt ? 1 : 0;
        a.ratio[f] = cnt == 0 ? 0.0 : (cnt == 1 ? sum : __ddiv_rn(sum, (double)cnt));
        if (a.n_debug) a.n_debug[f] = (uint8_t)cnt;
    }
}

}  // namespace smh
