// kernels_icons.cuh — marker icon template match.
//
// The reference carries this stage commented out in both backends (vision-cpu/src/lib.rs:270-277,
// 282-355; vision-gpu/src/lib.rs:373-430 + cuda.cu:548-614) and in the orchestrator
// (src/vision/mod.rs:95-107).  Semantics restated here (the CPU backend's flavour):
//
//   candidates   marker-coloured pixels (x, y) of the map crop with size <= x < w-size-1 and
//                size <= y < h-size-1 (lib.rs:269-277)
//   score        for every candidate and every template t:  sum over template pixels with alpha != 0 of
//                u32( f32(|r-tr| + |g-tg| + |b-tb|) * (f32(alpha) / 255) )  against the ISOLATED crop
//                (non-marker pixels are black), the window's top-left corner at the candidate (lib.rs:307-331)
//   pick         the global minimum; D17: ties -> first candidate in raster order
//   edit         black out [x, min(x+size, w-1)) x [y, min(y+size, h-1)), then a 4x4 pure-green square at
//                (x + size/2 - 2, y + round(0.70*size) - 2) (lib.rs:333-352) — so mask_marker_lines sees
//                the icon gone and its tip kept.
//
// B200 design: the isolated crop is never materialised (the fused pixel pass writes only the dilated
// mask), so k_icons rebuilds it tile by tile in shared memory from the resident capture and the 2^24-bit
// predicate table, and only for tiles whose dilated-mask bitmap has a set bit (~90 % of the tiles of a
// map have none: one 128-byte read each).  Candidates of a tile are compacted with ballots; one warp
// scores one candidate, lanes grouped by template and split over the template's rows; a window pixel that
// is black (the common case) costs nothing — its contribution is precomputed per template and added up
// front, and the pixels that do count are found with one AND of the row's non-black mask and the
// template's opaque mask.  k_icon_apply then edits the dilated mask and its bitmap in place.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/smh_consts.h"
#include "kernels.cuh"

namespace smh {

constexpr int ICON_THREADS = 256;
constexpr int ICON_MAX_TEMPLATES = 8;
constexpr uint32_t ICON_TILE = 32;
constexpr unsigned long long ICON_NONE = 0xFFFFFFFFFFFFFFFFull;

// Per-handle template table (built on the host by smh_b200_load_map_markers), p = my*size + mx:
//   rgb/af/c0[p*n+t]  = packed b|g<<8|r<<16, alpha/255 (f32), contribution against a black pixel
//   base[t]           = sum_p c0[p*n+t]
//   opq[t*size+my]    = bit mx set where template t is opaque (alpha != 0) at (mx, my); size <= 32
struct IconTable {
    const uint32_t* rgb; const float* af; const uint32_t* c0; const uint32_t* base; const uint32_t* opq;
    uint32_t n, size;
};

struct IconArgs {
    const uint8_t* frames; uint64_t frame_stride; uint32_t pitch;  // BGRA captures
    uint32_t map_x, map_y, w, h;                                    // map crop, relative to `frames`
    const uint32_t* lut;                                            // marker predicate bit table
    const uint32_t* bits; uint64_t bits_stride; BitsGeom g;         // dilated mask, tile-major bitmap
    const uint8_t* open;                                            // nullptr: every frame
    IconTable t;
    uint32_t n_frames;
    unsigned long long* best;                                       // [n] sad << 32 | y << 16 | x, ICON_NONE-initialised
    uint32_t* counter;                                              // item dispenser, zeroed before the launch
};

// a BGRA word of the capture as the isolated crop holds it: b|g<<8|r<<16, or 0 when the colour is not a
// marker colour (vision-cpu/src/lib.rs:266-268)
__device__ __forceinline__ uint32_t icon_isolate(uint32_t word, const uint32_t* __restrict__ lut) {
    const uint32_t b = word & 0xFFu, g = (word >> 8) & 0xFFu, r = (word >> 16) & 0xFFu;
    return marker_bit(word, max(r, max(g, b)), min(r, min(g, b)), lut) ? (word & 0x00FFFFFFu) : 0u;
}

// Persistent CTAs over (frame, tile-row) items handed out by an atomic counter.  A tile is one 32 x 32
// tile of the dilated-mask bitmap (padded columns X = x + pad), so "any white pixel?" is one coalesced
// 128-byte read per warp.  Dynamic shared memory: window [(32+size-1)^2] words and one 64-bit row mask of
// its non-black pixels per row, candidate list [1024], queue of the row's live tiles, template table.
__global__ void __launch_bounds__(ICON_THREADS, 5) k_icons(const IconArgs a) {
    extern __shared__ __align__(16) uint32_t sm[];
    constexpr uint32_t NW = ICON_THREADS / 32;
    const uint32_t size = a.t.size, n = a.t.n, np = size * size;
    const uint32_t WIN = ICON_TILE + size - 1u, WP = WIN | 1u;  // <= 63 columns; odd pitch: rows fall in different banks
    unsigned long long* rowbits = reinterpret_cast<unsigned long long*>(sm);  // [64] bit c of row r: win[r][c] != 0
    uint32_t* win = sm + 128;                            // [WIN][WP]
    uint32_t* cand = win + WIN * WP;                     // [1024] cx | cy << 8 (tile coordinates)
    uint32_t* s_rgb = cand + ICON_TILE * ICON_TILE;      // [np * n]
    float* s_af = reinterpret_cast<float*>(s_rgb + np * n);
    uint32_t* s_c0 = reinterpret_cast<uint32_t*>(s_af + np * n);
    uint32_t* s_opq = s_c0 + np * n;                     // [n][size]
    uint32_t* queue = s_opq + n * size;                  // [g.tx] live tiles of the current tile-row
    __shared__ uint32_t s_item, s_nq, s_ncand;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    for (uint32_t j = tid; j < np * n; j += ICON_THREADS) { s_rgb[j] = a.t.rgb[j]; s_af[j] = a.t.af[j]; s_c0[j] = a.t.c0[j]; }
    // scoring: G lanes per template, each lane takes every G-th row of its template
    const uint32_t G = 32u / n, my_t = lane / G, sub = lane - my_t * G;
    const bool scorer = my_t < n;
    const uint32_t my_base = scorer ? a.t.base[my_t] : 0u;
    for (uint32_t j = tid; j < n * size; j += ICON_THREADS) s_opq[j] = a.t.opq[j];
    const uint32_t* my_opq = s_opq + (scorer ? my_t : 0u) * size;
    const uint32_t my_rows = scorer ? size : 0u;

    // candidate bounds (lib.rs:270-273), [lo, hi)
    const int x_lo = (int)size, y_lo = (int)size, x_hi = (int)(a.w - size - 1u), y_hi = (int)(a.h - size - 1u);
    const uint32_t items = a.g.ty * a.n_frames;

    uint32_t cur_frame = 0xFFFFFFFFu;
    unsigned long long my_best = ICON_NONE;  // every lane of a warp tracks the warp's best of the current frame
    for (;;) {
        __syncthreads();  // previous item's shared state is dead (also orders the table fill on the first pass)
        if (tid == 0) { s_item = atomicAdd(a.counter, 1u); s_nq = 0u; }
        __syncthreads();
        const uint32_t it = s_item;
        if (it >= items) break;
        const uint32_t f = it / a.g.ty, tyi = it - f * a.g.ty;
        if (f != cur_frame) {
            if (cur_frame != 0xFFFFFFFFu && lane == 0 && my_best != ICON_NONE) atomicMin(a.best + cur_frame, my_best);
            my_best = ICON_NONE;
            cur_frame = f;
        }
        const int ty0 = (int)(tyi * ICON_TILE);
        if ((a.open && !a.open[f]) || ty0 >= y_hi || ty0 + (int)ICON_TILE <= y_lo) continue;
        // live tiles of this tile-row: a tile without a white pixel in the dilated mask has no candidate
        const uint32_t* fb = a.bits + (uint64_t)f * a.bits_stride + (((uint64_t)tyi * a.g.tx) << 5);
        for (uint32_t t0 = 0; t0 < a.g.tx; t0 += NW) {
            const uint32_t ti = t0 + warp;
            uint32_t any = 0u;
            if (ti < a.g.tx) any = __ldg(fb + (ti << 5) + lane);
            any = __reduce_or_sync(0xFFFFFFFFu, any);
            const int tx0 = (int)(ti * ICON_TILE) - (int)a.g.pad;
            if (any && lane == 0 && tx0 < x_hi && tx0 + (int)ICON_TILE > x_lo) queue[atomicAdd(&s_nq, 1u)] = ti;
        }
        __syncthreads();
        const uint32_t nq = s_nq;
        const uint8_t* frame = a.frames + (uint64_t)f * a.frame_stride;
        for (uint32_t q = 0; q < nq; q++) {
            const int tx0 = (int)(queue[q] * ICON_TILE) - (int)a.g.pad;  // crop x of the tile's first column (may be < 0)
            if (tid == 0) s_ncand = 0u;
            __syncthreads();
            // isolated crop of the tile plus the template's reach, its non-black row masks, the tile's candidates
            for (uint32_t r0 = 0; r0 < WIN; r0 += NW) {
                const uint32_t r = r0 + warp;
                uint32_t nzw[2] = {0u, 0u};
#pragma unroll
                for (uint32_t half = 0; half < 2; half++) {  // whole warps iterate: the ballots are full-mask
                    const uint32_t c = half * 32u + lane;
                    bool is_cand = false, nz = false;
                    if (r < WIN && c < WIN) {
                        const int x = tx0 + (int)c, y = ty0 + (int)r;
                        uint32_t v = 0u;  // black is not a marker colour: pixels outside the crop stay 0
                        if (x >= 0 && (uint32_t)x < a.w && (uint32_t)y < a.h)
                            v = icon_isolate(__ldg(reinterpret_cast<const uint32_t*>(frame + (uint64_t)(a.map_y + (uint32_t)y) * a.pitch +
                                                                                     (uint64_t)(a.map_x + (uint32_t)x) * 4u)), a.lut);
                        win[r * WP + c] = v;
                        nz = v != 0u;
                        is_cand = nz && r < ICON_TILE && c < ICON_TILE && x >= x_lo && x < x_hi && y >= y_lo && y < y_hi;
                    }
                    nzw[half] = __ballot_sync(0xFFFFFFFFu, nz);
                    const uint32_t m = __ballot_sync(0xFFFFFFFFu, is_cand);
                    if (m) {
                        // raster order inside the tile is not needed: the key carries (y, x)
                        uint32_t pos = 0;
                        const uint32_t leader = __ffs(m) - 1u;
                        if (lane == leader) pos = atomicAdd(&s_ncand, (uint32_t)__popc(m));
                        pos = __shfl_sync(0xFFFFFFFFu, pos, leader) + __popc(m & ((1u << lane) - 1u));
                        if (is_cand) cand[pos] = c | (r << 8);
                    }
                }
                if (lane == 0 && r < WIN) rowbits[r] = ((unsigned long long)nzw[1] << 32) | nzw[0];
            }
            __syncthreads();
            const uint32_t nc = s_ncand;
            for (uint32_t k = warp; k < nc; k += NW) {
                const uint32_t cv = cand[k], cx = cv & 0xFFu, cy = cv >> 8;
                const uint32_t* wp = win + cy * WP + cx;
                // a black crop pixel scores a per-template constant (summed in `base`); only the template's
                // opaque pixels that see a non-black crop pixel are evaluated:
                // ((ad as f32) * (alpha as f32 / 255.0)) as u32 — lib.rs:323 — minus what a black pixel gives
                uint32_t acc = 0u;
                for (uint32_t my = sub; my < my_rows; my += G) {
                    uint32_t bits = (uint32_t)(rowbits[cy + my] >> cx) & my_opq[my];
                    const uint32_t* wr = wp + my * WP;
                    const uint32_t e0 = my * size * n + my_t;
                    while (bits) {
                        const uint32_t mx = __ffs(bits) - 1u;
                        bits &= bits - 1u;
                        const uint32_t idx = e0 + mx * n;
                        acc += (uint32_t)__fmul_rn((float)__vsadu4(wr[mx], s_rgb[idx]), s_af[idx]) - s_c0[idx];
                    }
                }
                uint32_t tot = 0u;  // sum over the template's lanes (a per-group REDUX with six different masks measured slower)
                for (uint32_t d = 0; d < G; d++) tot += __shfl_sync(0xFFFFFFFFu, acc, scorer ? my_t * G + d : lane);
                const uint32_t best_sad = __reduce_min_sync(0xFFFFFFFFu, scorer ? tot + my_base : 0xFFFFFFFFu);
                const unsigned long long key = ((unsigned long long)best_sad << 32) | ((unsigned long long)(uint32_t)(ty0 + (int)cy) << 16) |
                                               (uint32_t)(tx0 + (int)cx);
                my_best = key < my_best ? key : my_best;
            }
            __syncthreads();  // window, row masks and candidate list are rewritten by the next tile
        }
    }
    if (cur_frame != 0xFFFFFFFFu && lane == 0 && my_best != ICON_NONE) atomicMin(a.best + cur_frame, my_best);
}

struct IconApplyArgs {
    const uint8_t* frames; uint64_t frame_stride; uint32_t pitch;
    uint32_t map_x, map_y, w, h;
    const uint32_t* lut;
    uint8_t* lsd; uint64_t lsd_stride; uint32_t lsd_pitch, pad;   // dilated mask, L8 (pixel x at byte pad + x)
    uint32_t* bits; uint64_t bits_stride; BitsGeom g;              // and its bitmap (may be nullptr)
    uint32_t size;
    const unsigned long long* best;   // [n]
    uint8_t* found; uint32_t* xy; uint32_t* sad;   // [n], [n][2], [n]
};

// One CTA per frame: decode the winner, then rewrite the dilated mask where the edit can change it.
__global__ void __launch_bounds__(ICON_THREADS) k_icon_apply(const IconApplyArgs a) {
    const uint32_t f = blockIdx.x;
    const unsigned long long key = a.best[f];
    if (key == ICON_NONE) {
        if (threadIdx.x == 0) { a.found[f] = 0; a.xy[2 * f] = a.xy[2 * f + 1] = 0u; a.sad[f] = 0u; }
        return;
    }
    const uint32_t x0 = (uint32_t)(key & 0xFFFFu), y0 = (uint32_t)((key >> 16) & 0xFFFFu);
    if (threadIdx.x == 0) { a.found[f] = 1; a.xy[2 * f] = x0; a.xy[2 * f + 1] = y0; a.sad[f] = (uint32_t)(key >> 32); }
    const uint32_t size = a.size;
    // erased rectangle [x0, ex) x [y0, ey) and the green square [sx0, sx1) x [sy0, sy1)
    const uint32_t ex = min(x0 + size, a.w - 1u), ey = min(y0 + size, a.h - 1u);
    const uint32_t px = x0 + size / 2u, py = y0 + (uint32_t)roundf(__fmul_rn((float)size, SMH_MAP_MARKER_POI_LOCATION));
    const uint32_t sx0 = px >= 2u ? px - 2u : 0u, sy0 = py >= 2u ? py - 2u : 0u, sx1 = min(px + 2u, a.w), sy1 = min(py + 2u, a.h);
    // is pure green a marker colour?  (it is — Alpha's hue; asked of the table rather than assumed)
    const uint32_t green = 0x0000FF00u;
    const bool green_on = (__ldg(a.lut + (green >> 5)) >> (green & 31u)) & 1u;
    // every mask pixel within one step of an edited pixel
    const uint32_t bx0 = min(x0, sx0), by0 = min(y0, sy0), bx1 = max(ex, sx1), by1 = max(ey, sy1);
    const uint32_t rx0 = bx0 ? bx0 - 1u : 0u, ry0 = by0 ? by0 - 1u : 0u, rx1 = min(bx1 + 1u, a.w), ry1 = min(by1 + 1u, a.h);
    const uint32_t rw = rx1 - rx0, rh = ry1 - ry0;
    const uint8_t* frame = a.frames + (uint64_t)f * a.frame_stride;
    auto raw = [&](int x, int y) -> bool {  // predicate on the edited, isolated crop
        if (x < 0 || y < 0 || (uint32_t)x >= a.w || (uint32_t)y >= a.h) return false;
        const uint32_t ux = (uint32_t)x, uy = (uint32_t)y;
        if (ux >= sx0 && ux < sx1 && uy >= sy0 && uy < sy1) return green_on;
        if (ux >= x0 && ux < ex && uy >= y0 && uy < ey) return false;
        const uint32_t word = __ldg(reinterpret_cast<const uint32_t*>(frame + (uint64_t)(a.map_y + uy) * a.pitch + (uint64_t)(a.map_x + ux) * 4u));
        const uint32_t b = word & 0xFFu, g = (word >> 8) & 0xFFu, r = (word >> 16) & 0xFFu;
        return marker_bit(word, max(r, max(g, b)), min(r, min(g, b)), a.lut) != 0u;
    };
    for (uint32_t i = threadIdx.x; i < rw * rh; i += ICON_THREADS) {
        const int x = (int)(rx0 + i % rw), y = (int)(ry0 + i / rw);
        // 3x3 cross dilation — vision-cpu/src/lib.rs:372 (imageproc L1, k = 1)
        const bool on = raw(x, y) || raw(x - 1, y) || raw(x + 1, y) || raw(x, y - 1) || raw(x, y + 1);
        a.lsd[(uint64_t)f * a.lsd_stride + (uint64_t)y * a.lsd_pitch + a.pad + (uint32_t)x] = on ? 255 : 0;
        if (a.bits) {
            const uint32_t X = (uint32_t)x + a.g.pad;
            uint32_t* wd = a.bits + (uint64_t)f * a.bits_stride + (((((uint32_t)y >> 5) * a.g.tx + (X >> 5)) << 5) + ((uint32_t)y & 31u));
            if (on) atomicOr(wd, 1u << (X & 31u)); else atomicAnd(wd, ~(1u << (X & 31u)));
        }
    }
}

// debug view LSDPreprocess after the edit: black rectangle, green square (RGBA words, w pixels per row)
__global__ void k_icon_overlay(uint32_t* rgba, uint32_t w, uint32_t h, uint32_t x0, uint32_t y0, uint32_t size) {
    const uint32_t ex = min(x0 + size, w - 1u), ey = min(y0 + size, h - 1u);
    const uint32_t px = x0 + size / 2u, py = y0 + (uint32_t)roundf(__fmul_rn((float)size, SMH_MAP_MARKER_POI_LOCATION));
    const uint32_t sx0 = px >= 2u ? px - 2u : 0u, sy0 = py >= 2u ? py - 2u : 0u, sx1 = min(px + 2u, w), sy1 = min(py + 2u, h);
    const uint32_t bx1 = max(ex, sx1), by1 = max(ey, sy1), bx0 = min(x0, sx0), by0 = min(y0, sy0);
    const uint32_t rw = bx1 - bx0, rh = by1 - by0;
    for (uint32_t i = threadIdx.x; i < rw * rh; i += blockDim.x) {
        const uint32_t x = bx0 + i % rw, y = by0 + i / rw;
        if (x >= sx0 && x < sx1 && y >= sy0 && y < sy1) rgba[(uint64_t)y * w + x] = 0xFF00FF00u;   // (0,255,0,255)
        else if (x >= x0 && x < ex && y >= y0 && y < ey) rgba[(uint64_t)y * w + x] = 0xFF000000u;  // (0,0,0,255)
    }
}

__global__ void k_fill_u64(unsigned long long* p, uint32_t n, unsigned long long v) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace smh
