// kernels.cuh — sm_100a kernels of the map-analysis vision path.
//
//   k_pixels   fused pixel pass: one read of the captured BGRA map ROI produces ui_map,
//              the dilated marker mask (lsd_image), ocr_out and scales, plus the
//              close-deployment-button test.  Replaces reference kernels K1-K9
//              (vision-gpu/cuda/cuda.cu:312-546,616-635 + NPP dilate) and restates the
//              vision-cpu arithmetic (vision-cpu/src/lib.rs:110-280,357-375) bit for bit.
//   k_lsd      whole lsd::find_lines seed scan + 3600-ray search per frame, one CTA per
//              frame, no host round trips.  Replaces K10 + the host loop
//              (vision-common/src/lsd.rs:60-105, vision-cpu/src/lib.rs:387-449).
//   k_longest  single find_longest_line call (trait method).
//   k_synth    synthetic capture rasteriser; k_colour_table exhaustive predicate tables.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "smh_pixel.h"
#include "smh_synth.h"

namespace smh {

// ------------------------------------------------------------------------------------------
// small PTX helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_v4(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void stg_u32(void* p, uint32_t v) {
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// ------------------------------------------------------------------------------------------
// fused pixel pass
// ------------------------------------------------------------------------------------------
struct PixelArgs {
    const uint8_t* frames;  // captured BGRA frames; ROI coordinates below are relative to this origin
    uint64_t frame_stride;
    uint32_t pitch;
    uint32_t map_x, map_y, w, h, bqw, bqh;
    uint32_t btn_x, btn_y, btn_w, btn_h;
    uint32_t gx0;   // map_x & ~3: frame column of vector 0
    uint32_t pad;   // map_x - gx0 (0..3)
    uint32_t nvec;  // 16-byte vectors per map row = ceil((pad + w) / 4)
    uint32_t qpad;  // (pad + bqw) & 3
    uint8_t* ui; uint64_t ui_stride; uint32_t ui_pitch;
    uint8_t* lsd; uint64_t lsd_stride; uint32_t lsd_pitch;
    uint8_t* ocr; uint8_t* scales; uint64_t q_stride; uint32_t q_pitch;
    const uint32_t* start_y;  // device [n] or nullptr
    uint32_t start_y_uniform;
    uint32_t* red; uint8_t* open;
    uint32_t rows_per_band, nbands;
    uint32_t grayscale;
    uint32_t ring_words;  // words per ring row = nvec + 2 (one zero guard word each side)
    const uint32_t* lut;  // marker predicate, 1 bit per colour (r<<16|g<<8|b)
};

constexpr int PX_THREADS = 256;
constexpr int PX_TX = 64;              // threads along a row
constexpr int PX_G = PX_THREADS / PX_TX;  // rows per group (4)
constexpr int PX_RING = 16;            // ring rows (power of two >= G + 6 + G)
constexpr int PX_UNR = 4;              // 16-byte loads in flight per thread

// flag bits per pixel byte in the ring
constexpr uint32_t FL_M = 1u;      // marker colour
constexpr uint32_t FL_CORE = 2u;   // bright monochrome pixel that an OCR neighbour search may find
constexpr uint32_t FL_EDGE = 4u;   // OCR edge candidate (keep only with a core pixel in its window)

// byte `k` of a little-endian word as an exact float: PRMT builds 0x4B0000xx (= 2^23 + xx), one
// FADD removes the bias.  Full-rate replacement for I2F.
template <int K>
__device__ __forceinline__ float byte_as_f32(uint32_t word) {
    return __fsub_rn(__uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540 + K)), 8388608.0f);
}
// luma8 of a BGRA word (image 0.23.14 rgb_to_luma, unfused, truncating): identical bits to
// smh_luma8 (proved over all 2^24 colours on the device, tests/test_gpu_parity.py); the
// truncation is a round-down add of 2^23 whose low mantissa byte is floor(l) (l < 256).
__device__ __forceinline__ uint32_t luma8_bgra(uint32_t word) {
    const float l = __fadd_rn(__fadd_rn(__fmul_rn(SMH_LUMA_R, byte_as_f32<2>(word)), __fmul_rn(SMH_LUMA_G, byte_as_f32<1>(word))),
                              __fmul_rn(SMH_LUMA_B, byte_as_f32<0>(word)));
    return __float_as_uint(__fadd_rd(l, 8388608.0f)) & 0xFFu;
}

// Marker predicate of a BGRA word: integer fast-reject, then one bit of the 2^24-entry table
// that k_build_marker_lut fills with smh_is_marker_exact (2 MB, L2 resident).
__device__ __forceinline__ uint32_t marker_bit(uint32_t word, uint32_t mx, uint32_t mn, const uint32_t* __restrict__ lut) {
    if (mx >= 178u && 100u * (mx - mn) >= 34u * mx) {
        const uint32_t c = word & 0x00FFFFFFu;  // r<<16 | g<<8 | b
        return (__ldg(lut + (c >> 5)) >> (c & 31u)) & 1u;
    }
    return 0u;
}

// One pixel of phase 1.  MODE 0: interior vector outside the OCR quadrant; MODE 1: interior vector
// wholly inside the quadrant and below the neighbour-search clip (D10); MODE 2: generic (row ends,
// quadrant borders) with every test per pixel.
template <bool GRAY, int MODE>
__device__ __forceinline__ void classify_px(uint32_t word, int e, int i, int jq, bool owned, bool inq_row,
                                            const PixelArgs& a, uint32_t& fl, uint32_t& uo, uint32_t& ocrw, uint32_t& scw) {
    const uint32_t b = word & 0xFFu, g = (word >> 8) & 0xFFu, r = (word >> 16) & 0xFFu;
    const uint32_t mx = max(r, max(g, b)), mn = min(r, min(g, b));
    const bool valid = MODE != 2 || (i >= 0 && i < (int)a.w);
    uint32_t fe = valid ? marker_bit(word, mx, mn, a.lut) : 0u;
    const bool inq = MODE == 1 || (MODE == 2 && inq_row && valid && i >= (int)a.bqw && i < (int)(2u * a.bqw));
    uint32_t luma = 0;
    if ((GRAY && owned) || inq) luma = luma8_bgra(word);
    if (inq) {
        // OCR classes — vision-cpu/src/lib.rs:39-53,188-193 (mono = 4*(max-min))
        if (mx == mn && mn >= SMH_OCR_BRIGHT) {
            // visible to a neighbour's search only below the clip min(x+3, w-3) (D10)
            if (MODE == 1 || (i - (int)a.bqw <= (int)a.bqw - SMH_OCR_RADIUS && jq <= (int)a.bqh - SMH_OCR_RADIUS)) fe |= FL_CORE;
            ocrw ^= luma << (8 * e);  // 255 - luma
        } else if (4u * (mx - mn) <= SMH_OCR_MONO_EDGE && mn >= SMH_OCR_BRIGHT_EDGE) {
            fe |= FL_EDGE;
        }
        scw |= (luma != 0u ? 255u : 0u) << (8 * e);
    }
    fl |= fe << (8 * e);
    // ui_map: (L,L,L,255) or (r,g,b,255) — vision-cpu/src/lib.rs:150-163
    uo = GRAY ? (luma * 0x00010101u) | 0xFF000000u : __byte_perm(word, 0xFF000000u, 0x7012);
}

template <bool GRAY, int MODE>
__device__ __forceinline__ void classify_vec(const uint4& v, int q, int jq, bool owned, bool inq_row, const PixelArgs& a,
                                             uint32_t& fl, uint4& uo, uint32_t& ocrw, uint32_t& scw) {
    const int i0 = 4 * q - (int)a.pad;
    classify_px<GRAY, MODE>(v.x, 0, i0 + 0, jq, owned, inq_row, a, fl, uo.x, ocrw, scw);
    classify_px<GRAY, MODE>(v.y, 1, i0 + 1, jq, owned, inq_row, a, fl, uo.y, ocrw, scw);
    classify_px<GRAY, MODE>(v.z, 2, i0 + 2, jq, owned, inq_row, a, fl, uo.z, ocrw, scw);
    classify_px<GRAY, MODE>(v.w, 3, i0 + 3, jq, owned, inq_row, a, fl, uo.w, ocrw, scw);
}

template <bool GRAY>
__global__ void __launch_bounds__(PX_THREADS) k_pixels(const PixelArgs a) {
    extern __shared__ __align__(16) uint32_t ring[];  // [PX_RING][ring_words]
    const uint32_t per_frame = a.nbands + 1u;
    const uint32_t f = blockIdx.x / per_frame;
    const uint32_t band = blockIdx.x % per_frame;
    const uint8_t* frame = a.frames + (uint64_t)f * a.frame_stride;
    const int tid = threadIdx.x;

    if (band == a.nbands) {
        // ---- close-deployment button test (vision-cpu/src/lib.rs:115-135) -----------------
        uint32_t cnt = 0;
        const uint32_t npx = a.btn_w * a.btn_h;
        for (uint32_t idx = tid; idx < npx; idx += PX_THREADS) {
            const uint32_t y = idx / a.btn_w, x = idx - y * a.btn_w;
            const uint32_t word = __ldg(reinterpret_cast<const uint32_t*>(
                frame + (uint64_t)(a.btn_y + y) * a.pitch + (uint64_t)(a.btn_x + x) * 4u));
            cnt += smh_button_red((word >> 16) & 0xFFu, (word >> 8) & 0xFFu, word & 0xFFu) ? 1u : 0u;
        }
        cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
        __shared__ uint32_t s_cnt[PX_THREADS / 32];
        if ((tid & 31) == 0) s_cnt[tid >> 5] = cnt;
        __syncthreads();
        if (tid == 0) {
            uint32_t total = 0;
            for (int k = 0; k < PX_THREADS / 32; k++) total += s_cnt[k];
            a.red[f] = total;
            // `red_pixels as f32 / (w*h) as f32 < 0.65` -> map closed
            a.open[f] = __fdiv_rn((float)total, (float)npx) < SMH_BTN_RED_RATIO_THRESHOLD ? 0 : 1;
        }
        return;
    }

    const int tx = tid % PX_TX, ty = tid / PX_TX;
    const int r0 = (int)(band * a.rows_per_band);
    const int r1 = min(r0 + (int)a.rows_per_band, (int)a.h);
    const int bqh = (int)a.bqh;
    const bool band_has_brq = r1 > bqh && r0 < 2 * bqh;
    const int halo = band_has_brq ? SMH_OCR_RADIUS : 1;
    const uint32_t start_y = a.start_y ? a.start_y[f] : a.start_y_uniform;
    const uint32_t rw = a.ring_words;

    // zero the guard words of every ring row (never written afterwards)
    for (int k = tid; k < PX_RING; k += PX_THREADS) { ring[k * rw] = 0u; ring[k * rw + rw - 1] = 0u; }

    uint8_t* ui_f = a.ui + (uint64_t)f * a.ui_stride;
    uint8_t* lsd_f = a.lsd + (uint64_t)f * a.lsd_stride;
    uint8_t* ocr_f = a.ocr + (uint64_t)f * a.q_stride;
    uint8_t* sc_f = a.scales + (uint64_t)f * a.q_stride;
    // vector range that intersects the brq columns [bqw, 2*bqw)
    const int qv_lo = ((int)a.bqw + (int)a.pad) / 4;                  // first vector with i >= bqw
    const int qv_hi = ((int)(2u * a.bqw) - 1 + (int)a.pad) / 4;       // last vector with i < 2*bqw
    const int qword0 = ((int)a.qpad - (int)a.pad - (int)a.bqw);       // + 4q = byte position in the brq row

    for (int a0 = r0 - halo; a0 - SMH_OCR_RADIUS < r1; a0 += PX_G) {
        // ---------------- phase 1: load, classify, emit the stencil-free outputs -----------
        {
            const int j = a0 + ty;
            const bool needed = j < r1 + halo;
            if (needed) {
                uint32_t* rrow = ring + (uint32_t)(j & (PX_RING - 1)) * rw + 1;
                if (j < 0 || j >= (int)a.h) {
                    for (int q = tx; q < (int)a.nvec; q += PX_TX) rrow[q] = 0u;
                } else {
                    const bool owned = j >= r0 && j < r1;
                    const int jq = j - bqh;
                    const bool inq_row = jq >= 0 && jq < bqh;
                    const uint8_t* src = frame + (uint64_t)(a.map_y + (uint32_t)j) * a.pitch + (uint64_t)a.gx0 * 4u;
                    uint8_t* ui_row = ui_f + (uint64_t)j * a.ui_pitch;
                    uint8_t* ocr_row = ocr_f + (uint64_t)(inq_row ? jq : 0) * a.q_pitch;
                    uint8_t* sc_row = sc_f + (uint64_t)(inq_row ? jq : 0) * a.q_pitch;
                    const bool write_scales = inq_row && (uint32_t)jq >= start_y;
                    for (int q0 = tx; q0 < (int)a.nvec; q0 += PX_TX * PX_UNR) {
                        uint4 v[PX_UNR];
#pragma unroll
                        for (int u = 0; u < PX_UNR; u++) {
                            const int q = q0 + u * PX_TX;
                            if (q < (int)a.nvec) v[u] = ldg_nc_v4(src + (uint64_t)q * 16u);
                        }
#pragma unroll
                        for (int u = 0; u < PX_UNR; u++) {
                            const int q = q0 + u * PX_TX;
                            if (q >= (int)a.nvec) continue;
                            uint32_t fl = 0, ocrw = 0xFFFFFFFFu, scw = 0;
                            uint4 uo;
                            const int i_lo = 4 * q - (int)a.pad, i_hi = i_lo + 3;
                            const bool row_end = q == 0 || q == (int)a.nvec - 1;
                            const bool q_out = !inq_row || i_hi < (int)a.bqw || i_lo >= (int)(2u * a.bqw);
                            const bool q_in = inq_row && i_lo >= (int)a.bqw && i_hi <= (int)(2u * a.bqw) - SMH_OCR_RADIUS &&
                                              jq <= bqh - SMH_OCR_RADIUS;
                            if (!row_end && q_out) classify_vec<GRAY, 0>(v[u], q, jq, owned, inq_row, a, fl, uo, ocrw, scw);
                            else if (!row_end && q_in) classify_vec<GRAY, 1>(v[u], q, jq, owned, inq_row, a, fl, uo, ocrw, scw);
                            else classify_vec<GRAY, 2>(v[u], q, jq, owned, inq_row, a, fl, uo, ocrw, scw);
                            rrow[q] = fl;
                            if (owned) {
                                stg_v4(ui_row + (uint64_t)q * 16u, uo);
                                if (inq_row && q >= qv_lo && q <= qv_hi) {
                                    const int pos = qword0 + 4 * q;  // multiple of 4, >= 0
                                    stg_u32(ocr_row + pos, ocrw);
                                    if (write_scales) stg_u32(sc_row + pos, scw);
                                }
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
        // ---------------- phase 2: stencils (3x3 cross dilation, OCR neighbour search) -----
        {
            const int j = a0 - SMH_OCR_RADIUS + ty;
            if (j >= r0 && j < r1) {
                const uint32_t* rc = ring + (uint32_t)(j & (PX_RING - 1)) * rw + 1;
                const uint32_t* ru = ring + (uint32_t)((j - 1) & (PX_RING - 1)) * rw + 1;
                const uint32_t* rd = ring + (uint32_t)((j + 1) & (PX_RING - 1)) * rw + 1;
                const bool has_u = j > 0, has_d = j + 1 < (int)a.h;
                uint8_t* lsd_row = lsd_f + (uint64_t)j * a.lsd_pitch;
                const int jq = j - bqh;
                const bool inq_row = jq >= 0 && jq < bqh;
                for (int q = tx; q < (int)a.nvec; q += PX_TX) {
                    const uint32_t c = rc[q], l = rc[q - 1], r = rc[q + 1];
                    const uint32_t mc = c & 0x01010101u;
                    uint32_t d = mc | (mc << 8) | ((l & 0x01000000u) >> 24) | (mc >> 8) | ((r & 1u) << 24);
                    if (has_u) d |= ru[q] & 0x01010101u;
                    if (has_d) d |= rd[q] & 0x01010101u;
                    // bytes of pixels outside [0, w) stay 0
                    uint32_t vm = 0;
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int i = 4 * q + e - (int)a.pad;
                        vm |= (i >= 0 && i < (int)a.w) ? (1u << (8 * e)) : 0u;
                    }
                    stg_u32(lsd_row + (uint64_t)q * 4u, (d & vm) * 255u);

                    uint32_t eb = (c >> 2) & 0x01010101u;
                    if (inq_row && eb) {
#pragma unroll 1
                        for (int e = 0; e < 4; e++) {
                            if (!((eb >> (8 * e)) & 1u)) continue;
                            const int col = 4 * q + e;  // byte column in the ring row (pad included)
                            bool found = false;
#pragma unroll 1
                            for (int dy = -SMH_OCR_RADIUS; dy <= SMH_OCR_RADIUS && !found; dy++) {
                                const int jj = j + dy;
                                if (jj < bqh || jj >= 2 * bqh) continue;
                                const uint8_t* rb = reinterpret_cast<const uint8_t*>(
                                    ring + (uint32_t)(jj & (PX_RING - 1)) * rw + 1);
                                for (int dx = -SMH_OCR_RADIUS; dx <= SMH_OCR_RADIUS; dx++) {
                                    const int cc = col + dx;
                                    if (cc < 0 || cc >= 4 * (int)a.nvec) continue;
                                    if (rb[cc] & FL_CORE) { found = true; break; }
                                }
                            }
                            if (found) {
                                const int i = col - (int)a.pad;
                                const uint32_t word = __ldg(reinterpret_cast<const uint32_t*>(
                                    frame + (uint64_t)(a.map_y + (uint32_t)j) * a.pitch +
                                    (uint64_t)(a.map_x + (uint32_t)i) * 4u));
                                const uint32_t L = luma8_bgra(word);
                                ocr_f[(uint64_t)jq * a.q_pitch + (uint32_t)(i - (int)a.bqw) + a.qpad] = (uint8_t)(255u - L);
                            }
                        }
                    }
                }
            }
        }
        // no second barrier: the next group's phase 1 writes ring rows a0+G..a0+2G-1, disjoint
        // (mod PX_RING) from the rows a0-6..a0+G-1 this phase 2 reads; the barrier of the next
        // group orders everything after that.
    }
}

// ------------------------------------------------------------------------------------------
// line search
// ------------------------------------------------------------------------------------------
struct LsdArgs {
    const uint8_t* lsd; uint64_t stride; uint32_t pitch, pad, w, h;
    const float2* rays;  // 3600 (cos, sin) pairs built on the host with libm (D11)
    float max_gap;
    const uint8_t* open;  // nullptr: every frame open
    float* lines;         // [n][32][4]
    uint32_t* n_lines;    // [n]
    unsigned long long* stats;  // [n][2] or nullptr
};

struct Img {
    const uint8_t* p; uint32_t pitch, w, h;
    __device__ __forceinline__ uint32_t at(uint32_t x, uint32_t y) const { return __ldg(p + (uint64_t)y * pitch + x); }
};

// find_line_in_image — vision-cpu/src/lib.rs:388-432.  Returns the squared length, writes the end point.
__device__ __forceinline__ float march(const Img& im, float px, float py, float max_gap, float dx, float dy,
                                       float& ex, float& ey, uint32_t& steps) {
    float x = px, y = py, xo = 0.0f, yo = 0.0f;
    float gap_n = 0.0f, gap_x = 0.0f, gap_y = 0.0f;
    const float wf = (float)im.w, hf = (float)im.h;
    while (x >= 0.0f && y >= 0.0f && x < wf && y < hf) {
        steps++;
        const uint32_t v = im.at(__float2uint_rz(x), __float2uint_rz(y));
        if (v == 255u) {
            gap_n = 0.0f; gap_x = 0.0f; gap_y = 0.0f;
        } else if (gap_n >= max_gap) {
            x = gap_x; y = gap_y;
            break;
        } else if (gap_n == 0.0f) {
            gap_n = 1.0f; gap_x = x; gap_y = y;
        } else {
            gap_n = __fadd_rn(gap_n, 1.0f);
        }
        xo = __fadd_rn(xo, dx); yo = __fadd_rn(yo, dy);
        x = __fadd_rn(xo, px); y = __fadd_rn(yo, py);
    }
    // get_pixel_checked(x as u32, y as u32) == Some(0) — lib.rs:426-429 (D7)
    const uint32_t xi = __float2uint_rz(x), yi = __float2uint_rz(y);  // saturating, NaN -> 0
    if (xi < im.w && yi < im.h && im.at(xi, yi) == 0u) {
        ex = __fsub_rn(x, dx); ey = __fsub_rn(y, dy);
    } else {
        ex = px; ey = py;
    }
    const float ddx = __fsub_rn(px, ex), ddy = __fsub_rn(py, ey);
    return __fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy));  // distance_sqr, util/src/geometry.rs:62-68
}

// get_centre — vision-common/src/lsd.rs:5-44 (indices clamped where the reference is UB, D8)
__device__ __forceinline__ uint32_t px_clamped(const Img& im, float fx, float fy) {
    uint32_t x = __float2uint_rz(fx), y = __float2uint_rz(fy);
    x = min(x, im.w - 1u); y = min(y, im.h - 1u);
    return im.at(x, y);
}
__device__ void get_centre(const Img& im, float ptx, float pty, float& cx, float& cy) {
    const float MAXD = SMH_LSD_CENTRE_MAX_DIST;
    float left = ptx;
    while (left > 0.0f && fabsf(__fsub_rn(left, ptx)) < MAXD && px_clamped(im, left, pty) == 255u) left = __fsub_rn(left, 1.0f);
    float right = ptx;
    while (right < (float)(im.w - 1u) && fabsf(__fsub_rn(right, ptx)) < MAXD && px_clamped(im, right, pty) == 255u) right = __fadd_rn(right, 1.0f);
    float up = pty;
    while (up > 0.0f && fabsf(__fsub_rn(up, pty)) < MAXD && px_clamped(im, ptx, up) == 255u) up = __fsub_rn(up, 1.0f);
    float down = pty;
    while (down < (float)(im.h - 1u) && fabsf(__fsub_rn(down, pty)) < MAXD && px_clamped(im, ptx, down) == 255u) down = __fadd_rn(down, 1.0f);
    cx = __fdiv_rn(__fadd_rn(left, right), 2.0f);
    cy = __fdiv_rn(__fadd_rn(up, down), 2.0f);
}

// seed suppression — lsd.rs:46-58,84-89: squared distance to the INFINITE line through l (D14)
__device__ __forceinline__ bool suppressed_by(const float4& l, float x, float y) {
    const float dx = __fsub_rn(l.z, l.x), dy = __fsub_rn(l.w, l.y);
    float nx, ny;
    if (dx == 0.0f && dy == 0.0f) { nx = l.x; ny = l.y; }
    else {
        const float num = __fadd_rn(__fmul_rn(__fsub_rn(x, l.x), dx), __fmul_rn(__fsub_rn(y, l.y), dy));
        const float den = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        const float u = __fdiv_rn(num, den);
        nx = __fadd_rn(l.x, __fmul_rn(u, dx));
        ny = __fadd_rn(l.y, __fmul_rn(u, dy));
    }
    const float ex = __fsub_rn(x, nx), ey = __fsub_rn(y, ny);
    return __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)) < SMH_LSD_SUPPRESS_DIST_SQ;
}

// All 3600 rays from (px,py), spread over the CTA; the winner is the LAST index among the
// longest rays (rayon's right-biased reduce, vision-cpu/src/lib.rs:440-446, D6).
// Result broadcast through shared memory: s_best = {len, ex, ey}.
template <int T>
__device__ __forceinline__ void longest_ray_cta(const Img& im, const float2* __restrict__ rays, float px, float py,
                                                float max_gap, float* s_best, unsigned long long* s_keys,
                                                float2* s_ends, uint32_t& steps) {
    const int tid = threadIdx.x;
    unsigned long long best_key = 0ull;  // (len_bits << 32) | k ; len >= 0 so the bit pattern is monotone
    float bex = px, bey = py;
    bool any = false;
    for (int k = tid; k < SMH_LSD_RAYS; k += T) {
        const float2 d = __ldg(&rays[k]);
        float ex, ey;
        const float len = march(im, px, py, max_gap, d.x, d.y, ex, ey, steps);
        const unsigned long long key = ((unsigned long long)__float_as_uint(len) << 32) | (uint32_t)k;
        if (!any || key >= best_key) { best_key = key; bex = ex; bey = ey; any = true; }
    }
    // warp argmax on the 64-bit key
    unsigned long long wkey = best_key;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, wkey, o);
        wkey = other > wkey ? other : wkey;
    }
    const int warp = tid >> 5;
    if (any && best_key == wkey) { s_keys[warp] = wkey; s_ends[warp] = make_float2(bex, bey); }
    __syncthreads();
    if (tid == 0) {
        int bw = 0;
        for (int k = 1; k < T / 32; k++) if (s_keys[k] > s_keys[bw]) bw = k;
        s_best[0] = __uint_as_float((uint32_t)(s_keys[bw] >> 32));
        s_best[1] = s_ends[bw].x;
        s_best[2] = s_ends[bw].y;
    }
    __syncthreads();
}

constexpr int LSD_THREADS = 256;
constexpr int LSD_VPT = 4;  // 16-byte mask vectors per thread per scan chunk

template <bool STATS>
__global__ void __launch_bounds__(LSD_THREADS) k_lsd(const LsdArgs a) {
    constexpr int T = LSD_THREADS;
    __shared__ float4 s_lines[SMH_LSD_MAX_LINES];
    __shared__ uint32_t s_min[T / 32];
    __shared__ unsigned long long s_keys[T / 32];
    __shared__ float2 s_ends[T / 32];
    __shared__ float s_best[3];
    __shared__ float s_pt[2];
    __shared__ int s_n;

    const uint32_t f = blockIdx.x;
    const int tid = threadIdx.x;
    float* out_lines = a.lines + (uint64_t)f * SMH_LSD_MAX_LINES * 4;
    if (a.open && a.open[f] == 0) {  // map closed: the reference returns before any line search
        if (tid == 0) { a.n_lines[f] = 0; if (STATS) { a.stats[2 * f] = 0; a.stats[2 * f + 1] = 0; } }
        return;
    }
    const uint8_t* base = a.lsd + (uint64_t)f * a.stride;  // 16-byte aligned rows; pixel i at byte i+pad
    Img im{base + a.pad, a.pitch, a.w, a.h};
    if (tid == 0) s_n = 0;
    __syncthreads();

    uint32_t steps = 0, seeds = 0;
    int n = 0;
    const uint32_t vpr = (a.pad + a.w + 15u) / 16u;      // vectors per row
    const uint32_t rows_per_chunk = max(1u, (uint32_t)(T * LSD_VPT) / vpr);
    const uint32_t items = rows_per_chunk * vpr;

    for (uint32_t row0 = 0; row0 < a.h && n < SMH_LSD_MAX_LINES; row0 += rows_per_chunk) {
        // ---- load this thread's slice of the chunk and find its white, in-image pixels ----
        uint32_t alive[LSD_VPT];  // 16 bits per vector: candidate pixels still to be tried
        uint32_t vrow[LSD_VPT], vcol[LSD_VPT];
        bool any_white = false;
#pragma unroll
        for (int k = 0; k < LSD_VPT; k++) {
            const uint32_t item = (uint32_t)tid * LSD_VPT + k;  // consecutive items => increasing raster order
            alive[k] = 0; vrow[k] = 0; vcol[k] = 0;
            if (item < items) {
                const uint32_t r = row0 + item / vpr, c = item % vpr;
                vrow[k] = r; vcol[k] = c;
                if (r < a.h) {
                    const uint4 v = *reinterpret_cast<const uint4*>(base + (uint64_t)r * a.pitch + (uint64_t)c * 16u);
                    const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
                    if (v.x | v.y | v.z | v.w) {
#pragma unroll
                        for (int b = 0; b < 16; b++) {
                            const int i = (int)(c * 16u) + b - (int)a.pad;
                            const bool white = ((wd[b >> 2] >> (8 * (b & 3))) & 0xFFu) == 255u;
                            if (white && i >= 0 && i < (int)a.w) alive[k] |= 1u << b;
                        }
                    }
                }
            }
            any_white |= alive[k] != 0;
        }
        if (!__syncthreads_or(any_white)) continue;

        // ---- seeds of this chunk, in raster order ------------------------------------------
        while (true) {
            // first untried, unsuppressed candidate of this thread (its items are in raster order)
            uint32_t cand = 0xFFFFFFFFu;
#pragma unroll
            for (int k = 0; k < LSD_VPT; k++) {
                while (cand == 0xFFFFFFFFu && alive[k]) {
                    const int b = __ffs(alive[k]) - 1;
                    const uint32_t xi = vcol[k] * 16u + (uint32_t)b - a.pad, yi = vrow[k];
                    bool sup = false;
                    for (int l = 0; l < n && !sup; l++) sup = suppressed_by(s_lines[l], (float)xi, (float)yi);
                    if (sup) alive[k] &= ~(1u << b);  // lines only accumulate: suppressed for good
                    else cand = yi * a.w + xi;
                }
            }
            const uint32_t wmin = __reduce_min_sync(0xFFFFFFFFu, cand);
            if ((tid & 31) == 0) s_min[tid >> 5] = wmin;
            __syncthreads();
            uint32_t seed = 0xFFFFFFFFu;
#pragma unroll
            for (int k = 0; k < T / 32; k++) seed = min(seed, s_min[k]);
            if (seed == 0xFFFFFFFFu) break;  // uniform
            if (cand == seed) {
                // this thread owns the seed: retire it and publish its centre
#pragma unroll
                for (int k = 0; k < LSD_VPT; k++) {
                    if (alive[k]) {
                        const int b = __ffs(alive[k]) - 1;
                        if (vrow[k] * a.w + vcol[k] * 16u + (uint32_t)b - a.pad == seed) alive[k] &= ~(1u << b);
                    }
                }
                float cx, cy;
                get_centre(im, (float)(seed % a.w), (float)(seed / a.w), cx, cy);
                s_pt[0] = cx; s_pt[1] = cy;
            }
            __syncthreads();
            const float px = s_pt[0], py = s_pt[1];
            seeds++;
            longest_ray_cta<T>(im, a.rays, px, py, a.max_gap, s_best, s_keys, s_ends, steps);
            if (s_best[0] > SMH_LSD_MIN_LEN_SQ) {  // uniform
                if (tid == 0) {
                    float c1x, c1y;
                    get_centre(im, s_best[1], s_best[2], c1x, c1y);
                    s_lines[n] = make_float4(px, py, c1x, c1y);
                    out_lines[4 * n + 0] = px; out_lines[4 * n + 1] = py;
                    out_lines[4 * n + 2] = c1x; out_lines[4 * n + 3] = c1y;
                }
                n++;
                if (n == SMH_LSD_MAX_LINES) break;  // `break 'row`
            }
            __syncthreads();
        }
        __syncthreads();
    }
    if (tid == 0) a.n_lines[f] = (uint32_t)n;
    if (STATS) {
        steps = __reduce_add_sync(0xFFFFFFFFu, steps);
        if ((tid & 31) == 0) atomicAdd(&a.stats[2 * f + 1], (unsigned long long)steps);
        if (tid == 0) a.stats[2 * f] = seeds;
    }
}

// single find_longest_line call (Vision::find_longest_line); out = {p0x,p0y,p1x,p1y,len}
__global__ void __launch_bounds__(LSD_THREADS) k_longest(const uint8_t* lsd, uint32_t pitch, uint32_t pad, uint32_t w,
                                                         uint32_t h, const float2* rays, float px, float py,
                                                         float max_gap, float* out) {
    __shared__ unsigned long long s_keys[LSD_THREADS / 32];
    __shared__ float2 s_ends[LSD_THREADS / 32];
    __shared__ float s_best[3];
    Img im{lsd + pad, pitch, w, h};
    uint32_t steps = 0;
    longest_ray_cta<LSD_THREADS>(im, rays, px, py, max_gap, s_best, s_keys, s_ends, steps);
    if (threadIdx.x == 0) { out[0] = px; out[1] = py; out[2] = s_best[1]; out[3] = s_best[2]; out[4] = s_best[0]; }
}

// ------------------------------------------------------------------------------------------
// debug views (rare path): cropped_map after isolation / cropped brq as RGBA — K2/K4/K7 outputs
// ------------------------------------------------------------------------------------------
__global__ void k_debug_rgb(const uint8_t* frame, uint32_t pitch, uint32_t x0, uint32_t y0, uint32_t w, uint32_t h,
                            int isolate, uint8_t* out_rgba) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= w || j >= h) return;
    const uint32_t word = *reinterpret_cast<const uint32_t*>(frame + (uint64_t)(y0 + j) * pitch + (uint64_t)(x0 + i) * 4u);
    uint32_t b = word & 0xFFu, g = (word >> 8) & 0xFFu, r = (word >> 16) & 0xFFu;
    if (isolate && !smh_is_marker(r, g, b)) { r = g = b = 0; }  // vision-cpu/src/lib.rs:266-268
    reinterpret_cast<uint32_t*>(out_rgba)[(uint64_t)j * w + i] = r | (g << 8) | (b << 16) | 0xFF000000u;
}

// ------------------------------------------------------------------------------------------
// exhaustive colour tables + synthetic captures
// ------------------------------------------------------------------------------------------
// 2^24-bit table of the exact marker predicate: thread t fills word t (colours 32t .. 32t+31)
__global__ void k_build_marker_lut(uint32_t* lut) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t word = 0;
    for (uint32_t k = 0; k < 32; k++) {
        const uint32_t c = t * 32u + k;
        word |= (smh_is_marker(c >> 16, (c >> 8) & 0xFFu, c & 0xFFu) ? 1u : 0u) << k;
    }
    lut[t] = word;
}

__global__ void k_colour_table_fast(uint32_t which, const uint32_t* lut, uint8_t* out) {
    // the exact code paths k_pixels uses: which 0 = fast-reject + table bit, 1 = luma8_bgra
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t r = c >> 16, g = (c >> 8) & 0xFFu, b = c & 0xFFu;
    const uint32_t word = c | 0xFF000000u;
    const uint32_t mx = max(r, max(g, b)), mn = min(r, min(g, b));
    out[c] = which == 0 ? (uint8_t)marker_bit(word, mx, mn, lut) : (uint8_t)luma8_bgra(word);
}

__global__ void k_colour_table(uint32_t which, uint8_t* out) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t r = c >> 16, g = (c >> 8) & 0xFFu, b = c & 0xFFu;
    uint8_t v;
    if (which == 0) v = smh_is_marker(r, g, b) ? 1 : 0;
    else if (which == 1) v = (uint8_t)smh_luma8(r, g, b);
    else if (which == 2) v = smh_button_red(r, g, b) ? 1 : 0;
    else v = (uint8_t)((smh_ocr_core(r, g, b) ? 1 : 0) | ((!smh_ocr_core(r, g, b) && smh_ocr_edge(r, g, b)) ? 2 : 0));
    out[c] = v;
}

struct SynthArgs {
    const smh_synth_frame_desc* desc;  // [n]
    const smh_synth_segment* segs;     // [n][n_segments]
    const smh_synth_blip* blips;       // [n][n_blips]
    uint32_t n_segments, n_blips, W, H;
    uint8_t* out;                      // [n][H][W] BGRA words
};
__global__ void k_synth(const SynthArgs a) {
    const uint32_t f = blockIdx.z;
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= a.W) return;
    const uint32_t word = smh_synth_pixel(&a.desc[f], a.segs + (uint64_t)f * a.n_segments,
                                          a.blips + (uint64_t)f * a.n_blips, (int32_t)x, (int32_t)y);
    reinterpret_cast<uint32_t*>(a.out)[((uint64_t)f * a.H + y) * a.W + x] = word;
}

}  // namespace smh
