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
// 16-byte asynchronous global->shared copy (LDGSTS) bypassing L1, plus group bookkeeping
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ uint4 lds_v4(uint32_t smem_addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(smem_addr));
    return r;
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
    uint32_t n_frames, nstrips;  // strips of PX_STRIP vector columns per row
    const uint32_t* lut;  // marker predicate, 1 bit per colour (r<<16|g<<8|b)
};



// Channel k of a BGRA word as the bit pattern of the float 2^23 + value (0x4B0000vv).  The same
// word serves three purposes: FADD -2^23 turns it into the exact float (full-rate replacement for
// I2F), unsigned min/max of these words order like the channel values, and the low byte is the value.
constexpr uint32_t BIAS = 0x4B000000u;
template <int K>
__device__ __forceinline__ uint32_t chan(uint32_t word) { return __byte_perm(word, BIAS, 0x7540 + K); }
__device__ __forceinline__ float unbias(uint32_t c) { return __fsub_rn(__uint_as_float(c), 8388608.0f); }

// luma8 (image 0.23.14 rgb_to_luma: unfused f32, truncating) as a biased word 0x4B0000LL: the
// truncation is a round-down add of 2^23 (l < 256).  Same bits as smh_luma8 for all 2^24 colours
// (device-side exhaustive test in tests/test_gpu_parity.py).
__device__ __forceinline__ uint32_t luma_biased(uint32_t cr, uint32_t cg, uint32_t cb) {
    const float l = __fadd_rn(__fadd_rn(__fmul_rn(SMH_LUMA_R, unbias(cr)), __fmul_rn(SMH_LUMA_G, unbias(cg))),
                              __fmul_rn(SMH_LUMA_B, unbias(cb)));
    return __float_as_uint(__fadd_rd(l, 8388608.0f));
}
__device__ __forceinline__ uint32_t luma8_bgra(uint32_t word) {
    return luma_biased(chan<2>(word), chan<1>(word), chan<0>(word)) & 0xFFu;
}

// Marker predicate of a BGRA word given max/min channel values: integer fast-reject (a necessary
// condition, smh_marker_prefilter), then one bit of the 2^24-entry table that k_build_marker_lut
// fills with the exact float predicate (2 MB, L2 resident).
__device__ __forceinline__ uint32_t marker_bit(uint32_t word, uint32_t mx, uint32_t mn, const uint32_t* __restrict__ lut) {
    if (mx >= 178u && 100u * (mx - mn) >= 34u * mx) {
        const uint32_t c = word & 0x00FFFFFFu;  // r<<16 | g<<8 | b
        return (__ldg(lut + (c >> 5)) >> (c & 31u)) & 1u;
    }
    return 0u;
}

struct PxOut {
    uint32_t m4, c4, e4;  // per-pixel flag nibbles: marker colour, OCR core, OCR edge candidate
    uint4 ui;             // ui_map words
    uint32_t ocrw;        // provisional ocr_out bytes (255 or 255-luma)
    uint32_t scw;         // scales bytes (luma != 0 ? 255 : 0)
};

// An opaque 0xFF000000 so that PRMT takes it as a register operand and keeps its selector immediate.
__device__ __forceinline__ uint32_t opaque_alpha() {
    uint32_t c;
    asm("mov.u32 %0, 0xFF000000;" : "=r"(c));
    return c;
}

// luma8 of a BGRA word through I2F.U8 byte-select conversions (one XU op per channel instead of
// PRMT + FADD), result biased (0x4B0000LL); bit-identical to smh_luma8 (device-side exhaustive test).
__device__ __forceinline__ uint32_t luma_biased_i2f(uint32_t w) {
    const float r = (float)((w >> 16) & 0xFFu), g = (float)((w >> 8) & 0xFFu), b = (float)(w & 0xFFu);
    const float l = __fadd_rn(__fadd_rn(__fmul_rn(SMH_LUMA_R, r), __fmul_rn(SMH_LUMA_G, g)), __fmul_rn(SMH_LUMA_B, b));
    return __float_as_uint(__fadd_rd(l, 8388608.0f));
}

// rare path shared by both vector kinds: exact marker bits of the 4 pixels
__device__ __noinline__ uint32_t marker_nibble(uint4 v, const uint32_t* __restrict__ lut) {
    const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
    uint32_t m4 = 0;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const uint32_t cb = wd[e] & 0xFFu, cg = (wd[e] >> 8) & 0xFFu, cr = (wd[e] >> 16) & 0xFFu;
        m4 |= marker_bit(wd[e], max(cr, max(cg, cb)), min(cr, min(cg, cb)), lut) << e;
    }
    return m4;
}

// ---- vector with no pixel inside the OCR quadrant (flags: marker only), owned row -------------
// Marker fast-reject for the whole vector: a marker colour has max(r,g,b) >= 178, i.e. some colour
// byte with bit 7 set and low 7 bits >= 50; SIMD-in-register test over the three colour bytes.
__device__ __forceinline__ uint32_t bright_bytes(uint32_t w) {  // nonzero iff some of b,g,r is >= 178
    return ((w & 0x007F7F7Fu) + 0x004E4E4Eu) & w & 0x00808080u;
}

template <bool GRAY>
__device__ __forceinline__ void vec_plain(const uint4& v, const uint32_t* __restrict__ lut, uint32_t alpha, PxOut& o) {
    if (GRAY) {  // (L,L,L,255) — vision-cpu/src/lib.rs:150-157
        o.ui.x = __byte_perm(luma_biased_i2f(v.x), alpha, 0x7000);
        o.ui.y = __byte_perm(luma_biased_i2f(v.y), alpha, 0x7000);
        o.ui.z = __byte_perm(luma_biased_i2f(v.z), alpha, 0x7000);
        o.ui.w = __byte_perm(luma_biased_i2f(v.w), alpha, 0x7000);
    } else {     // (r,g,b,255) — vision-cpu/src/lib.rs:158-163
        o.ui.x = __byte_perm(v.x, alpha, 0x7012);
        o.ui.y = __byte_perm(v.y, alpha, 0x7012);
        o.ui.z = __byte_perm(v.z, alpha, 0x7012);
        o.ui.w = __byte_perm(v.w, alpha, 0x7012);
    }
    o.m4 = 0;
    if (bright_bytes(v.x) | bright_bytes(v.y) | bright_bytes(v.z) | bright_bytes(v.w)) o.m4 = marker_nibble(v, lut);  // rare
}

// marker flags only (halo rows: nothing is stored)
__device__ __forceinline__ uint32_t vec_marker_only(const uint4& v, const uint32_t* __restrict__ lut) {
    if (bright_bytes(v.x) | bright_bytes(v.y) | bright_bytes(v.z) | bright_bytes(v.w)) return marker_nibble(v, lut);
    return 0u;
}

// ---- vector with pixels inside the OCR quadrant (flags: marker, core, edge; the caller masks
//      out what a pixel may not carry: outside the quadrant, above the search clip) ------------
template <bool GRAY>
__device__ __forceinline__ void vec_quadrant(const uint4& v, bool owned, const uint32_t* __restrict__ lut, uint32_t alpha, PxOut& o) {
    const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
    uint32_t uo[4], dmin = 0xFFFFFFFFu, lmin = 0xFFFFFFFFu;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const uint32_t cb = chan<0>(wd[e]), cg = chan<1>(wd[e]), cr = chan<2>(wd[e]);
        dmin = min(dmin, max(cr, max(cg, cb)) - min(cr, min(cg, cb)));
        const uint32_t lb = owned ? luma_biased(cr, cg, cb) : BIAS + 1u;
        lmin = min(lmin, lb);
        if (GRAY) uo[e] = __byte_perm(lb, alpha, 0x7000);
        else uo[e] = __byte_perm(wd[e], alpha, 0x7012);
    }
    o.ui = make_uint4(uo[0], uo[1], uo[2], uo[3]);
    o.m4 = 0; o.c4 = 0; o.e4 = 0; o.ocrw = 0xFFFFFFFFu; o.scw = 0xFFFFFFFFu;
    if (bright_bytes(v.x) | bright_bytes(v.y) | bright_bytes(v.z) | bright_bytes(v.w)) o.m4 = marker_nibble(v, lut);  // rare
    // OCR classes — vision-cpu/src/lib.rs:39-53,188-193; mono = sum |ci-cj| = 4*(max-min)
    if (dmin <= SMH_OCR_MONO_EDGE / 4u) {  // rare: a near-grey pixel
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const uint32_t cb = wd[e] & 0xFFu, cg = (wd[e] >> 8) & 0xFFu, cr = (wd[e] >> 16) & 0xFFu;
            const uint32_t mx = max(cr, max(cg, cb)), mn = min(cr, min(cg, cb));
            if (mx == mn && mn >= SMH_OCR_BRIGHT) {
                o.c4 |= 1u << e;
                if (owned) o.ocrw ^= luma8_bgra(wd[e]) << (8 * e);  // 255 - luma
            } else if (4u * (mx - mn) <= SMH_OCR_MONO_EDGE && mn >= SMH_OCR_BRIGHT_EDGE) {
                o.e4 |= 1u << e;
            }
        }
    }
    if (lmin == BIAS) {  // rare: a pixel with luma 0
#pragma unroll
        for (int e = 0; e < 4; e++)
            if (luma8_bgra(wd[e]) == 0u) o.scw &= ~(0xFFu << (8 * e));
    }
}

// Fused pixel pass, warp-autonomous.  A work item is (frame, band of rows, strip of 30 vector
// columns) or a frame's button test; every warp takes one item and needs no shared memory or
// block barrier.  Lane L owns the 16-byte vector column q = 30*strip - 1 + L: lanes 0 and 31 are a
// one-vector halo (enough for the 3x3 dilation and the 7x7 OCR window) that classify but never
// store.  A lane walks down the band with PX_DEPTH rows of 16-byte cp.async copies in flight into
// a per-warp shared-memory ring (each lane only ever reads back its own 16 bytes, so no barrier):
//   phase 1  classify the 4 pixels of row jn, store ui_map / provisional ocr_out / scales, and
//            shift the row's flag nibbles (marker, OCR core, OCR edge) into three window
//            registers that hold the last 5 / 7 / 4 rows of the lane's own column;
//   phase 2  (3 rows behind) dilate the marker bits with a 3x3 cross — up/down from the window,
//            left/right by warp shuffle — into lsd_image, and run the 7x7 neighbour search for
//            the rare OCR edge pixels, again by shuffle.
constexpr int PX_THREADS = 128;
constexpr int PX_WARPS = PX_THREADS / 32;
constexpr int PX_STRIP = 30;  // inner vector columns per warp
#ifndef PX_MIN_BLOCKS
#define PX_MIN_BLOCKS 8  // register cap 64/thread -> 32 warps per SM
#endif
#ifndef PX_DEPTH
#define PX_DEPTH 3       // rows of 128-bit copies in flight per lane (3 measured best on B200; 5-7 are ~4% slower)
#endif
constexpr int PX_SLOTS = 8;  // ring slots per warp (power of two > PX_DEPTH), 512 B each
constexpr int PX_SMEM_BYTES = PX_WARPS * PX_SLOTS * 512;

template <bool GRAY>
__global__ void __launch_bounds__(PX_THREADS, PX_MIN_BLOCKS) k_pixels(const PixelArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t item = blockIdx.x * PX_WARPS + (threadIdx.x >> 5);
    const uint32_t per_frame = a.nbands * a.nstrips + 1u;
    const uint32_t f = item / per_frame;
    if (f >= a.n_frames) return;
    const uint32_t sub = item - f * per_frame;
    const uint8_t* frame = a.frames + (uint64_t)f * a.frame_stride;

    if (sub == per_frame - 1u) {
        // ---- close-deployment button test (vision-cpu/src/lib.rs:115-135) -----------------
        uint32_t cnt = 0;
        const uint32_t npx = a.btn_w * a.btn_h;
        for (uint32_t idx = lane; idx < npx; idx += 32) {
            const uint32_t y = idx / a.btn_w, x = idx - y * a.btn_w;
            const uint32_t word = __ldg(reinterpret_cast<const uint32_t*>(
                frame + (uint64_t)(a.btn_y + y) * a.pitch + (uint64_t)(a.btn_x + x) * 4u));
            cnt += smh_button_red((word >> 16) & 0xFFu, (word >> 8) & 0xFFu, word & 0xFFu) ? 1u : 0u;
        }
        cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
        if (lane == 0) {
            a.red[f] = cnt;
            // `red_pixels as f32 / (w*h) as f32 < 0.65` -> map closed
            a.open[f] = __fdiv_rn((float)cnt, (float)npx) < SMH_BTN_RED_RATIO_THRESHOLD ? 0 : 1;
        }
        return;
    }

    const uint32_t band = sub / a.nstrips, strip = sub - band * a.nstrips;
    const int r0 = (int)(band * a.rows_per_band);
    const int r1 = min(r0 + (int)a.rows_per_band, (int)a.h);
    const int bqh = (int)a.bqh, bqw = (int)a.bqw;
    const int halo = (r1 > bqh && r0 < 2 * bqh) ? SMH_OCR_RADIUS : 1;
    const uint32_t start_y = a.start_y ? a.start_y[f] : a.start_y_uniform;
    const uint32_t* __restrict__ lut = a.lut;
    const uint32_t alpha = opaque_alpha();

    // ---- per-lane column constants ----------------------------------------------------------
    const int q = (int)strip * PX_STRIP - 1 + lane;
    const bool active = q >= 0 && q < (int)a.nvec;
    const bool inner = active && lane >= 1 && lane <= PX_STRIP;  // lanes that store
    const int i_lo = 4 * q - (int)a.pad;
    uint32_t vm4 = 0, cm4 = 0, em4 = 0;  // nibble masks: pixel in the image / may carry a core bit / an edge bit
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const int i = i_lo + e;
        const bool valid = active && i >= 0 && i < (int)a.w;
        const bool inq = valid && i >= bqw && i < 2 * bqw;
        vm4 |= valid ? (1u << e) : 0u;
        em4 |= inq ? (1u << e) : 0u;
        // an OCR core pixel is visible to a neighbour's search only below the clip min(x+3, w-3) (D10)
        cm4 |= (inq && i - bqw <= bqw - SMH_OCR_RADIUS) ? (1u << e) : 0u;
    }
    const bool col_brq = em4 != 0u;  // some pixel inside the quadrant columns
    const bool brq_strip = __any_sync(0xFFFFFFFFu, col_brq && inner);
    const uint32_t qo = active ? (uint32_t)q : 0u;
    const uint32_t qpos = (uint32_t)((int)a.qpad - (int)a.pad - bqw + 4 * (int)qo);  // byte position in a quadrant row (used when col_brq)
    uint8_t* ocr_f = a.ocr + (uint64_t)f * a.q_stride;
    uint8_t* sc_f = a.scales + (uint64_t)f * a.q_stride;

    // rows this item reads: [j_lo, j_hi); rows it owns (stores): [r0, r1)
    const int j_lo = max(r0 - halo, 0), j_hi = min(r1 + halo, (int)a.h);
    const int j_first = r0 - halo;  // the windows also slide over rows outside the image (flags 0)

    // running per-lane pointers: row jn of the capture and of ui_map, row jn-3 of lsd_image
    const uint8_t* src_p = frame + ((int64_t)a.map_y + j_first) * (int64_t)a.pitch + ((uint64_t)a.gx0 + 4u * qo) * 4u;
    uint8_t* ui_p = a.ui + (uint64_t)f * a.ui_stride + (int64_t)j_first * (int64_t)a.ui_pitch + qo * 16u;
    uint8_t* lsd_p = a.lsd + (uint64_t)f * a.lsd_stride + (int64_t)(j_first - SMH_OCR_RADIUS) * (int64_t)a.lsd_pitch + qo * 4u;

    uint32_t mwin = 0, cwin = 0, ewin = 0;  // flag nibbles of rows jn, jn-1, ... (newest in bits 0-3)

    auto row = [&](const uint4& cur, int jn) {
        // ---------------- phase 1: row jn -------------------------------------------------------
        uint32_t m4 = 0, c4 = 0, e4 = 0;
        if (active && jn >= j_lo && jn < j_hi) {
            const bool owned = jn >= r0 && jn < r1;
            const int jq = jn - bqh;
            if ((uint32_t)jq < (uint32_t)bqh && col_brq) {
                PxOut o;
                vec_quadrant<GRAY>(cur, owned, lut, alpha, o);
                c4 = jq <= bqh - SMH_OCR_RADIUS ? o.c4 & cm4 : 0u;  // rows above the vertical clip only (D10)
                e4 = o.e4 & em4;
                m4 = o.m4;
                if (owned && inner) {
                    stg_v4(ui_p, o.ui);
                    stg_u32(ocr_f + ((uint32_t)jq * a.q_pitch + qpos), o.ocrw);
                    if ((uint32_t)jq >= start_y) stg_u32(sc_f + ((uint32_t)jq * a.q_pitch + qpos), o.scw);
                }
            } else if (owned) {
                PxOut o;
                vec_plain<GRAY>(cur, lut, alpha, o);
                m4 = o.m4;
                if (inner) stg_v4(ui_p, o.ui);
            } else {
                m4 = vec_marker_only(cur, lut);  // halo row: flags only
            }
            m4 &= vm4;
        }
        mwin = (mwin << 4) | m4;
        if (brq_strip) { cwin = (cwin << 4) | c4; ewin = (ewin << 4) | e4; }
        ui_p += a.ui_pitch; lsd_p += a.lsd_pitch;

        // ---------------- phase 2: row jc = jn - 3 ----------------------------------------------
        const int jc = jn - SMH_OCR_RADIUS;
        if (jc < r0 || jc >= r1) return;  // warp-uniform
        const uint32_t mc = (mwin >> 12) & 0xFu;  // centre row; up = bits 16-19, down = bits 8-11 (0 outside the image)
        const uint32_t l = __shfl_up_sync(0xFFFFFFFFu, mc, 1), rr = __shfl_down_sync(0xFFFFFFFFu, mc, 1);
        const uint32_t d4 = (mc | (mc << 1) | (mc >> 1) | (l >> 3) | (rr << 3) | (mwin >> 16) | (mwin >> 8)) & vm4;
        // nibble -> 4 bytes of 0/255 (pixels outside [0, w) stay 0); lsd_p already points one row ahead
        if (inner) stg_u32(lsd_p - a.lsd_pitch, ((d4 * 0x00204081u) & 0x01010101u) * 255u);

        const int jqc = jc - bqh;
        if (!brq_strip || (uint32_t)jqc >= (uint32_t)bqh) return;  // warp-uniform
        const uint32_t eb = inner ? (ewin >> 12) & 0xFu : 0u;
        if (__any_sync(0xFFFFFFFFu, eb != 0u)) {
            // rare: an OCR edge pixel keeps its value iff a core pixel lies in its 7x7 window
            // (vision-cpu/src/lib.rs:196-216); rows/columns outside the quadrant carry no core bit
            const uint32_t lwin = __shfl_up_sync(0xFFFFFFFFu, cwin, 1), rwin = __shfl_down_sync(0xFFFFFFFFu, cwin, 1);
            uint32_t found = 0;  // bit e: pixel e has a core neighbour
#pragma unroll
            for (int k = 0; k < 7; k++) {
                // core bits of columns -4 .. +7 around the lane's first pixel, row jn - k
                const uint32_t cm = ((lwin >> (4 * k)) & 0xFu) | (((cwin >> (4 * k)) & 0xFu) << 4) | (((rwin >> (4 * k)) & 0xFu) << 8);
#pragma unroll
                for (int e = 0; e < 4; e++) found |= (cm & (0x7Fu << (e + 1))) ? (1u << e) : 0u;
            }
            const uint32_t todo = eb & found;
#pragma unroll 1
            for (int e = 0; e < 4; e++) {
                if ((todo >> e) & 1u) {
                    const uint32_t word = __ldg(reinterpret_cast<const uint32_t*>(
                        frame + (uint64_t)(a.map_y + (uint32_t)jc) * a.pitch + (uint64_t)(a.map_x + (uint32_t)(i_lo + e)) * 4u));
                    ocr_f[(uint32_t)jqc * a.q_pitch + qpos + e] = (uint8_t)(255u - luma8_bgra(word));
                }
            }
        }
    };

    // software pipeline through the warp's shared-memory ring: row jn + PX_DEPTH is requested while
    // row jn is consumed; one commit group per row (empty when the row is not read)
    extern __shared__ __align__(16) uint8_t px_ring[];
    const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(px_ring) + (uint32_t)(threadIdx.x >> 5) * (PX_SLOTS * 512u) + (uint32_t)lane * 16u;
    const uint8_t* fetch_p = src_p;  // row j_first + k of the lane's column
#pragma unroll
    for (int k = 0; k < PX_DEPTH; k++) {
        const int j = j_first + k;
        if (active && j >= j_lo && j < j_hi) cp_async16(ring0 + (uint32_t)(k & (PX_SLOTS - 1)) * 512u, fetch_p);
        cp_async_commit();
        fetch_p += a.pitch;
    }
    uint32_t slot = 0;
#pragma unroll 1
    for (int jn = j_first; jn - SMH_OCR_RADIUS < r1; jn++) {
        {
            const int j = jn + PX_DEPTH;
            if (active && j >= j_lo && j < j_hi) cp_async16(ring0 + ((slot + PX_DEPTH) & (PX_SLOTS - 1)) * 512u, fetch_p);
            cp_async_commit();
            fetch_p += a.pitch;
        }
        cp_async_wait<PX_DEPTH>();  // the group of row jn has landed
        const uint4 cur = lds_v4(ring0 + slot * 512u);
        slot = (slot + 1) & (PX_SLOTS - 1);
        row(cur, jn);
    }
}

// ------------------------------------------------------------------------------------------
// line search
// ------------------------------------------------------------------------------------------
struct LsdArgs {
    const uint8_t* lsd; uint64_t stride; uint32_t pitch, pad, w, h;
    const float2* rays;  // 3600 (cos, sin) pairs built on the host with libm (D11)
    float max_gap;
    int gap_limit;        // ceil(max_gap): `gap.0 >= max_gap` on the integer-valued counter
    int prune;            // 1: skip seeds proven unable to yield a line (isolated blobs)
    uint32_t n_frames;
    uint32_t* counter;    // persistent-CTA frame dispenser (zeroed before the launch)
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

// Same automaton for seeds that lie inside the image (all coordinates within (-2^22, 2^22)) and
// gap_limit = ceil(max_gap) >= 1: floor() by a round-down add of 1.5*2^23 (full-rate FADD instead
// of F2I), both bounds in one unsigned compare each, integer gap counter.  With gap_limit >= 1 a
// ray can only stop on a gap that has a recorded start, so the reference's reset of the saved
// position to (0,0) on white pixels is unobservable and dropped.  Bit-identical end points to
// march() — floor(x) in [0,w) <=> 0 <= x < w, and floor == trunc for x >= 0.
template <bool STATS>
__device__ __forceinline__ float march_fast(const uint8_t* __restrict__ img, uint32_t pitch, uint32_t w, uint32_t h,
                                            float px, float py, int gap_limit, float dx, float dy,
                                            float& ex, float& ey, uint32_t& steps) {
    float x = px, y = py, xo = 0.0f, yo = 0.0f, gap_x = px, gap_y = py;
    int gap = 0;
    bool broke = false;
    while (true) {
        const uint32_t ix = (uint32_t)(__float_as_int(__fadd_rd(x, 12582912.0f)) - 0x4B400000);
        const uint32_t iy = (uint32_t)(__float_as_int(__fadd_rd(y, 12582912.0f)) - 0x4B400000);
        if (ix >= w || iy >= h) break;
        if (STATS) steps++;
        const bool black = __ldg(img + (iy * pitch + ix)) != 255u;
        if (black && gap >= gap_limit) { broke = true; break; }
        if (black && gap == 0) { gap_x = x; gap_y = y; }
        gap = black ? gap + 1 : 0;
        xo = __fadd_rn(xo, dx); yo = __fadd_rn(yo, dy);
        x = __fadd_rn(xo, px); y = __fadd_rn(yo, py);
    }
    if (broke) { x = gap_x; y = gap_y; }
    const uint32_t xi = __float2uint_rz(x), yi = __float2uint_rz(y);  // `as u32`: saturating (D7)
    if (xi < w && yi < h && __ldg(img + (yi * pitch + xi)) == 0u) {
        ex = __fsub_rn(x, dx); ey = __fsub_rn(y, dy);
    } else {
        ex = px; ey = py;
    }
    const float ddx = __fsub_rn(px, ex), ddy = __fsub_rn(py, ey);
    return __fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy));
}

// get_centre evaluated by one warp in parallel (every lane returns the result): lane = 8*dir + t
// tests the loop condition of step t of direction dir (left, right, up, down); the walk stops at
// the first failing t, so the stop index is the number of trailing ones of the ballot byte.
// Each lane rebuilds its v by t successive +-1.0 adds, because the reference's `(v - pt).abs() < 5`
// sees the accumulated rounding: across a binade boundary (e.g. 511.6 -> 512.6) five increments can
// leave |v - pt| just below 5 and the walk takes a sixth step.
__device__ __forceinline__ void get_centre_warp(const Img& im, float ptx, float pty, float& cx, float& cy) {
    const int lane = threadIdx.x & 31, dir = lane >> 3, t = lane & 7;
    const float pt = dir < 2 ? ptx : pty;
    const float step = (dir & 1) ? 1.0f : -1.0f;
    float v = pt;
    for (int i = 0; i < t; i++) v = __fadd_rn(v, step);
    bool ok = fabsf(__fsub_rn(v, pt)) < SMH_LSD_CENTRE_MAX_DIST;
    if (dir == 0) ok = ok && v > 0.0f && px_clamped(im, v, pty) == 255u;
    else if (dir == 1) ok = ok && v < (float)(im.w - 1u) && px_clamped(im, v, pty) == 255u;
    else if (dir == 2) ok = ok && v > 0.0f && px_clamped(im, ptx, v) == 255u;
    else ok = ok && v < (float)(im.h - 1u) && px_clamped(im, ptx, v) == 255u;
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, ok);
    // value after k steps = the v of lane 8*dir + k (k <= 7: lane t = 7 always fails the distance test)
    const int kl = __ffs(~(bal & 0xFFu)) - 1, kr = __ffs(~((bal >> 8) & 0xFFu)) - 1;
    const int ku = __ffs(~((bal >> 16) & 0xFFu)) - 1, kd = __ffs(~((bal >> 24) & 0xFFu)) - 1;
    const float left = __shfl_sync(0xFFFFFFFFu, v, kl), right = __shfl_sync(0xFFFFFFFFu, v, 8 + kr);
    const float up = __shfl_sync(0xFFFFFFFFu, v, 16 + ku), down = __shfl_sync(0xFFFFFFFFu, v, 24 + kd);
    cx = __fdiv_rn(__fadd_rn(left, right), 2.0f);
    cy = __fdiv_rn(__fadd_rn(up, down), 2.0f);
}

// Exact pruning of hopeless seeds.  A ray stops at its K-th consecutive non-white sample,
// K = gap_limit + 1.  Samples move by at most one pixel per step in Chebyshev distance d from
// the seed pixel (monotonically; a float rounding can make one step span two pixels, hence the
// +1 below).  If no white pixel has d in [a+1, a+K+1], every ray stops inside that ring, its
// restored gap start lies within d <= a+2 and its squared length is < 2(a+3)^2; a ray can only
// leave through the left/top border first (the D7 quirk that still yields a long line) when the
// seed is closer than a+K+2 to that border, and then its length is < 2(a+K+2)^2.  When the bound
// is <= 2500 = SMH_LSD_MIN_LEN_SQ the seed cannot produce a line (lsd.rs:95) and the 3600-ray
// search is skipped — find_lines' output is unchanged.  Uniform result for the whole CTA.
template <int T>
__device__ __forceinline__ bool seed_is_isolated(const uint8_t* __restrict__ base, uint32_t pitch, uint32_t pad, uint32_t w,
                                                 uint32_t h, float px, float py, int gap_limit,
                                                 unsigned long long* s_mask) {
    if (gap_limit < 0 || gap_limit > 22) return false;  // keep the inspected window <= 113 px wide
    const int K = gap_limit + 1;
    const int R = 32 + K + 1;
    const int cx = (int)floorf(px), cy = (int)floorf(py);
    const int x0 = max(cx - R, 0), x1 = min(cx + R, (int)w - 1);
    const int y0 = max(cy - R, 0), y1 = min(cy + R, (int)h - 1);
    const int wlo = (x0 + (int)pad) >> 2, whi = (x1 + (int)pad) >> 2;
    const int nw = whi - wlo + 1, total = nw * (y1 - y0 + 1);
    // *s_mask was zeroed one seed ago (double-buffered by the caller)
    unsigned long long m = 0ull;
    for (int t = threadIdx.x; t < total; t += T) {
        const int r = t / nw, c = t - r * nw;
        const int y = y0 + r;
        const uint32_t word = __ldg(reinterpret_cast<const uint32_t*>(base + (uint32_t)y * pitch) + (wlo + c));
        if (word) {
            const int dy = abs(y - cy);
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int x = (wlo + c) * 4 + b - (int)pad;
                if (((word >> (8 * b)) & 0xFFu) == 255u && x >= x0 && x <= x1) m |= 1ull << max(abs(x - cx), dy);
            }
        }
    }
    const uint32_t lo = __reduce_or_sync(0xFFFFFFFFu, (uint32_t)m), hi = __reduce_or_sync(0xFFFFFFFFu, (uint32_t)(m >> 32));
    if ((threadIdx.x & 31) == 0 && (lo | hi)) atomicOr(s_mask, ((unsigned long long)hi << 32) | lo);
    __syncthreads();
    const unsigned long long mask = *s_mask;
    const unsigned long long ring = (1ull << (K + 1)) - 1ull;
    for (int a = 0; a <= 32; a++) {
        if ((mask & (ring << (a + 1))) == 0ull) {
            const bool border_exit = cx < a + K + 2 || cy < a + K + 2;
            return !border_exit || a + K <= 33;
        }
    }
    return false;
}

// All 3600 rays from (px,py), spread over the CTA; the winner is the LAST index among the
// longest rays (rayon's right-biased reduce, vision-cpu/src/lib.rs:440-446, D6).
// Every thread returns best = {len, ex, ey} (one barrier; s_keys/s_ends may be reused after the next one).
template <int T, bool FAST, bool STATS>
__device__ __forceinline__ void longest_ray_cta(const Img& im, const float2* __restrict__ rays, float px, float py,
                                                float max_gap, int gap_limit, float (&best)[3], unsigned long long* s_keys,
                                                float2* s_ends, uint32_t& steps) {
    const int tid = threadIdx.x;
    unsigned long long best_key = 0ull;  // (len_bits << 32) | k ; len >= 0 so the bit pattern is monotone
    float bex = px, bey = py;
    bool any = false;
    for (int k = tid; k < SMH_LSD_RAYS; k += T) {
        const float2 d = __ldg(&rays[k]);
        float ex, ey;
        const float len = FAST ? march_fast<STATS>(im.p, im.pitch, im.w, im.h, px, py, gap_limit, d.x, d.y, ex, ey, steps)
                               : march(im, px, py, max_gap, d.x, d.y, ex, ey, steps);
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
    int bw = 0;
    unsigned long long bk = s_keys[0];
#pragma unroll
    for (int k = 1; k < T / 32; k++) { const unsigned long long o = s_keys[k]; if (o > bk) { bk = o; bw = k; } }
    best[0] = __uint_as_float((uint32_t)(bk >> 32));
    best[1] = s_ends[bw].x;
    best[2] = s_ends[bw].y;
}

constexpr int LSD_THREADS = 256;
constexpr int LSD_VPT = 4;  // 16-byte mask vectors per thread per scan chunk

// lsd::find_lines for whole frames (vision-common/src/lsd.rs:60-105).  One CTA works on one frame
// at a time (persistent CTAs pull frames from a counter).  Per frame: the mask is scanned in
// raster order in chunks held in registers as per-thread bitmaps of white pixels; pixels are
// struck off as soon as a line suppresses them, so the next seed is just the lowest remaining
// bit (block-wide min).  Per seed: get_centre (per warp, redundantly), the exact isolated-seed
// test, then the 3600-ray search spread over the CTA with a last-index argmax.
template <bool STATS>
__global__ void __launch_bounds__(LSD_THREADS, 4) k_lsd(const LsdArgs a) {
    constexpr int T = LSD_THREADS;
    __shared__ float4 s_lines[SMH_LSD_MAX_LINES];
    __shared__ uint32_t s_min[T / 32];
    __shared__ unsigned long long s_keys[T / 32];
    __shared__ float2 s_ends[T / 32];
    __shared__ unsigned long long s_mask[2];
    __shared__ uint32_t s_frame;

    const int tid = threadIdx.x;
    const uint32_t vpr = (a.pad + a.w + 15u) / 16u;  // 16-byte vectors per mask row
    const uint32_t rows_per_chunk = max(1u, (uint32_t)(T * LSD_VPT) / vpr);
    const uint32_t items = rows_per_chunk * vpr;
    const bool fast = a.gap_limit >= 1;  // march_fast's precondition

    while (true) {
        __syncthreads();
        if (tid == 0) { s_frame = atomicAdd(a.counter, 1u); s_mask[0] = 0ull; s_mask[1] = 0ull; }
        __syncthreads();
        const uint32_t f = s_frame;
        if (f >= a.n_frames) break;
        float* out_lines = a.lines + (uint64_t)f * SMH_LSD_MAX_LINES * 4;
        if (a.open && a.open[f] == 0) {  // map closed: the reference returns before any line search
            if (tid == 0) { a.n_lines[f] = 0; if (STATS) { a.stats[3 * f] = 0; a.stats[3 * f + 1] = 0; a.stats[3 * f + 2] = 0; } }
            continue;
        }
        const uint8_t* base = a.lsd + (uint64_t)f * a.stride;  // 16-byte aligned rows; pixel i at byte i+pad
        const Img im{base + a.pad, a.pitch, a.w, a.h};

        uint32_t steps = 0, seeds = 0, pruned = 0;
        int n = 0, par = 0;
        for (uint32_t row0 = 0; row0 < a.h && n < SMH_LSD_MAX_LINES; row0 += rows_per_chunk) {
            // ---- this thread's slice of the chunk: bitmaps of its white, in-image pixels -------
            uint32_t alive[LSD_VPT];  // 16 bits per vector
            uint32_t vrow[LSD_VPT], vx0[LSD_VPT];
            bool any_white = false;
#pragma unroll
            for (int k = 0; k < LSD_VPT; k++) {
                const uint32_t item = (uint32_t)tid * LSD_VPT + k;  // consecutive items => increasing raster order
                alive[k] = 0; vrow[k] = 0; vx0[k] = 0;
                if (item < items) {
                    const uint32_t r = row0 + item / vpr, c = item % vpr;
                    vrow[k] = r; vx0[k] = c * 16u - a.pad;  // x of byte 0 (may wrap below 0 for c = 0)
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
                // strike off what the lines found so far already suppress (lsd.rs:84-89): one pixel
                // per warp step, lane l tests line l (n <= 32), the vote decides
                if (n > 0) {
                    const int lane = tid & 31;
                    uint32_t pend = alive[k];
                    uint32_t bal;
                    while ((bal = __ballot_sync(0xFFFFFFFFu, pend != 0u)) != 0u) {
                        const int src = __ffs(bal) - 1;
                        const uint32_t bits = __shfl_sync(0xFFFFFFFFu, pend, src);
                        const int b = __ffs(bits) - 1;
                        const float fx = (float)(__shfl_sync(0xFFFFFFFFu, vx0[k], src) + (uint32_t)b);
                        const float fy = (float)__shfl_sync(0xFFFFFFFFu, vrow[k], src);
                        const bool sup = lane < n && suppressed_by(s_lines[lane], fx, fy);
                        const bool hit = __any_sync(0xFFFFFFFFu, sup);
                        if (lane == src) { pend &= pend - 1u; if (hit) alive[k] &= ~(1u << b); }
                    }
                }
                any_white |= alive[k] != 0;
            }
            if (!__syncthreads_or(any_white)) continue;

            // ---- seeds of this chunk, in raster order ------------------------------------------
            while (true) {
                uint32_t cand = 0xFFFFFFFFu;  // lowest remaining pixel of this thread (items are in raster order)
#pragma unroll
                for (int k = LSD_VPT - 1; k >= 0; k--)
                    if (alive[k]) cand = vrow[k] * a.w + vx0[k] + (uint32_t)(__ffs(alive[k]) - 1);
                const uint32_t wmin = __reduce_min_sync(0xFFFFFFFFu, cand);
                if ((tid & 31) == 0) s_min[tid >> 5] = wmin;
                __syncthreads();
                uint32_t seed = 0xFFFFFFFFu;
#pragma unroll
                for (int k = 0; k < T / 32; k++) seed = min(seed, s_min[k]);
                if (seed == 0xFFFFFFFFu) break;  // uniform
                if (cand == seed) {
#pragma unroll
                    for (int k = 0; k < LSD_VPT; k++)
                        if (alive[k] && vrow[k] * a.w + vx0[k] + (uint32_t)(__ffs(alive[k]) - 1) == seed) alive[k] &= alive[k] - 1u;
                }
                float px, py;
                get_centre_warp(im, (float)(seed % a.w), (float)(seed / a.w), px, py);  // same value in every warp
                seeds++;
                if (a.prune) {
                    const bool isolated = seed_is_isolated<T>(base, a.pitch, a.pad, a.w, a.h, px, py, a.gap_limit, &s_mask[par]);
                    if (tid == 0) s_mask[par ^ 1] = 0ull;  // last read one seed ago, behind two barriers
                    par ^= 1;
                    if (isolated) { pruned++; continue; }  // uniform
                }
                float best[3];
                if (fast) longest_ray_cta<T, true, STATS>(im, a.rays, px, py, a.max_gap, a.gap_limit, best, s_keys, s_ends, steps);
                else longest_ray_cta<T, false, STATS>(im, a.rays, px, py, a.max_gap, a.gap_limit, best, s_keys, s_ends, steps);
                if (best[0] > SMH_LSD_MIN_LEN_SQ) {  // uniform
                    float c1x, c1y;
                    get_centre_warp(im, best[1], best[2], c1x, c1y);
                    const float4 line = make_float4(px, py, c1x, c1y);
                    if (tid == 0) {
                        s_lines[n] = line;  // read by later chunks only, behind the barriers in between
                        out_lines[4 * n + 0] = px; out_lines[4 * n + 1] = py;
                        out_lines[4 * n + 2] = c1x; out_lines[4 * n + 3] = c1y;
                    }
                    n++;
                    if (n == SMH_LSD_MAX_LINES) break;  // `break 'row`
                    // the new line suppresses pixels for good: strike them off now
#pragma unroll
                    for (int k = 0; k < LSD_VPT; k++)
                        for (uint32_t bits = alive[k]; bits; bits &= bits - 1u) {
                            const int b = __ffs(bits) - 1;
                            if (suppressed_by(line, (float)(vx0[k] + (uint32_t)b), (float)vrow[k])) alive[k] &= ~(1u << b);
                        }
                }
            }
            __syncthreads();
        }
        if (tid == 0) a.n_lines[f] = (uint32_t)n;
        if (STATS) {
            steps = __reduce_add_sync(0xFFFFFFFFu, steps);
            if ((tid & 31) == 0) atomicAdd(&a.stats[3 * f + 1], (unsigned long long)steps);
            if (tid == 0) { a.stats[3 * f] = seeds; a.stats[3 * f + 2] = pruned; }
        }
    }
}

// single find_longest_line call (Vision::find_longest_line); out = {p0x,p0y,p1x,p1y,len}
__global__ void __launch_bounds__(LSD_THREADS) k_longest(const uint8_t* lsd, uint32_t pitch, uint32_t pad, uint32_t w,
                                                         uint32_t h, const float2* rays, float px, float py,
                                                         float max_gap, float* out) {
    __shared__ unsigned long long s_keys[LSD_THREADS / 32];
    __shared__ float2 s_ends[LSD_THREADS / 32];
    Img im{lsd + pad, pitch, w, h};
    uint32_t steps = 0;
    float best[3];
    longest_ray_cta<LSD_THREADS, false, false>(im, rays, px, py, max_gap, 0, best, s_keys, s_ends, steps);
    if (threadIdx.x == 0) { out[0] = px; out[1] = py; out[2] = best[1]; out[3] = best[2]; out[4] = best[0]; }
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
    // the exact code paths k_pixels uses: which 0 = vector fast-reject + table bit,
    // 1 = luma through I2F.U8 (plain vectors), 6 = luma through PRMT-biased words (quadrant vectors)
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t word = c | 0xFF000000u;
    if (which == 0) out[c] = (uint8_t)(bright_bytes(word) ? marker_nibble(make_uint4(word, 0u, 0u, 0u), lut) & 1u : 0u);
    else if (which == 1) out[c] = (uint8_t)(luma_biased_i2f(word) & 0xFFu);
    else out[c] = (uint8_t)luma8_bgra(word);
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
