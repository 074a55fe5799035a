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
#include "smh_lsd_filter.h"
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
    // 1-bit tile-major copy of the dilated mask (the line search's input, layout: see BitsGeom), written
    // alongside lsd_image when non-null: words per frame, tiles per tile-row
    uint32_t* bits; uint64_t bits_stride; uint32_t bits_tx;
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

// Marker fast-reject for a whole vector: a marker colour has max(r,g,b) >= 178, i.e. some colour
// byte with bit 7 set and low 7 bits >= 50; SIMD-in-register test over the three colour bytes.
__device__ __forceinline__ uint32_t bright_bytes(uint32_t w) {  // nonzero iff some of b,g,r is >= 178
    return ((w & 0x007F7F7Fu) + 0x004E4E4Eu) & w & 0x00808080u;
}

// rare path shared by both vector kinds: exact marker bits of the 4 pixels.  Entered when the OR of
// the four words has a colour byte >= 178 (necessary, not sufficient), so the exact per-word test
// comes first.
__device__ __noinline__ uint32_t marker_nibble(uint4 v, const uint32_t* __restrict__ lut) {
    if (!(bright_bytes(v.x) | bright_bytes(v.y) | bright_bytes(v.z) | bright_bytes(v.w))) return 0u;
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
// Vector-level fast rejects work on the OR of the four BGRA words (2 instructions, shared by both
// tests): a byte of the OR is >= any of the four bytes it combines and keeps their bit 7, so
//   some colour byte >= 178  =>  bright_bytes(or4) != 0          (marker colours: max(r,g,b) >= 178)
//   some pixel with b,g,r all >= 130  =>  bit 7 of all three colour bytes of or4   (OCR classes)
__device__ __forceinline__ uint32_t or4(const uint4& v) { return v.x | v.y | v.z | v.w; }

template <bool GRAY>
__device__ __forceinline__ void vec_plain(const uint4& v, const uint32_t* __restrict__ lut, uint32_t alpha, PxOut& o) {
#if defined(PX_EXP) && PX_EXP >= 1
    if (false) {
#else
    if (GRAY) {  // (L,L,L,255) — vision-cpu/src/lib.rs:150-157
#endif
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
#if !defined(PX_EXP) || PX_EXP < 2
    if (bright_bytes(or4(v))) o.m4 = marker_nibble(v, lut);  // rare
#endif
}

// marker flags only (halo rows: nothing is stored)
__device__ __forceinline__ uint32_t vec_marker_only(const uint4& v, const uint32_t* __restrict__ lut) {
    if (bright_bytes(or4(v))) return marker_nibble(v, lut);
    return 0u;
}

// ---- vector with pixels inside the OCR quadrant (flags: marker, core, edge; the caller masks
//      out what a pixel may not carry: outside the quadrant, above the search clip) ------------
template <bool GRAY>
__device__ __forceinline__ void vec_quadrant(const uint4& v, bool owned, const uint32_t* __restrict__ lut, uint32_t alpha, PxOut& o) {
    const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
    uint32_t uo[4], lmin = 0xFFFFFFFFu;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const uint32_t lb = owned ? luma_biased_i2f(wd[e]) : BIAS + 1u;
        lmin = min(lmin, lb);
        if (GRAY) uo[e] = __byte_perm(lb, alpha, 0x7000);
        else uo[e] = __byte_perm(wd[e], alpha, 0x7012);
    }
    o.ui = make_uint4(uo[0], uo[1], uo[2], uo[3]);
    o.m4 = 0; o.c4 = 0; o.e4 = 0; o.ocrw = 0xFFFFFFFFu; o.scw = 0xFFFFFFFFu;
    const uint32_t any = or4(v);
    if (bright_bytes(any)) o.m4 = marker_nibble(v, lut);  // rare
    // OCR classes — vision-cpu/src/lib.rs:39-53,188-193; mono = sum |ci-cj| = 4*(max-min).  Both
    // classes need every channel >= SMH_OCR_BRIGHT_EDGE (130 > 128): bit 7 of b, g and r.
    static_assert(SMH_OCR_BRIGHT_EDGE >= 128u && SMH_OCR_BRIGHT >= SMH_OCR_BRIGHT_EDGE, "fast reject assumes min channel >= 128");
    if ((any & 0x00808080u) == 0x00808080u) {  // rare: possibly a bright near-grey pixel
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
#define PX_DEPTH 2       // rows of 128-bit copies in flight per lane (B200, 32-row bands: 1 -> 3.07 ms, 2 -> 2.91, 3 -> 3.05 per 1024 frames)
#endif
#ifndef PX_SLOTS_N
#define PX_SLOTS_N 8
#endif
constexpr int PX_SLOTS = PX_SLOTS_N;  // ring slots per warp (power of two > PX_DEPTH), 512 B each
constexpr int PX_SMEM_BYTES = PX_WARPS * PX_SLOTS * 512;

template <bool GRAY, bool BITS = false>
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
    // rows of halo: 3 where the strip and band meet the OCR quadrant (7x7 search), else 1 (3x3 cross)
    const int halo = (brq_strip && r1 > bqh && r0 < 2 * bqh) ? SMH_OCR_RADIUS : 1;
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
#if defined(PX_EXP) && PX_EXP >= 3  // memory pattern only (timing experiment; 4: reads only, 5: writes only)
#if PX_EXP == 8  // lsd plane only, 32 lanes x 4 B = one aligned 128-byte line per warp-row
        if (jn >= r0 && jn < r1 && strip * 128u + 128u <= a.lsd_pitch)
            stg_u32(a.lsd + (uint64_t)f * a.lsd_stride + (int64_t)jn * (int64_t)a.lsd_pitch + strip * 128u + lane * 4u, cur.x);
#endif
#if PX_EXP == 4  // reads only: keep the value alive without storing
        if (cur.x == 0x12345678u && cur.y == 0x9ABCDEF0u) stg_u32(lsd_p + 3 * a.lsd_pitch, cur.z);
        else if (false) {
#else
        if (active && jn >= r0 && jn < r1 && inner) {
#endif
            if (PX_EXP != 7 && PX_EXP < 8) stg_v4(ui_p, cur);
            if (PX_EXP != 6 && PX_EXP != 8) stg_u32(lsd_p + 3 * a.lsd_pitch, cur.x);
            const int jq = jn - bqh;
            if (PX_EXP != 6 && PX_EXP < 8 && (uint32_t)jq < (uint32_t)bqh && col_brq) {
                stg_u32(ocr_f + ((uint32_t)jq * a.q_pitch + qpos), cur.y);
                stg_u32(sc_f + ((uint32_t)jq * a.q_pitch + qpos), cur.z);
            }
        }
        ui_p += a.ui_pitch; lsd_p += a.lsd_pitch;
        return;
#endif
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
        if (BITS) {
            // the same nibble into the bitmap: X = 4q + e, so a lane's nibble is half a byte of word q >> 3 of tile
            // row jc >> 5.  Inner lanes come in pairs (odd lane: even q), the odd one stores the pair's byte.
            const uint32_t hi = __shfl_down_sync(0xFFFFFFFFu, d4, 1);
            if (inner && active && (lane & 1))
                reinterpret_cast<uint8_t*>(a.bits + (uint64_t)f * a.bits_stride)[
                    (((((uint32_t)jc >> 5) * a.bits_tx + (qo >> 3)) << 5) + ((uint32_t)jc & 31u)) * 4u + ((qo & 7u) >> 1)] = (uint8_t)(d4 | (hi << 4));
        }

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
#if !defined(PX_EXP) || PX_EXP < 5
            if (active && j >= j_lo && j < j_hi) cp_async16(ring0 + ((slot + PX_DEPTH) & (PX_SLOTS - 1)) * 512u, fetch_p);
#endif
            cp_async_commit();
            fetch_p += a.pitch;
        }
        cp_async_wait<PX_DEPTH>();  // the group of row jn has landed
#if defined(PX_EXP) && PX_EXP >= 5
        const uint4 cur = make_uint4(jn, slot, lane, 0);
#else
        const uint4 cur = lds_v4(ring0 + slot * 512u);
#endif
        slot = (slot + 1) & (PX_SLOTS - 1);
        row(cur, jn);
    }
}

// ------------------------------------------------------------------------------------------
// line search
// ------------------------------------------------------------------------------------------
// The line search reads the dilated marker mask as a 1-bit-per-pixel, TILE-MAJOR bitmap: tiles of
// 32 x 32 pixels, 128 bytes each (word r of a tile = row r, bit b = pixel column 32*tx + b), tiles
// of a tile-row contiguous.  Columns are in padded coordinates X = x + pad (the byte mask's
// 16-byte vector grid).  A ray then touches a new 128-byte line every ~32 steps whatever its
// direction, the whole frame is 8x smaller than the L8 mask, and a tile-row is one contiguous
// block for the raster seed scan.
struct BitsGeom {
    uint32_t tx, ty;      // tiles per tile-row, tile-rows
    uint32_t pad, w, h;   // mask geometry (pixels); bit X = x + pad
};
struct Bits {
    const uint32_t* __restrict__ p;
    uint32_t tx, pad, w, h;
    __device__ __forceinline__ uint32_t word(uint32_t X32, uint32_t y) const { return __ldg(p + ((((y >> 5) * tx + X32) << 5) + (y & 31u))); }
    __device__ __forceinline__ bool white(uint32_t x, uint32_t y) const {
        const uint32_t X = x + pad;
        return (word(X >> 5, y) >> (X & 31u)) & 1u;
    }
};

// L8 mask -> tile-major bitmap.  Warp = one tile, lane = row of the tile: 32 bytes in, one word out
// (coalesced 128-byte stores).  grid = (ceil(tx / 8), ty, n_frames), block = 256.
__global__ void __launch_bounds__(256) k_pack_bits(const uint8_t* __restrict__ lsd, uint64_t stride, uint32_t pitch, BitsGeom g,
                                                   uint32_t* __restrict__ bits, uint64_t bits_stride) {
    const uint32_t tile = blockIdx.x * 8u + (threadIdx.x >> 5), r = threadIdx.x & 31u, tyi = blockIdx.y, f = blockIdx.z;
    if (tile >= g.tx) return;
    const uint32_t y = tyi * 32u + r;
    uint32_t out = 0;
    if (y < g.h) {
        const uint4* src = reinterpret_cast<const uint4*>(lsd + (uint64_t)f * stride + (uint64_t)y * pitch + tile * 32u);
        const uint4 v0 = ldg_nc_v4(src), v1 = ldg_nc_v4(src + 1);
        const uint32_t wd[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int q = 0; q < 8; q++)  // bytes equal to 255 -> 4 bits (the multiply gathers bits 0,8,16,24 into 24..27)
            out |= (((__vcmpeq4(wd[q], 0xFFFFFFFFu) & 0x01010101u) * 0x01020408u) >> 24) << (4 * q);
    }
    bits[(uint64_t)f * bits_stride + (((uint64_t)tyi * g.tx + tile) << 5) + r] = out;
}

struct LsdArgs {
    const uint32_t* bits; uint64_t bits_stride;  // tile-major bitmaps, words per frame
    BitsGeom g;
    const float2* rays;  // 3600 (cos, sin) pairs built on the host with libm (D11)
    const float2* ckpt; uint32_t n_ckpt;  // accumulated ray offsets at every LSD_CKPT-th step, [3600][n_ckpt]
    const uint32_t* span_tab;             // ray span per pixel offset (k_build_span_table)
    float max_gap;
    int gap_limit;        // ceil(max_gap): `gap.0 >= max_gap` on the integer-valued counter
    int prune;            // 1: exact ray filter + dead clusters; 0: march all 3600 rays of every seed
    uint32_t n_frames;
    uint32_t* counter;    // persistent-CTA frame dispenser (zeroed before the launch)
    const uint8_t* open;  // nullptr: every frame open
    float* lines;         // [n][32][4]
    uint32_t* n_lines;    // [n]
    unsigned long long* stats;  // [n][3] or nullptr
};

struct Img {
    const uint8_t* p; uint32_t pitch, w, h;
    __device__ __forceinline__ uint32_t at(uint32_t x, uint32_t y) const { return __ldg(p + (uint64_t)y * pitch + x); }
};

// find_line_in_image — vision-cpu/src/lib.rs:388-432, literally (any max_gap, any seed).  Used by the
// single-call trait method and when ceil(max_gap) < 1.  Returns the squared length, writes the end point.
template <typename IMG>
__device__ __forceinline__ float march(const IMG& im, float px, float py, float max_gap, float dx, float dy,
                                       float& ex, float& ey, uint32_t& steps) {
    float x = px, y = py, xo = 0.0f, yo = 0.0f;
    float gap_n = 0.0f, gap_x = 0.0f, gap_y = 0.0f;
    const float wf = (float)im.w, hf = (float)im.h;
    while (x >= 0.0f && y >= 0.0f && x < wf && y < hf) {
        steps++;
        if (im.white(__float2uint_rz(x), __float2uint_rz(y))) {
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
    if (xi < im.w && yi < im.h && !im.white(xi, yi)) {
        ex = __fsub_rn(x, dx); ey = __fsub_rn(y, dy);
    } else {
        ex = px; ey = py;
    }
    const float ddx = __fsub_rn(px, ex), ddy = __fsub_rn(py, ey);
    return __fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy));  // distance_sqr, util/src/geometry.rs:62-68
}
struct ImgL8 {  // byte-mask view with the accessor the generic march wants
    Img im; uint32_t w, h;
    __device__ __forceinline__ bool white(uint32_t x, uint32_t y) const { return im.at(x, y) == 255u; }
};

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

// Columns of row y that line l suppresses FOR SURE / POSSIBLY: exact distance to the infinite line
// < 6.99 (squared < 48.9) resp. < 7.15 (squared < 51.1), while the float formula above errs by
// < 0.02 around its threshold of 50.  Pixels outside the outer interval are surely kept, pixels
// inside the inner one surely struck off; only the thin fringe in between (and every pixel for a
// degenerate line) takes the exact test.  Integer pixel columns clamped to [0, w-1]; an interval
// with lo > hi is empty.
struct Band { int ilo, ihi, olo, ohi; };
__device__ __forceinline__ Band line_band(const float4& l, float y, int w) {
    Band b{0, -1, 0, -1};
    const float dx = l.z - l.x, dy = l.w - l.y;
    const float len = sqrtf(dx * dx + dy * dy);
    if (!(len > 1.0f)) { b.olo = 0; b.ohi = w - 1; return b; }  // degenerate: test everything
    const float ci = 6.99f * len, co = 7.15f * len, a = (y - l.y) * dx;
    if (fabsf(dy) < 1.0e-4f * len) {  // |(x - p0x) * dy| < slack over the whole row
        const float slack = 1.0e-4f * (float)w * len;
        if (fabsf(a) < ci - slack) { b.ilo = 0; b.ihi = w - 1; }
        if (fabsf(a) < co + slack) { b.olo = 0; b.ohi = w - 1; }
        return b;
    }
    const float inv = 1.0f / dy;
    const float i1 = (a - ci) * inv, i2 = (a + ci) * inv, o1 = (a - co) * inv, o2 = (a + co) * inv;
    const float wf = (float)(w - 1);
    const float ilo = fmaxf(l.x + fminf(i1, i2) + 0.01f, 0.0f), ihi = fminf(l.x + fmaxf(i1, i2) - 0.01f, wf);
    const float olo = fmaxf(l.x + fminf(o1, o2) - 0.01f, 0.0f), ohi = fminf(l.x + fmaxf(o1, o2) + 0.01f, wf);
    if (ilo <= ihi) { b.ilo = (int)ceilf(ilo); b.ihi = (int)floorf(ihi); }
    if (olo <= ohi) { b.olo = (int)floorf(olo); b.ohi = (int)ceilf(ohi); }
    return b;
}

// Same automaton for seeds that lie inside the image (all coordinates within (-2^22, 2^22)) and
// gap_limit = ceil(max_gap) >= 1: floor() by a round-down add of 1.5*2^23 (full-rate FADD instead
// of F2I), both bounds in one unsigned compare each, integer gap counter.  With gap_limit >= 1 a
// ray can only stop on a gap that has a recorded start, so the reference's reset of the saved
// position to (0,0) on white pixels is unobservable and dropped.  Bit-identical end points to
// march() — floor(x) in [0,w) <=> 0 <= x < w, and floor == trunc for x >= 0.
// The samples are generated and fetched LSD_BATCH at a time: the positions of a ray do not depend
// on the pixels it reads, so the loads of a batch are independent and their latency overlaps; the
// gap automaton then consumes the batch in order and ignores what lies beyond the stopping sample.
#ifndef LSD_BATCH
#define LSD_BATCH 8
#endif
template <bool STATS>
__device__ __forceinline__ float march_fast(const Bits& im, float px, float py, int gap_limit, float dx, float dy,
                                            float& ex, float& ey, uint32_t& steps) {
    float x = px, y = py, xo = 0.0f, yo = 0.0f, gap_x = px, gap_y = py;
    int gap = 0;
    int state = 0;  // 0 marching, 1 stopped on a gap, 2 left the image
    while (state == 0) {
        float xs[LSD_BATCH], ys[LSD_BATCH];
        uint32_t wh = 0, inb = 0;  // bit u: sample u is white / inside the image
#pragma unroll
        for (int u = 0; u < LSD_BATCH; u++) {
            xs[u] = x; ys[u] = y;
            const uint32_t ix = (uint32_t)(__float_as_int(__fadd_rd(x, 12582912.0f)) - 0x4B400000);
            const uint32_t iy = (uint32_t)(__float_as_int(__fadd_rd(y, 12582912.0f)) - 0x4B400000);
            if (ix < im.w && iy < im.h) {
                inb |= 1u << u;
                const uint32_t X = ix + im.pad;
                wh |= ((im.word(X >> 5, iy) >> (X & 31u)) & 1u) << u;
            }
            xo = __fadd_rn(xo, dx); yo = __fadd_rn(yo, dy);
            x = __fadd_rn(xo, px); y = __fadd_rn(yo, py);
        }
#pragma unroll
        for (int u = 0; u < LSD_BATCH; u++) {
            if (state == 0) {
                if (!((inb >> u) & 1u)) { state = 2; gap_x = xs[u]; gap_y = ys[u]; }
                else {
                    if (STATS) steps++;
                    const bool black = !((wh >> u) & 1u);
                    if (black && gap >= gap_limit) state = 1;
                    else {
                        if (black && gap == 0) { gap_x = xs[u]; gap_y = ys[u]; }
                        gap = black ? gap + 1 : 0;
                    }
                }
            }
        }
    }
    // stopped on a gap: back to where the gap began; left the image: the first outside sample
    x = gap_x; y = gap_y;
    const uint32_t xi = __float2uint_rz(x), yi = __float2uint_rz(y);  // `as u32`: saturating (D7)
    if (xi < im.w && yi < im.h && !im.white(xi, yi)) {
        ex = __fsub_rn(x, dx); ey = __fsub_rn(y, dy);
    } else {
        ex = px; ey = py;
    }
    const float ddx = __fsub_rn(px, ex), ddy = __fsub_rn(py, ey);
    return __fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy));
}

// ---- exact ray filter (smh_lsd_filter.h): per seed, scan the pixel window around it, list the
//      white pixels that lie in one of the step-window rings, mark the rays each can serve, and
//      march only the rays marked in every window ---------------------------------------------
#ifndef LSD_THREADS_N
#define LSD_THREADS_N 256
#endif
#ifndef LSD_MIN_CTAS
#define LSD_MIN_CTAS 3
#endif
constexpr int LSD_THREADS = LSD_THREADS_N;
constexpr int LSD_DEAD = 16;      // remembered dead clusters per frame

// Optional per-phase cycle accounting of k_lsd (build with -DLSD_PROFILE; tools/lsd_profile.py reads it)
__device__ unsigned long long g_lsd_prof[16];
#ifdef LSD_PROFILE
#define LSD_PROF_DECL long long prof_t = clock64(); unsigned long long prof_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define LSD_CNT(k, v) do { prof_acc[k] += (unsigned long long)(v); } while (0)
#define LSD_PROF(k) do { const long long now__ = clock64(); prof_acc[k] += (unsigned long long)(now__ - prof_t); prof_t = now__; } while (0)
#define LSD_PROF_FLUSH do { if (threadIdx.x == 0) for (int k__ = 0; k__ < 16; k__++) atomicAdd(&g_lsd_prof[k__], prof_acc[k__]); } while (0)
#else
#define LSD_PROF_DECL
#define LSD_CNT(k, v) do { } while (0)
#define LSD_PROF(k) do { } while (0)
#define LSD_PROF_FLUSH do { } while (0)
#endif

struct LsdShared {
    float4 lines[SMH_LSD_MAX_LINES];
    float4 dead[LSD_DEAD];  // (x, y, rho^2, raw^2): centred seeds closer than rho to (x, y) cannot yield a line
    unsigned long long keys[LSD_THREADS / 32];
    float2 ends[LSD_THREADS / 32];
    uint32_t smin[2][LSD_THREADS / 32];  // double-buffered: a seed may be dismissed without another barrier
    uint32_t wsum[LSD_THREADS / 32];
    uint32_t bits[SMH_LSD_FILTER_MAX_WINDOWS][SMH_LSD_BITMAP_WORDS + 3];
    uint32_t pre[SMH_LSD_BITMAP_WORDS + 3];  // rays before each word of the combined bitmap
    uint32_t grp_cnt[LSD_THREADS / 32];  // entries per warp segment of grp[]
    uint32_t near2, far2, n_dead, frame, surv_n;
    uint32_t surv[LSD_THREADS];    // rays still alive after the in-window march: k | gap << 16
    union {
        uint32_t grp[1024];            // LSD_GRP_CAP non-empty pixel quarters of the window around the seed
        uint16_t rays[SMH_LSD_RAYS];   // compacted indices of the rays to march
    };
};

constexpr int LSD_SPAN_R = 99;                   // largest window radius (smh_lsd_filter_init: gap_limit <= 48)
constexpr int LSD_SPAN_N = 2 * LSD_SPAN_R + 1;

// The neighbourhood of a seed, copied once from the bitmap into shared memory: rows [y0, y1], 8 words
// per row starting at bitmap word wlo.  get_centre of the seed, the ring scan and the first
// LSD_WIN_STEPS samples of every ray read it instead of global memory.
constexpr int LSD_WIN_EXTRA = 3;                                   // get_centre moves the seed by < 3 px
constexpr int LSD_WIN_ROWS = 2 * (LSD_SPAN_R + LSD_WIN_EXTRA) + 1;  // 205
struct WinView {
    const uint32_t* p;
    int x0, y0, x1, y1, wlo;
    uint32_t pad, w, h;  // image geometry (w, h: for get_centre's clamping)
    __device__ __forceinline__ bool white(uint32_t x, uint32_t y) const {
        const uint32_t X = x + pad;
        return (p[((int)y - y0) * 8 + (int)(X >> 5) - wlo] >> (X & 31u)) & 1u;
    }
};
template <int T>
__device__ __forceinline__ WinView lsd_load_window(const Bits& im, uint32_t* win, int sx, int sy, int radius) {
    WinView v;
    v.p = win; v.pad = im.pad; v.w = im.w; v.h = im.h;
    v.x0 = max(sx - radius, 0); v.x1 = min(sx + radius, (int)im.w - 1);
    v.y0 = max(sy - radius, 0); v.y1 = min(sy + radius, (int)im.h - 1);
    v.wlo = (v.x0 + (int)im.pad) >> 5;
    const int nwords = ((v.x1 + (int)im.pad) >> 5) - v.wlo + 1;  // <= 8 for radius <= 102
    const int total = (v.y1 - v.y0 + 1) * 8;
    for (int i = threadIdx.x; i < total; i += T) {  // independent loads: one memory round trip
        const int r = i >> 3, wi = i & 7;
        win[i] = wi < nwords ? im.word((uint32_t)(v.wlo + wi), (uint32_t)(v.y0 + r)) : 0u;
    }
    return v;
}

// ---- one ray per WARP -------------------------------------------------------------------------
// The accumulated offsets xo_n = fl(xo_{n-1} + dx) of a ray do not depend on the seed, so they are
// tabulated once per handle for every ray direction at every LSD_CKPT-th step (k_build_ray_ckpt).
// Lane i of a warp then rebuilds step 32c + i of block c with at most LSD_CKPT - 1 adds — the same
// float sequence as the serial loop — and the 32 samples of a block are fetched and tested in
// parallel; the gap automaton runs on the two ballot masks with bit tricks.  A ray costs ~2
// instructions per step instead of ~30, and its latency is one table + one bitmap load per 32 steps.
constexpr int LSD_CKPT = 4;
__global__ void k_build_ray_ckpt(const float2* __restrict__ rays, uint32_t n_ckpt, float2* __restrict__ ckpt) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= SMH_LSD_RAYS) return;
    const float2 d = rays[k];
    float xo = 0.0f, yo = 0.0f;
    for (uint32_t c = 0; c < n_ckpt; c++) {
        ckpt[(uint64_t)k * n_ckpt + c] = make_float2(xo, yo);
#pragma unroll
        for (int j = 0; j < LSD_CKPT; j++) { xo = __fadd_rn(xo, d.x); yo = __fadd_rn(yo, d.y); }
    }
}

// AND of k consecutive bits ending at each position (k >= 1)
__device__ __forceinline__ unsigned long long run_and(unsigned long long r, int k) {
    int len = 1;
    while (2 * len <= k) { r &= r << len; len *= 2; }
    if (len < k) r &= r << (k - len);
    return r;
}

// One block (32 consecutive samples) of one ray, lane i = sample 32c + i.
struct RayBlock { float x, y; bool ok; uint32_t word; uint32_t bit; };
__device__ __forceinline__ RayBlock ray_block_fetch(const Bits& im, float2 ck, int sub, float dx, float dy, float px, float py) {
    RayBlock b;
    float xo = ck.x, yo = ck.y;
#pragma unroll
    for (int j = 1; j < LSD_CKPT; j++)
        if (sub >= j) { xo = __fadd_rn(xo, dx); yo = __fadd_rn(yo, dy); }
    b.x = __fadd_rn(xo, px); b.y = __fadd_rn(yo, py);
    const uint32_t ix = (uint32_t)(__float_as_int(__fadd_rd(b.x, 12582912.0f)) - 0x4B400000);
    const uint32_t iy = (uint32_t)(__float_as_int(__fadd_rd(b.y, 12582912.0f)) - 0x4B400000);
    b.ok = ix < im.w && iy < im.h;
    const uint32_t X = ix + im.pad;
    b.bit = X & 31u;
    b.word = 0u;
    if (b.ok) b.word = im.word(X >> 5, iy);
    return b;
}
// Runs the gap automaton over the block.  Returns true when the ray ends in it: (gx, gy) is then the
// start of the fatal gap or the first sample outside the image.  Uniform across the warp.
template <bool STATS>
__device__ __forceinline__ bool ray_block_decide(const RayBlock& b, float xprev, float yprev, uint32_t& gap_in, int gap_limit,
                                                 float& gx, float& gy, uint32_t& steps) {
    const uint32_t inb = __ballot_sync(0xFFFFFFFFu, b.ok), wh = __ballot_sync(0xFFFFFFFFu, (b.word >> b.bit) & 1u);
    if ((inb & wh) == 0xFFFFFFFFu) {  // all inside, all white (the body of a marker line)
        if (STATS) steps += 32u;
        gap_in = 0u;
        return false;
    }
    const int first_out = __ffs(~inb) - 1;  // -1: the whole block is inside
    const uint32_t valid = first_out < 0 ? 0xFFFFFFFFu : ((1u << first_out) - 1u);
    const uint32_t black = ~wh & valid;
    // bit 32+i of r: sample i is the (gap_limit+1)-th consecutive non-white one
    const unsigned long long carry = gap_in ? ((gap_in >= 32u ? 0xFFFFFFFFull : ((1ull << gap_in) - 1ull)) << (32u - min(gap_in, 32u))) : 0ull;
    const uint32_t stop = (uint32_t)(run_and(((unsigned long long)black << 32) | carry, gap_limit + 1) >> 32);
    if (stop) {
        const int i = __ffs(stop) - 1, g = i - gap_limit;  // the gap began at sample g (previous block when negative)
        if (STATS) steps += (uint32_t)(i + 1);
        const float cx = __shfl_sync(0xFFFFFFFFu, b.x, g & 31), cy = __shfl_sync(0xFFFFFFFFu, b.y, g & 31);
        const float qx = __shfl_sync(0xFFFFFFFFu, xprev, g & 31), qy = __shfl_sync(0xFFFFFFFFu, yprev, g & 31);
        gx = g >= 0 ? cx : qx; gy = g >= 0 ? cy : qy;
        return true;
    }
    if (first_out >= 0) {
        if (STATS) steps += (uint32_t)first_out;
        gx = __shfl_sync(0xFFFFFFFFu, b.x, first_out); gy = __shfl_sync(0xFFFFFFFFu, b.y, first_out);
        return true;
    }
    if (STATS) steps += 32u;
    gap_in = black == 0xFFFFFFFFu ? gap_in + 32u : (uint32_t)__clz((int)~black);  // trailing non-white run
    return false;
}
// end-point rule (vision-cpu/src/lib.rs:426-429, D7) and squared length
__device__ __forceinline__ float ray_finish(const Bits& im, float gx, float gy, float px, float py, float dx, float dy, float& ex, float& ey) {
    const uint32_t xi = __float2uint_rz(gx), yi = __float2uint_rz(gy);  // `as u32`: saturating
    if (xi < im.w && yi < im.h && !im.white(xi, yi)) {
        ex = __fsub_rn(gx, dx); ey = __fsub_rn(gy, dy);
    } else {
        ex = px; ey = py;
    }
    const float ddx = __fsub_rn(px, ex), ddy = __fsub_rn(py, ey);
    return __fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy));
}

// First `limit` samples (a multiple of LSD_BATCH) of one ray per LANE, read from the seed's window in
// shared memory (all of them lie inside it).  Returns 0: still marching (gap = current run of
// non-white samples), 1: stopped on a gap that began at (gx, gy), 2: (gx, gy) is the first sample
// outside the image.  Same arithmetic as march_fast.
template <bool STATS, typename WIN>
__device__ __forceinline__ int march_window(const WIN& wv, int limit, float px, float py, int gap_limit, float dx, float dy,
                                            float& gx, float& gy, int& gap, uint32_t& steps) {
    float x = px, y = py, xo = 0.0f, yo = 0.0f;
    gx = px; gy = py; gap = 0;
    int state = 0;
    for (int n0 = 0; n0 < limit && state == 0; n0 += LSD_BATCH) {
        float xs[LSD_BATCH], ys[LSD_BATCH];
        uint32_t wh = 0, inb = 0;  // bit u: sample u is white / inside the image
#pragma unroll
        for (int u = 0; u < LSD_BATCH; u++) {
            xs[u] = x; ys[u] = y;
            const uint32_t ix = (uint32_t)(__float_as_int(__fadd_rd(x, 12582912.0f)) - 0x4B400000);
            const uint32_t iy = (uint32_t)(__float_as_int(__fadd_rd(y, 12582912.0f)) - 0x4B400000);
            if (ix < wv.w && iy < wv.h) {
                inb |= 1u << u;
                wh |= (wv.white(ix, iy) ? 1u : 0u) << u;
            }
            xo = __fadd_rn(xo, dx); yo = __fadd_rn(yo, dy);
            x = __fadd_rn(xo, px); y = __fadd_rn(yo, py);
        }
#pragma unroll
        for (int u = 0; u < LSD_BATCH; u++) {
            if (state == 0) {
                if (!((inb >> u) & 1u)) { state = 2; gx = xs[u]; gy = ys[u]; }
                else {
                    if (STATS) steps++;
                    const bool black = !((wh >> u) & 1u);
                    if (black && gap >= gap_limit) state = 1;
                    else {
                        if (black && gap == 0) { gx = xs[u]; gy = ys[u]; }
                        gap = black ? gap + 1 : 0;
                    }
                }
            }
        }
    }
    return state;
}

// The listed rays of one seed.  Phase 1, one ray per lane: the first `win_steps` samples from the
// window in shared memory — most rays end there.  Phase 2, one ray per warp: the survivors continue
// block by block on the global bitmap (ray_block_*), the next checkpoint prefetched.  Same
// last-index argmax as longest_ray_cta.  Preconditions: 1 <= gap_limit <= 32, seed inside the
// image, the offsets table covers every step a ray can take inside the image plus two blocks,
// win_steps a multiple of 32 whose samples all fall inside the window.
template <int T, bool STATS>
__device__ __forceinline__ void longest_ray_hybrid(const Bits& im, const WinView& wv, int win_steps, const float2* __restrict__ rays,
                                                   const float2* __restrict__ ckpt, uint32_t n_ckpt, const uint16_t* list, int n_rays,
                                                   float px, float py, int gap_limit, float (&best)[3], LsdShared& sh, uint32_t& steps) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, sub = lane & (LSD_CKPT - 1);
    unsigned long long best_key = 0ull;
    float bex = px, bey = py;
    uint32_t st = 0;
    for (int base = 0; base < n_rays; base += T) {
        // ---- phase 1 ---------------------------------------------------------------------------
        const int i = base + tid;
        bool alive = false;
        int k = 0, gap = 0;
        if (i < n_rays) {
            k = (int)list[i];
            const float2 d = __ldg(&rays[k]);
            float gx, gy;
            const int state = win_steps > 0 ? march_window<STATS>(wv, win_steps, px, py, gap_limit, d.x, d.y, gx, gy, gap, steps) : 0;
            if (state == 0) alive = true;
            else {
                float ex, ey, len;
                if (state == 1) {  // the gap's first sample is inside the image and not white: lib.rs:426-429 holds
                    ex = __fsub_rn(gx, d.x); ey = __fsub_rn(gy, d.y);
                    const float ddx = __fsub_rn(px, ex), ddy = __fsub_rn(py, ey);
                    len = __fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy));
                } else len = ray_finish(im, gx, gy, px, py, d.x, d.y, ex, ey);
                const unsigned long long key = ((unsigned long long)__float_as_uint(len) << 32) | (uint32_t)k;
                if (key >= best_key) { best_key = key; bex = ex; bey = ey; }
            }
        }
        const uint32_t am = __ballot_sync(0xFFFFFFFFu, alive);
        if (am) {
            uint32_t slot = 0;
            if (lane == 0) slot = atomicAdd(&sh.surv_n, (uint32_t)__popc(am));
            slot = __shfl_sync(0xFFFFFFFFu, slot, 0) + (uint32_t)__popc(am & ((1u << lane) - 1u));
            if (alive) sh.surv[slot] = (uint32_t)k | ((uint32_t)gap << 16);
        }
        __syncthreads();
        // ---- phase 2 ---------------------------------------------------------------------------
        const int n_surv = (int)sh.surv_n;
        const int c0 = win_steps / 32;
        for (int j = warp; j < n_surv; j += T / 32) {
            const uint32_t e = sh.surv[j];
            const int k2 = (int)(e & 0xFFFFu);
            uint32_t gap_in = e >> 16;
            const float2 d = __ldg(&rays[k2]);
            const float2* cp = ckpt + (uint64_t)k2 * n_ckpt + (lane / LSD_CKPT) + c0 * (32 / LSD_CKPT);
            float2 c1 = __ldg(cp);
            float xprev = px, yprev = py;
            if (c0 > 0) {  // this lane's sample of block c0 - 1 (a gap may have begun there)
                const float2 cq = __ldg(cp - 32 / LSD_CKPT);
                float xo = cq.x, yo = cq.y;
#pragma unroll
                for (int q = 1; q < LSD_CKPT; q++)
                    if (sub >= q) { xo = __fadd_rn(xo, d.x); yo = __fadd_rn(yo, d.y); }
                xprev = __fadd_rn(xo, px); yprev = __fadd_rn(yo, py);
            }
            // two blocks in flight: block c+1 is fetched (table value loaded one iteration earlier) before
            // block c is decided, so bitmap and table latencies overlap; what is fetched beyond the
            // block where the ray ends is dropped
            float gx, gy;
            cp += 32 / LSD_CKPT;
            float2 c2 = __ldg(cp);
            RayBlock b = ray_block_fetch(im, c1, sub, d.x, d.y, px, py);
            while (true) {
                cp += 32 / LSD_CKPT;
                const float2 c3 = __ldg(cp);
                const RayBlock bn = ray_block_fetch(im, c2, sub, d.x, d.y, px, py);
                if (ray_block_decide<STATS>(b, xprev, yprev, gap_in, gap_limit, gx, gy, st)) break;
                xprev = b.x; yprev = b.y; b = bn; c2 = c3;
            }
            float ex, ey;
            const float len = ray_finish(im, gx, gy, px, py, d.x, d.y, ex, ey);
            const unsigned long long key = ((unsigned long long)__float_as_uint(len) << 32) | (uint32_t)k2;
            if (key >= best_key) { best_key = key; bex = ex; bey = ey; }  // same in every lane
        }
        __syncthreads();
        if (tid == 0) sh.surv_n = 0u;  // next appended to behind another barrier (next round / next seed)
        if (base + T < n_rays) __syncthreads();
    }
    if (STATS && lane == 0) steps += st;  // phase-2 samples: once per warp (steps are summed over all threads later)
    // warp argmax of (len, k), then across the warps
    unsigned long long wkey = best_key;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, wkey, o);
        wkey = other > wkey ? other : wkey;
    }
    const int src = __ffs(__ballot_sync(0xFFFFFFFFu, best_key == wkey)) - 1;
    bex = __shfl_sync(0xFFFFFFFFu, bex, src); bey = __shfl_sync(0xFFFFFFFFu, bey, src);
    if (lane == 0) { sh.keys[warp] = wkey; sh.ends[warp] = make_float2(bex, bey); }
    __syncthreads();
    int bw = 0;
    unsigned long long bk = sh.keys[0];
#pragma unroll
    for (int k2 = 1; k2 < T / 32; k2++) { const unsigned long long o = sh.keys[k2]; if (o > bk) { bk = o; bw = k2; } }
    best[0] = __uint_as_float((uint32_t)(bk >> 32));
    best[1] = sh.ends[bw].x;
    best[2] = sh.ends[bw].y;
}

// get_centre — vision-common/src/lsd.rs:5-44 (indices clamped where the reference is UB, D8) —
// evaluated by one warp in parallel (every lane returns the result): lane = 8*dir + t tests the
// loop condition of step t of direction dir (left, right, up, down); the walk stops at the first
// failing t, so the stop index is the number of trailing ones of the ballot byte.  Each lane
// rebuilds its v by t successive +-1.0 adds, because the reference's `(v - pt).abs() < 5` sees the
// accumulated rounding: across a binade boundary (e.g. 511.6 -> 512.6) five increments can leave
// |v - pt| just below 5 and the walk takes a sixth step.
template <typename IMG>
__device__ __forceinline__ bool white_clamped(const IMG& im, float fx, float fy) {
    uint32_t x = __float2uint_rz(fx), y = __float2uint_rz(fy);
    x = min(x, im.w - 1u); y = min(y, im.h - 1u);
    return im.white(x, y);
}
template <typename IMG>
__device__ __forceinline__ void get_centre_warp(const IMG& im, float ptx, float pty, float& cx, float& cy) {
    const int lane = threadIdx.x & 31, dir = lane >> 3, t = lane & 7;
    const float pt = dir < 2 ? ptx : pty;
    const float step = (dir & 1) ? 1.0f : -1.0f;
    float v = pt;
    for (int i = 0; i < t; i++) v = __fadd_rn(v, step);
    bool ok = fabsf(__fsub_rn(v, pt)) < SMH_LSD_CENTRE_MAX_DIST;
    if (dir == 0) ok = ok && v > 0.0f && white_clamped(im, v, pty);
    else if (dir == 1) ok = ok && v < (float)(im.w - 1u) && white_clamped(im, v, pty);
    else if (dir == 2) ok = ok && v > 0.0f && white_clamped(im, ptx, v);
    else ok = ok && v < (float)(im.h - 1u) && white_clamped(im, ptx, v);
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, ok);
    // value after k steps = the v of lane 8*dir + k (k <= 7: lane t = 7 always fails the distance test)
    const int kl = __ffs(~(bal & 0xFFu)) - 1, kr = __ffs(~((bal >> 8) & 0xFFu)) - 1;
    const int ku = __ffs(~((bal >> 16) & 0xFFu)) - 1, kd = __ffs(~((bal >> 24) & 0xFFu)) - 1;
    const float left = __shfl_sync(0xFFFFFFFFu, v, kl), right = __shfl_sync(0xFFFFFFFFu, v, 8 + kr);
    const float up = __shfl_sync(0xFFFFFFFFu, v, 16 + ku), down = __shfl_sync(0xFFFFFFFFu, v, 24 + kd);
    cx = __fdiv_rn(__fadd_rn(left, right), 2.0f);
    cy = __fdiv_rn(__fadd_rn(up, down), 2.0f);
}

// Ray spans by pixel offset.  A seed centre is a multiple of 0.5 in each axis (get_centre of an
// integer pixel), so the ray interval smh_lsd_filter_span() assigns to the pixel at integer offset
// (ox, oy) from floor(seed) only depends on the offset and on the two half-pixel flags: it is
// tabulated once per handle ([4][SPAN_N][SPAN_N] x (lo | hi << 16), int16 each) and the per-pixel
// atan2 / sqrt of the marking step become one 4-byte load.
__global__ void k_build_span_table(uint32_t* __restrict__ tab) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 4 * LSD_SPAN_N * LSD_SPAN_N) return;
    const int v = i / (LSD_SPAN_N * LSD_SPAN_N), rem = i - v * LSD_SPAN_N * LSD_SPAN_N;
    const int oy = rem / LSD_SPAN_N - LSD_SPAN_R, ox = rem % LSD_SPAN_N - LSD_SPAN_R;
    const float cx = (float)ox + 0.5f - ((v & 1) ? 0.5f : 0.0f), cy = (float)oy + 0.5f - ((v & 2) ? 0.5f : 0.0f);
    const float d2 = cx * cx + cy * cy;
    int lo = 0, hi = SMH_LSD_RAYS - 1;
    if (d2 >= 1.0f) smh_lsd_filter_span(cx, cy, d2, &lo, &hi);  // lo >= -1803, hi <= 1803 + span
    tab[i] = (uint32_t)(uint16_t)(int16_t)lo | ((uint32_t)(uint16_t)(int16_t)hi << 16);
}

// OR `m` into a shared bitmap word unless it is already there (most marks of a seed repeat: the
// pixels of a line all point the same way, and same-address shared atomics serialise)
__device__ __forceinline__ void or_if_new(uint32_t* w, uint32_t m) {
    if ((*reinterpret_cast<volatile uint32_t*>(w) & m) != m) atomicOr(w, m);
}

// Stage 1: list the non-empty pixel quarters (8 pixels of one word of one row) of the (2R+1)^2 window
// around the centred seed: sh.grp[] = mask | (x offset of the quarter's first pixel + 160) << 8 |
// (row offset + 128) << 17, offsets relative to floor(seed).
constexpr int LSD_GRP_SEG = 1024 / (LSD_THREADS / 32);  // entries per warp segment of sh.grp
template <int T>
__device__ __forceinline__ void lsd_scan_window(const WinView& wv, float px, float py, const smh_lsd_filter& F, LsdShared& sh) {
    const int fx = (int)px, fy = (int)py, R = F.radius;  // px, py >= 0
    const int x0 = max(fx - R, wv.x0), x1 = min(fx + R, wv.x1);
    const int y0 = max(fy - R, wv.y0), y1 = min(fy + R, wv.y1);
    const int X0 = x0 + (int)wv.pad, X1 = x1 + (int)wv.pad;
    const int wlo = X0 >> 5, nwords = (X1 >> 5) - wlo + 1;
    const int lane = threadIdx.x & 31, wi = lane >> 2, q = lane & 3;
    const int Xw = (wlo + wi) << 5;
    uint32_t keep = 0u;  // this lane's 8 pixels of the word, inside [X0, X1]
    if (wi < nwords) {
        keep = 0xFFu << (8 * q);
        if (Xw < X0) keep &= 0xFFFFFFFFu << (X0 - Xw);
        if (Xw + 31 > X1) keep &= 0xFFFFFFFFu >> (Xw + 31 - X1);
    }
    const uint32_t gx = (uint32_t)(Xw + 8 * q - (int)wv.pad - fx + 160);
    const uint32_t* col = wv.p + (wlo + wi - wv.wlo);
    const int warp = threadIdx.x >> 5;
    uint32_t cnt = 0;  // entries of this warp's segment so far (uniform)
    for (int y = y0 + warp; y <= y1; y += T / 32) {  // warp-uniform trip count
        const uint32_t m8 = keep ? (col[(y - wv.y0) * 8] & keep) >> (8 * q) : 0u;
        const uint32_t have = __ballot_sync(0xFFFFFFFFu, m8 != 0u);
        if (have == 0u) continue;  // nothing white in this row of the window (uniform)
        const uint32_t slot = cnt + (uint32_t)__popc(have & ((1u << lane) - 1u));
        if (m8 && slot < (uint32_t)LSD_GRP_SEG) sh.grp[warp * LSD_GRP_SEG + (int)slot] = m8 | (gx << 8) | ((uint32_t)(y - fy + 128) << 17);
        cnt += (uint32_t)__popc(have);
    }
    if (lane == 0) sh.grp_cnt[warp] = cnt;  // > LSD_GRP_SEG: overflow
}

// bits lo..hi (any integers) that fall into word w of a bitmap
__device__ __forceinline__ uint32_t range_mask(int w, int lo, int hi) {
    const int s = max(lo - 32 * w, 0), e = min(hi - 32 * w, 31);
    return s <= e ? ((0xFFFFFFFFu << s) & (0xFFFFFFFFu >> (31 - e))) : 0u;
}

// Stage 2: every warp marks the quarters it listed.  The ray spans of horizontally adjacent pixels
// overlap, so a quarter is marked as a whole: the union of the spans of its first and last white
// pixel, in every ring that the distance range of the pixels between them meets (a superset of the
// per-pixel marks).  Lanes first work out the quarters' parameters in parallel (one each), then the
// warp walks through them with every lane OR-ing the broadcast span into the bitmap words it owns
// IN REGISTERS (word lane + 32 i of each ring) — no shared-memory atomics in the loop; the
// accumulated words are merged into sh.bits at the end.  Also reduces, conservatively, the nearest /
// farthest white distances around the inner disc into sh.near2 / sh.far2 (float bits; positive
// floats order like their bit patterns).
template <int T>
__device__ __forceinline__ void lsd_mark_rays(const uint32_t* __restrict__ span_tab, float px, float py, const smh_lsd_filter& F, LsdShared& sh) {
    bool overflow = false;
#pragma unroll
    for (int k = 0; k < T / 32; k++) overflow |= sh.grp_cnt[k] > (uint32_t)LSD_GRP_SEG;
    if (overflow) {  // too many white pixels around: mark everything (uniform)
        for (int i = threadIdx.x; i < SMH_LSD_FILTER_MAX_WINDOWS * (SMH_LSD_BITMAP_WORDS + 3); i += T) (&sh.bits[0][0])[i] = 0xFFFFFFFFu;
        if (threadIdx.x == 0) sh.far2 = 0u;  // nothing is known about the surroundings: no dead cluster
        return;
    }
    const int fx = (int)px, fy = (int)py, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float hx = px - (float)fx, hy = py - (float)fy;  // 0 or 0.5
    const uint32_t* tab = span_tab + ((hx != 0.0f ? 1 : 0) + (hy != 0.0f ? 2 : 0)) * (LSD_SPAN_N * LSD_SPAN_N) + LSD_SPAN_R * LSD_SPAN_N + LSD_SPAN_R;
    const float a2 = F.r2_lo[0];
    float near2 = 0.0f, far2 = 3.0e38f;
    uint32_t acc[SMH_LSD_FILTER_MAX_WINDOWS][4];
#pragma unroll
    for (int j = 0; j < SMH_LSD_FILTER_MAX_WINDOWS; j++)
#pragma unroll
        for (int i = 0; i < 4; i++) acc[j][i] = 0u;
    const int cnt = (int)sh.grp_cnt[warp];
    for (int base = 0; base < cnt; base += 32) {
        // ---- this lane's quarter: span [lo, hi] with 0 <= lo < 3600 (hi may run past 3599), rings -----
        int lo = 0, hi = -1;
        uint32_t rings = 0;
        if (base + lane < cnt) {
            const uint32_t e = sh.grp[warp * LSD_GRP_SEG + base + lane], m8 = e & 0xFFu;
            const int oxa = (int)((e >> 8) & 0x1FFu) - 160 + (__ffs(m8) - 1), oxb = (int)((e >> 8) & 0x1FFu) - 160 + (31 - __clz((int)m8));
            const int oy = (int)(e >> 17) - 128;
            const float cy = (float)oy + 0.5f - hy, cxa = (float)oxa + 0.5f - hx, cxb = (float)oxb + 0.5f - hx;  // exact
            const float cy2 = cy * cy, da = cxa * cxa + cy2, db = cxb * cxb + cy2;
            const float dmax = fmaxf(da, db), dmin = (cxa < 0.0f && cxb > 0.0f) ? cy2 : fminf(da, db);  // over [cxa, cxb]
            if (dmax < a2) near2 = fmaxf(near2, dmax);
            else {
                if (dmin < a2) near2 = fmaxf(near2, a2);  // straddles the inner disc's rim: no dead cluster from this seed
                far2 = fminf(far2, fmaxf(dmin, a2));
#pragma unroll
                for (int j = 0; j < SMH_LSD_FILTER_MAX_WINDOWS; j++)
                    if (j < F.nw && dmax >= F.r2_lo[j] && dmin <= F.r2_hi[j]) rings |= 1u << j;
                if (rings) {
                    const uint32_t sa = __ldg(tab + oy * LSD_SPAN_N + oxa), sb = __ldg(tab + oy * LSD_SPAN_N + oxb);
                    // direction is monotonic along the row between the two pixels, the half-width is not: take the
                    // range of the centres (sum / 2, +-0.5 from the integer rounding) widened by the larger width
                    const int loa = (int)(int16_t)(sa & 0xFFFFu), hia = (int)(int16_t)(sa >> 16);
                    const int lob = (int)(int16_t)(sb & 0xFFFFu), hib = (int)(int16_t)(sb >> 16);
                    const int wide = max(hia - loa, hib - lob);
                    lo = ((min(loa + hia, lob + hib) - wide) >> 1) - 1; hi = ((max(loa + hia, lob + hib) + wide + 1) >> 1) + 1;
                    if (hi - lo >= SMH_LSD_RAYS - 1) { lo = 0; hi = SMH_LSD_RAYS - 1; }
                    else {
                        if (lo < 0) { lo += SMH_LSD_RAYS; hi += SMH_LSD_RAYS; }
                        if (lo >= SMH_LSD_RAYS) { lo -= SMH_LSD_RAYS; hi -= SMH_LSD_RAYS; }
                    }
                }
            }
        }
        // ---- every lane ORs each quarter's span into the words it owns ---------------------------------
        for (uint32_t todo = __ballot_sync(0xFFFFFFFFu, rings != 0u); todo; todo &= todo - 1u) {
            const int src = __ffs(todo) - 1;
            const int l = __shfl_sync(0xFFFFFFFFu, lo, src), h = __shfl_sync(0xFFFFFFFFu, hi, src);
            const uint32_t r = __shfl_sync(0xFFFFFFFFu, rings, src);
            // the span touches words [l >> 5, h >> 5] (and [0, (h - 3600) >> 5] where it wraps): lanes own words
            // lane + 32 i, so only the i whose 32 words meet those ranges can change (uniform tests)
            const uint32_t need = ((0xFu << (l >> 10)) & (0xFu >> (3 - (min(h, SMH_LSD_RAYS - 1) >> 10)))) |
                                  (h >= SMH_LSD_RAYS ? 0xFu >> (3 - ((h - SMH_LSD_RAYS) >> 10)) : 0u);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (!((need >> i) & 1u)) continue;
                uint32_t mk = range_mask(lane + 32 * i, l, h);
                if (h >= SMH_LSD_RAYS) mk |= range_mask(lane + 32 * i, l - SMH_LSD_RAYS, h - SMH_LSD_RAYS);  // uniform
#pragma unroll
                for (int j = 0; j < SMH_LSD_FILTER_MAX_WINDOWS; j++)
                    if ((r >> j) & 1u) acc[j][i] |= mk;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < SMH_LSD_FILTER_MAX_WINDOWS; j++)
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (acc[j][i] && lane + 32 * i < SMH_LSD_BITMAP_WORDS) atomicOr(&sh.bits[j][lane + 32 * i], acc[j][i]);
    const uint32_t n2 = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(near2));
    const uint32_t f2 = __reduce_min_sync(0xFFFFFFFFu, __float_as_uint(far2));
    if (lane == 0) {
        if (n2) atomicMax(&sh.near2, n2);
        if (f2 != __float_as_uint(3.0e38f)) atomicMin(&sh.far2, f2);
    }
}

// position of the n-th (0-based) set bit of mask (n < popc(mask))
__device__ __forceinline__ int nth_set_bit(uint32_t mask, int n) {
    int pos = 0;
#pragma unroll
    for (int s = 16; s; s >>= 1) {
        const int c = __popc((mask >> pos) & ((1u << s) - 1u));
        if (n >= c) { n -= c; pos += s; }
    }
    return pos;
}

// Stage 3: AND the usable windows, compact the surviving ray indices into sh.rays (ascending).
// Returns the number of rays (uniform).  Two barriers inside; clears the per-seed scratch.
template <int T>
__device__ __forceinline__ uint32_t lsd_build_ray_list(float px, float py, const smh_lsd_filter& F, LsdShared& sh) {
    static_assert(T >= 128, "one thread per bitmap word");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t wv = 0;
    if (tid < SMH_LSD_BITMAP_WORDS) {
        wv = 0xFFFFFFFFu;
#pragma unroll
        for (int j = 0; j < SMH_LSD_FILTER_MAX_WINDOWS; j++)
            if (j < F.nw && smh_lsd_filter_window_on(&F, j, px, py)) wv &= sh.bits[j][tid];
        if (tid == SMH_LSD_BITMAP_WORDS - 1) wv &= (1u << (SMH_LSD_RAYS & 31)) - 1u;  // 3600 = 112*32 + 16
    }
    const uint32_t c = __popc(wv);
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) sh.wsum[warp] = incl;
    if (tid < SMH_LSD_BITMAP_WORDS) sh.bits[0][tid] = wv;  // window 0's bitmap now holds the AND
    __syncthreads();
    uint32_t off = incl - c, total = 0;
#pragma unroll
    for (int k = 0; k < T / 32; k++) { const uint32_t s = sh.wsum[k]; if (k < warp) off += s; total += s; }
    if (tid < SMH_LSD_BITMAP_WORDS) sh.pre[tid] = off;  // rays before word tid
    __syncthreads();
    // slot t of the list: the word whose prefix range holds t, then the (t - prefix)-th set bit of it
    for (uint32_t t = tid; t < total; t += T) {
        int lo = 0, hi = SMH_LSD_BITMAP_WORDS - 1;  // last word with pre[word] <= t
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (sh.pre[mid] <= t) lo = mid; else hi = mid - 1;
        }
        sh.rays[t] = (uint16_t)(lo * 32 + nth_set_bit(sh.bits[0][lo], (int)(t - sh.pre[lo])));
    }
    __syncthreads();
    for (int i = tid; i < SMH_LSD_FILTER_MAX_WINDOWS * (SMH_LSD_BITMAP_WORDS + 3); i += T) (&sh.bits[0][0])[i] = 0u;
    if (tid == 0) { sh.near2 = 0u; sh.far2 = __float_as_uint(3.0e38f); }
    return total;  // the next writers of bits / near2 / far2 are behind the barrier of the ray search or of the next seed pick
}

// The rays of one seed, spread over the CTA; the winner is the LAST index among the longest rays
// (rayon's right-biased reduce, vision-cpu/src/lib.rs:440-446, D6).  n_rays < 0: all 3600 rays,
// else the first n_rays entries of `list`.  Every thread returns best = {len, ex, ey} (one
// barrier; keys/ends may be reused after the next one).
template <int T, typename MARCH>
__device__ __forceinline__ void longest_ray_cta(const float2* __restrict__ rays, const uint16_t* list, int n_rays,
                                                float px, float py, float (&best)[3],
                                                unsigned long long* s_keys, float2* s_ends, MARCH march_one) {
    const int tid = threadIdx.x;
    unsigned long long best_key = 0ull;  // (len_bits << 32) | k ; len >= 0 so the bit pattern is monotone
    float bex = px, bey = py;
    bool any = false;
    const int n = n_rays < 0 ? SMH_LSD_RAYS : n_rays;
    for (int i = tid; i < n; i += T) {
        const int k = n_rays < 0 ? i : (int)list[i];
        const float2 d = __ldg(&rays[k]);
        float ex, ey;
        const float len = march_one(d.x, d.y, ex, ey);
        const unsigned long long key = ((unsigned long long)__float_as_uint(len) << 32) | (uint32_t)k;
        if (!any || key >= best_key) { best_key = key; bex = ex; bey = ey; any = true; }
    }
    // warp argmax on the 64-bit key (threads without a ray carry key 0 and never claim the slot)
    unsigned long long wkey = best_key;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, wkey, o);
        wkey = other > wkey ? other : wkey;
    }
    const int warp = tid >> 5;
    const bool warp_any = __any_sync(0xFFFFFFFFu, any);
    if ((tid & 31) == 0 && !warp_any) { s_keys[warp] = 0ull; s_ends[warp] = make_float2(px, py); }
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

// Clear columns [xlo, xhi] (pixel coordinates) of row r of the tile-row held row-major in shared memory.
__device__ __forceinline__ void chunk_clear(uint32_t* chunk, uint32_t tx, uint32_t pad, int r, int xlo, int xhi) {
    const int Xlo = xlo + (int)pad, Xhi = xhi + (int)pad;
    for (int wd = Xlo >> 5; wd <= (Xhi >> 5); wd++) {
        uint32_t m = 0xFFFFFFFFu;
        if (wd == (Xlo >> 5)) m &= 0xFFFFFFFFu << (Xlo & 31);
        if (wd == (Xhi >> 5)) m &= 0xFFFFFFFFu >> (31 - (Xhi & 31));
        atomicAnd(&chunk[(uint32_t)r * tx + (uint32_t)wd], ~m);
    }
}

__device__ __forceinline__ void chunk_set(uint32_t* bm, uint32_t tx, uint32_t pad, int r, int xlo, int xhi) {
    const int Xlo = xlo + (int)pad, Xhi = xhi + (int)pad;
    for (int wd = Xlo >> 5; wd <= (Xhi >> 5); wd++) {
        uint32_t m = 0xFFFFFFFFu;
        if (wd == (Xlo >> 5)) m &= 0xFFFFFFFFu << (Xlo & 31);
        if (wd == (Xhi >> 5)) m &= 0xFFFFFFFFu >> (31 - (Xhi & 31));
        atomicOr(&bm[(uint32_t)r * tx + (uint32_t)wd], m);
    }
}

// lsd::find_lines for whole frames (vision-common/src/lsd.rs:60-105).  One CTA works on one frame
// at a time (persistent CTAs pull frames from a counter).  Per frame the bitmap is walked one
// tile-row (32 pixel rows) at a time, held row-major in shared memory as the set of pixels that
// can still become seeds:
//   * lines found so far strike off what they suppress (lsd.rs:84-89) — per (row, line) the
//     interval that is suppressed for sure is cleared with word masks, the few remaining pixels
//     take the reference's float test — so the next seed is simply the first set bit;
//   * per seed: dead-cluster test, get_centre, the exact ray filter, the search over the surviving
//     rays with a last-index argmax, get_centre of the far end;
//   * a seed whose first ring is empty proves a whole disc of seeds hopeless (dead cluster): they
//     are struck off in bulk (or, when counting seeds for the statistics, dismissed one by one).
// Registers: capped at 64 (four CTAs per SM).  Alone the search is as fast with 80 registers and three CTAs
// (4.41 against 4.45 ms per 1024 frames), but with a second batch in flight the smaller footprint lets the next
// batch's pixel pass and line search move in earlier: 163.9 k -> 170.8 k frames/s (56 / 60 / 64 / 68 / 72 / 80
// registers: 169.4 / 169.2 / 170.8 / 164.3 / 165.0 / 163.9 k).  LSD_MAXNREG=0 restores __launch_bounds__.
#ifndef LSD_MAXNREG
#define LSD_MAXNREG 64
#endif
template <bool STATS>
#if LSD_MAXNREG > 0
__global__ void __maxnreg__(LSD_MAXNREG) k_lsd(const LsdArgs a) {
#else
__global__ void __launch_bounds__(LSD_THREADS, LSD_MIN_CTAS) k_lsd(const LsdArgs a) {
#endif
    constexpr int T = LSD_THREADS;
    __shared__ LsdShared sh;
    extern __shared__ uint32_t chunk[];  // [32][tx] current tile-row, row-major: pixels that can still become seeds
    uint32_t* fringe = chunk + 32u * a.g.tx;  // [32][tx] pixels inside the uncertain rim of some line's band
    uint32_t* win = fringe + 32u * a.g.tx;    // [LSD_WIN_ROWS][8] neighbourhood of the current seed

    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t TX = a.g.tx, W = a.g.w, H = a.g.h, PAD = a.g.pad;
    const uint32_t NW = 32u * TX;             // words per tile-row
    const uint32_t E = (NW + T - 1) / T;      // consecutive raster words owned by a thread
    const uint32_t own0 = min((uint32_t)tid * E, NW), own1 = min(own0 + E, NW);
    const bool fast = a.gap_limit >= 1;  // march_fast's precondition
    smh_lsd_filter F;
    smh_lsd_filter_init(&F, a.prune ? a.gap_limit : 0);
    const bool dead_ok = F.nw >= 1 && F.border[0] == 0;  // "ring 1 empty => no line" holds for every seed
    const bool hybrid = F.nw > 0 && a.gap_limit <= 32;   // filtered search (march_window / ray_block_* preconditions)
    // samples 0 .. win_steps-1 of any ray stay inside the seed's window (|P_n - seed| < n + 1)
    const int win_steps = F.radius >= 65 ? 64 : (F.radius >= 33 ? 32 : 0);

    for (int i = tid; i < SMH_LSD_FILTER_MAX_WINDOWS * (SMH_LSD_BITMAP_WORDS + 3); i += T) (&sh.bits[0][0])[i] = 0u;
    if (tid == 0) { sh.surv_n = 0u; sh.near2 = 0u; sh.far2 = __float_as_uint(3.0e38f); }

    while (true) {
        __syncthreads();
        if (tid == 0) { sh.frame = atomicAdd(a.counter, 1u); sh.n_dead = 0u; }
        __syncthreads();
        const uint32_t f = sh.frame;
        if (f >= a.n_frames) break;
        float* out_lines = a.lines + (uint64_t)f * SMH_LSD_MAX_LINES * 4;
        if (a.open && a.open[f] == 0) {  // map closed: the reference returns before any line search
            if (tid == 0) { a.n_lines[f] = 0; if (STATS) { a.stats[3 * f] = 0; a.stats[3 * f + 1] = 0; a.stats[3 * f + 2] = 0; } }
            continue;
        }
        const Bits im{a.bits + (uint64_t)f * a.bits_stride, TX, PAD, W, H};

        uint32_t steps = 0, seeds = 0, pruned = 0;
        int n = 0, par = 0;
        LSD_PROF_DECL
        for (uint32_t tyi = 0; tyi < a.g.ty && n < SMH_LSD_MAX_LINES; tyi++) {
            // ---- load the tile-row (contiguous in memory), transposed to row-major ------------
            {
                const uint32_t* src = im.p + (uint64_t)tyi * NW;
                for (uint32_t i = tid; i < NW; i += T) {
                    const uint32_t o = (i & 31u) * TX + (i >> 5);
                    chunk[o] = __ldg(src + i);
                    fringe[o] = 0u;
                }
                if (tyi + 1 < a.g.ty)  // pull the next tile-row into L2 while this one is searched
                    for (uint32_t i = tid * 32u; i < NW; i += T * 32u) asm volatile("prefetch.global.L2 [%0];" :: "l"(src + NW + i));
            }
            __syncthreads();
            const int row0 = (int)(tyi * 32u);
            const int nd = STATS ? 0 : (int)min(sh.n_dead, (uint32_t)LSD_DEAD);
            // ---- strike off what the lines (and dead clusters) found so far cover for sure; mark
            //      the fringe of every line's band for the exact test ------------------------------
            for (int i = tid; i < 32 * (n + nd); i += T) {
                const int r = i & 31, l = i >> 5, y = row0 + r;
                if (l < n) {
                    const Band b = line_band(sh.lines[l], (float)y, (int)W);
                    if (b.ilo <= b.ihi) {
                        chunk_clear(chunk, TX, PAD, r, b.ilo, b.ihi);
                        if (b.olo < b.ilo) chunk_set(fringe, TX, PAD, r, b.olo, b.ilo - 1);
                        if (b.ohi > b.ihi) chunk_set(fringe, TX, PAD, r, b.ihi + 1, b.ohi);
                    } else if (b.olo <= b.ohi) chunk_set(fringe, TX, PAD, r, b.olo, b.ohi);
                } else {
                    const float4 d = sh.dead[l - n];
                    const float dy = (float)y - d.y, r2 = d.w - dy * dy;
                    if (r2 > 0.0f) {
                        const float hw = sqrtf(r2) - 0.01f;
                        const int xlo = max((int)ceilf(d.x - hw), 0), xhi = min((int)floorf(d.x + hw), (int)W - 1);
                        if (xlo <= xhi) chunk_clear(chunk, TX, PAD, r, xlo, xhi);
                    }
                }
            }
            __syncthreads();
            // ---- the reference's float test for the fringe pixels (own words only) -------------
            bool any_white = false;
            for (uint32_t i = own0; i < own1; i++) {
                uint32_t m = chunk[i];
                uint32_t todo = m & fringe[i];
                if (todo) {
                    const uint32_t r = i / TX, wd = i - r * TX;
                    for (; todo; todo &= todo - 1u) {
                        const int b = __ffs(todo) - 1;
                        const float fx = (float)((int)(wd * 32u) + b - (int)PAD), fy = (float)(row0 + (int)r);
                        for (int l = 0; l < n; l++)
                            if (suppressed_by(sh.lines[l], fx, fy)) { m &= ~(1u << b); break; }
                    }
                    chunk[i] = m;
                }
                any_white |= m != 0u;
            }
            const bool some_white = __syncthreads_or(any_white);
            LSD_PROF(0);  // tile-row load + strike-off
            if (!some_white) continue;

            // ---- seeds of this tile-row, in raster order ---------------------------------------
            while (true) {
                uint32_t cand = 0xFFFFFFFFu;  // first remaining pixel among this thread's words: raster word index * 32 + bit
                for (uint32_t i = own0; i < own1; i++) {
                    const uint32_t m = chunk[i];
                    if (m) { cand = i * 32u + (uint32_t)(__ffs(m) - 1); break; }
                }
                const uint32_t wmin = __reduce_min_sync(0xFFFFFFFFu, cand);
                if (lane == 0) sh.smin[par][tid >> 5] = wmin;
                __syncthreads();
                uint32_t seed = 0xFFFFFFFFu;
#pragma unroll
                for (int k = 0; k < T / 32; k++) seed = min(seed, sh.smin[par][k]);
                par ^= 1;
                LSD_PROF(1);  // seed pick
                if (seed == 0xFFFFFFFFu) break;  // uniform
                if (cand == seed) chunk[seed >> 5] &= ~(1u << (seed & 31u));  // the owner takes it off
                const uint32_t sr = (seed >> 5) / TX, swd = (seed >> 5) - sr * TX;
                const int isx = (int)(swd * 32u + (seed & 31u)) - (int)PAD, isy = row0 + (int)sr;
                const float sx = (float)isx, sy = (float)isy;
                seeds++;
                const bool have_dead = dead_ok && lane < (int)min(sh.n_dead, (uint32_t)LSD_DEAD);
                if (dead_ok) {
                    // deep inside a remembered dead cluster?  get_centre moves a seed by at most 2.5 px per
                    // axis, so the raw pixel decides when it is sqrt(12.5) closer.  Lane l tests entry l.
                    bool in = false;
                    if (have_dead) {
                        const float4 d = sh.dead[lane];
                        const float ex = sx - d.x, ey = sy - d.y;
                        in = ex * ex + ey * ey < d.w;
                    }
                    if (__any_sync(0xFFFFFFFFu, in)) { pruned++; LSD_PROF(2); continue; }  // uniform: same data in every warp
                }
                float px, py;
                float best[3];
                if (hybrid) {
                    // ---- the seed's neighbourhood into shared memory: one memory round trip -----------
                    const WinView wv = lsd_load_window<T>(im, win, isx, isy, F.radius + LSD_WIN_EXTRA);
                    __syncthreads();
                    LSD_PROF(4);  // window load
                    LSD_CNT(10, 1);
                    get_centre_warp(wv, sx, sy, px, py);  // same value in every warp
                    LSD_PROF(9);  // get_centre
                    if (dead_ok) {
                        bool in = false;
                        if (have_dead) {
                            const float4 d = sh.dead[lane];
                            const float ex = px - d.x, ey = py - d.y;
                            in = ex * ex + ey * ey < d.z;
                        }
                        if (__any_sync(0xFFFFFFFFu, in)) { pruned++; LSD_PROF(2); continue; }  // (win is rewritten behind the next seed-pick barrier)
                    }
                    lsd_scan_window<T>(wv, px, py, F, sh);
                    LSD_PROF(3);  // ring scan
                    __syncthreads();
                    LSD_PROF(15);  // barrier after the ring scan
                    lsd_mark_rays<T>(a.span_tab, px, py, F, sh);
                    __syncthreads();
                    LSD_PROF(5);  // ray marking
                    const float near2 = __uint_as_float(sh.near2), far2 = __uint_as_float(sh.far2);
                    const int n_rays = (int)lsd_build_ray_list<T>(px, py, F, sh);
                    LSD_PROF(6);  // ray list
                    LSD_CNT(11, n_rays);
                    if (n_rays == 0) {  // uniform: no ray can exceed the minimum length
                        pruned++;
                        // every white pixel is within sqrt(near2) of the seed or beyond dfar: a seed at distance
                        // t < rho sees none of them in its first ring either
                        const float k1 = (float)F.k1;
                        const float dfar = fminf(sqrtf(far2), (float)(F.radius - 1));
                        const float rho = fminf(k1 - 0.75f - sqrtf(near2), dfar - (2.0f * k1 - 0.25f)) - 0.05f;
                        if (dead_ok && rho > 0.5f) {  // uniform
                            const float raw = fmaxf(rho - 3.6f, 0.0f);  // radius for raw (un-centred) seed pixels
                            if (tid == 0) {
                                sh.dead[sh.n_dead % (uint32_t)LSD_DEAD] = make_float4(px, py, rho * rho, raw * raw);
                                sh.n_dead = sh.n_dead + 1u;  // read by the other threads only after the next barrier
                            }
                            if (!STATS && raw > 0.0f) {  // strike the disc off this tile-row right away
                                const int r = tid;  // one row per thread
                                if (r < 32) {
                                    const float dy = (float)(row0 + r) - py, r2 = raw * raw - dy * dy;
                                    if (r2 > 0.0f) {
                                        const float hw = sqrtf(r2) - 0.01f;
                                        const int xlo = max((int)ceilf(px - hw), 0), xhi = min((int)floorf(px + hw), (int)W - 1);
                                        if (xlo <= xhi) chunk_clear(chunk, TX, PAD, r, xlo, xhi);
                                    }
                                }
                            }
                        }
                        __syncthreads();  // uniform branch: the next seed pick reads the cleared words, the next window load rewrites win
                        continue;
                    }
                    longest_ray_hybrid<T, STATS>(im, wv, win_steps, a.rays, a.ckpt, a.n_ckpt, sh.rays, n_rays, px, py, a.gap_limit, best, sh, steps);
                } else {
                    get_centre_warp(im, sx, sy, px, py);  // same value in every warp
                    LSD_PROF(3);
                    if (fast)
                        longest_ray_cta<T>(a.rays, nullptr, -1, px, py, best, sh.keys, sh.ends,
                                           [&](float dx, float dy, float& ex, float& ey) { return march_fast<STATS>(im, px, py, a.gap_limit, dx, dy, ex, ey, steps); });
                    else
                        longest_ray_cta<T>(a.rays, nullptr, -1, px, py, best, sh.keys, sh.ends,
                                           [&](float dx, float dy, float& ex, float& ey) { return march(im, px, py, a.max_gap, dx, dy, ex, ey, steps); });
                }
                LSD_PROF(7);  // march
                if (best[0] > SMH_LSD_MIN_LEN_SQ) {  // uniform
                    float c1x, c1y;
                    get_centre_warp(im, best[1], best[2], c1x, c1y);
                    const float4 line = make_float4(px, py, c1x, c1y);
                    if (tid == 0) {
                        sh.lines[n] = line;  // read by the others only behind the barriers below
                        out_lines[4 * n + 0] = px; out_lines[4 * n + 1] = py;
                        out_lines[4 * n + 2] = c1x; out_lines[4 * n + 3] = c1y;
                    }
                    n++;
                    if (n == SMH_LSD_MAX_LINES) break;  // `break 'row`
                    // the new line suppresses pixels for good: strike them off this tile-row now (own words;
                    // approximate distance first, the reference's float test only in the uncertain rim)
                    {
                        const float ldx = line.z - line.x, ldy = line.w - line.y, l2 = ldx * ldx + ldy * ldy;
                        const float in2 = 48.9f * l2, out2 = 51.1f * l2;
                        for (uint32_t i = own0; i < own1; i++) {
                            uint32_t m = chunk[i];
                            if (m) {
                                const uint32_t r = i / TX, wd = i - r * TX;
                                const float fy = (float)(row0 + (int)r);
                                for (uint32_t bits = m; bits; bits &= bits - 1u) {
                                    const int b = __ffs(bits) - 1;
                                    const float fx = (float)((int)(wd * 32u) + b - (int)PAD);
                                    const float cr = (fx - line.x) * ldy - (fy - line.y) * ldx, c2 = cr * cr;
                                    if (c2 < in2 || (!(c2 > out2) && suppressed_by(line, fx, fy))) m &= ~(1u << b);
                                }
                                chunk[i] = m;
                            }
                        }
                    }
                }
            }
            __syncthreads();
            LSD_PROF(8);  // accept + strike-off of new lines, end of tile-row
        }
        LSD_PROF_FLUSH;
        if (tid == 0) a.n_lines[f] = (uint32_t)n;
        if (STATS) {
            steps = __reduce_add_sync(0xFFFFFFFFu, steps);
            if (lane == 0) atomicAdd(&a.stats[3 * f + 1], (unsigned long long)steps);
            if (tid == 0) { a.stats[3 * f] = seeds; a.stats[3 * f + 2] = pruned; }
        }
    }
}

// single find_longest_line call (Vision::find_longest_line) on the L8 mask; out = {p0x,p0y,p1x,p1y,len}
__global__ void __launch_bounds__(LSD_THREADS) k_longest(const uint8_t* lsd, uint32_t pitch, uint32_t pad, uint32_t w,
                                                         uint32_t h, const float2* rays, float px, float py,
                                                         float max_gap, float* out) {
    __shared__ unsigned long long s_keys[LSD_THREADS / 32];
    __shared__ float2 s_ends[LSD_THREADS / 32];
    const ImgL8 im{Img{lsd + pad, pitch, w, h}, w, h};
    uint32_t steps = 0;
    float best[3];
    longest_ray_cta<LSD_THREADS>(rays, nullptr, -1, px, py, best, s_keys, s_ends,
                                 [&](float dx, float dy, float& ex, float& ey) { return march(im, px, py, max_gap, dx, dy, ex, ey, steps); });
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
