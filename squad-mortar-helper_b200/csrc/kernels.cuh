// kernels.cuh — shared device code of the map-analysis vision path (sm_100a) and its small kernels.
//
//   per-pixel helpers   luma / marker predicate on BGRA words, bit-identical to smh_pixel.h (used by k_pixels2,
//                       kernels_px.cuh: the fused pixel pass)
//   line-search pieces  the tile-major bitmap, the literal and the fast ray march, the one-ray-per-warp block march,
//                       get_centre, the suppression test and band (used by k_lsd2, kernels_lsd.cuh)
//   k_lsd      lsd::find_lines with EVERY ray of every seed marched (the reference's work; any max_gap): the
//              literal form behind SMH_BATCH_LSD_EXHAUSTIVE and for max_gap outside k_lsd2's range
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
#ifndef STG_POLICY
#define STG_POLICY "wb"  // (measured: 2-3 % faster than L1::no_allocate on the pixel pass)
#endif
__device__ __forceinline__ void stg_v4(void* p, const uint4& v) {
    asm volatile("st.global." STG_POLICY ".v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t smem_addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(smem_addr));
    return r;
}
__device__ __forceinline__ void stg_u32(void* p, uint32_t v) {
    asm volatile("st.global." STG_POLICY ".u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// ------------------------------------------------------------------------------------------
// per-pixel helpers of the fused pixel pass (kernels_px.cuh)
// ------------------------------------------------------------------------------------------
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

// Vector-level fast rejects work on the OR of the four BGRA words: a byte of the OR is >= any of the four bytes it
// combines and keeps their bit 7.
__device__ __forceinline__ uint32_t or4(const uint4& v) { return v.x | v.y | v.z | v.w; }

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
    // any (x, y): the coordinates are clamped into the image, the caller masks the result of a sample outside it
    __device__ __forceinline__ uint32_t white_anywhere(uint32_t x, uint32_t y) const {
        return white(min(x, w - 1u), min(y, h - 1u)) ? 1u : 0u;
    }
};

struct LsdArgs {
    const uint32_t* bits; uint64_t bits_stride;  // tile-major bitmaps, words per frame
    BitsGeom g;
    const float2* rays;  // 3600 (cos, sin) pairs built on the host with libm (D11)
    const float2* ckpt; uint32_t n_ckpt;  // accumulated ray offsets at every LSD_CKPT-th step, [3600][n_ckpt]
    float max_gap;
    int gap_limit;        // ceil(max_gap): `gap.0 >= max_gap` on the integer-valued counter
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

#ifndef LSD_THREADS_N
#define LSD_THREADS_N 256
#endif
constexpr int LSD_THREADS = LSD_THREADS_N;        // k_lsd / k_longest: threads per CTA
constexpr int LSD_WIN_EXTRA = 3;                  // get_centre moves the seed by < 3 px

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

// OR `m` into a shared bitmap word unless it is already there (most marks of a seed repeat: the
// pixels of a line all point the same way, and same-address shared atomics serialise)
__device__ __forceinline__ void or_if_new(uint32_t* w, uint32_t m) {
    if ((*reinterpret_cast<volatile uint32_t*>(w) & m) != m) atomicOr(w, m);
}

// bits lo..hi (any integers) that fall into word w of a bitmap
__device__ __forceinline__ uint32_t range_mask(int w, int lo, int hi) {
    const int s = max(lo - 32 * w, 0), e = min(hi - 32 * w, 31);
    return s <= e ? ((0xFFFFFFFFu << s) & (0xFFFFFFFFu >> (31 - e))) : 0u;
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

// lsd::find_lines for whole frames (vision-common/src/lsd.rs:60-105) with EVERY ray of every seed marched, as the
// reference does (SMH_BATCH_LSD_EXHAUSTIVE: its ray-step counter equals the reference algorithm's; also the path for
// max_gap outside k_lsd2's 1..32).  One CTA works on one frame at a time (persistent CTAs pull frames from a counter).
// Per frame the bitmap is walked one tile-row (32 pixel rows) at a time, held row-major in shared memory as the
// set of pixels that can still become seeds: lines found so far strike off what they suppress (lsd.rs:84-89) —
// per (row, line) the interval that is suppressed for sure is cleared with word masks, the pixels of the uncertain
// rim take the reference's float test — so the next seed is simply the first set bit.
template <bool STATS>
__global__ void __launch_bounds__(LSD_THREADS, 3) k_lsd(const LsdArgs a) {
    constexpr int T = LSD_THREADS;
    __shared__ float4 s_lines[SMH_LSD_MAX_LINES];
    __shared__ unsigned long long s_keys[T / 32];
    __shared__ float2 s_ends[T / 32];
    __shared__ uint32_t s_min[2][T / 32];  // double-buffered: a seed may be dismissed without another barrier
    __shared__ uint32_t s_frame;
    extern __shared__ uint32_t chunk[];  // [32][tx] current tile-row, row-major: pixels that can still become seeds
    uint32_t* fringe = chunk + 32u * a.g.tx;  // [32][tx] pixels inside the uncertain rim of some line's band

    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t TX = a.g.tx, W = a.g.w, H = a.g.h, PAD = a.g.pad;
    const uint32_t NW = 32u * TX;             // words per tile-row
    const uint32_t E = (NW + T - 1) / T;      // consecutive raster words owned by a thread
    const uint32_t own0 = min((uint32_t)tid * E, NW), own1 = min(own0 + E, NW);
    const bool fast = a.gap_limit >= 1;  // march_fast's precondition

    while (true) {
        __syncthreads();
        if (tid == 0) s_frame = atomicAdd(a.counter, 1u);
        __syncthreads();
        const uint32_t f = s_frame;
        if (f >= a.n_frames) break;
        float* out_lines = a.lines + (uint64_t)f * SMH_LSD_MAX_LINES * 4;
        if (a.open && a.open[f] == 0) {  // map closed: the reference returns before any line search
            if (tid == 0) { a.n_lines[f] = 0; if (STATS) { a.stats[3 * f] = 0; a.stats[3 * f + 1] = 0; a.stats[3 * f + 2] = 0; } }
            continue;
        }
        const Bits im{a.bits + (uint64_t)f * a.bits_stride, TX, PAD, W, H};

        uint32_t steps = 0, seeds = 0;
        int n = 0, par = 0;
        for (uint32_t tyi = 0; tyi < a.g.ty && n < SMH_LSD_MAX_LINES; tyi++) {
            // ---- load the tile-row (contiguous in memory), transposed to row-major ------------
            {
                const uint32_t* src = im.p + (uint64_t)tyi * NW;
                for (uint32_t i = tid; i < NW; i += T) {
                    const uint32_t o = (i & 31u) * TX + (i >> 5);
                    chunk[o] = __ldg(src + i);
                    fringe[o] = 0u;
                }
            }
            __syncthreads();
            const int row0 = (int)(tyi * 32u);
            // ---- strike off what the lines found so far cover for sure; mark the fringe of every line's band ----
            for (int i = tid; i < 32 * n; i += T) {
                const int r = i & 31, l = i >> 5, y = row0 + r;
                const Band b = line_band(s_lines[l], (float)y, (int)W);
                if (b.ilo <= b.ihi) {
                    chunk_clear(chunk, TX, PAD, r, b.ilo, b.ihi);
                    if (b.olo < b.ilo) chunk_set(fringe, TX, PAD, r, b.olo, b.ilo - 1);
                    if (b.ohi > b.ihi) chunk_set(fringe, TX, PAD, r, b.ihi + 1, b.ohi);
                } else if (b.olo <= b.ohi) chunk_set(fringe, TX, PAD, r, b.olo, b.ohi);
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
                            if (suppressed_by(s_lines[l], fx, fy)) { m &= ~(1u << b); break; }
                    }
                    chunk[i] = m;
                }
                any_white |= m != 0u;
            }
            if (!__syncthreads_or(any_white)) continue;

            // ---- seeds of this tile-row, in raster order ---------------------------------------
            while (true) {
                uint32_t cand = 0xFFFFFFFFu;  // first remaining pixel among this thread's words: raster word index * 32 + bit
                for (uint32_t i = own0; i < own1; i++) {
                    const uint32_t m = chunk[i];
                    if (m) { cand = i * 32u + (uint32_t)(__ffs(m) - 1); break; }
                }
                const uint32_t wmin = __reduce_min_sync(0xFFFFFFFFu, cand);
                if (lane == 0) s_min[par][tid >> 5] = wmin;
                __syncthreads();
                uint32_t seed = 0xFFFFFFFFu;
#pragma unroll
                for (int k = 0; k < T / 32; k++) seed = min(seed, s_min[par][k]);
                par ^= 1;
                if (seed == 0xFFFFFFFFu) break;  // uniform
                if (cand == seed) chunk[seed >> 5] &= ~(1u << (seed & 31u));  // the owner takes it off
                const uint32_t sr = (seed >> 5) / TX, swd = (seed >> 5) - sr * TX;
                const float sx = (float)((int)(swd * 32u + (seed & 31u)) - (int)PAD), sy = (float)(row0 + (int)sr);
                seeds++;
                float px, py;
                float best[3];
                get_centre_warp(im, sx, sy, px, py);  // same value in every warp
                if (fast)
                    longest_ray_cta<T>(a.rays, nullptr, -1, px, py, best, s_keys, s_ends,
                                       [&](float dx, float dy, float& ex, float& ey) { return march_fast<STATS>(im, px, py, a.gap_limit, dx, dy, ex, ey, steps); });
                else
                    longest_ray_cta<T>(a.rays, nullptr, -1, px, py, best, s_keys, s_ends,
                                       [&](float dx, float dy, float& ex, float& ey) { return march(im, px, py, a.max_gap, dx, dy, ex, ey, steps); });
                if (best[0] > SMH_LSD_MIN_LEN_SQ) {  // uniform
                    float c1x, c1y;
                    get_centre_warp(im, best[1], best[2], c1x, c1y);
                    const float4 line = make_float4(px, py, c1x, c1y);
                    if (tid == 0) {
                        s_lines[n] = line;  // read by the others only behind the barriers below
                        out_lines[4 * n + 0] = px; out_lines[4 * n + 1] = py;
                        out_lines[4 * n + 2] = c1x; out_lines[4 * n + 3] = c1y;
                    }
                    n++;
                    if (n == SMH_LSD_MAX_LINES) break;  // `break 'row`
                    // the new line suppresses pixels for good: strike them off this tile-row now (own words;
                    // approximate distance first, the reference's float test only in the uncertain rim)
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
            __syncthreads();
        }
        if (tid == 0) a.n_lines[f] = (uint32_t)n;
        if (STATS) {
            steps = __reduce_add_sync(0xFFFFFFFFu, steps);
            if (lane == 0) atomicAdd(&a.stats[3 * f + 1], (unsigned long long)steps);
            if (tid == 0) { a.stats[3 * f] = seeds; a.stats[3 * f + 2] = 0; }
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
