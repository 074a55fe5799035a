// smh_pixel.h — exact per-pixel arithmetic of the vision path, shared by the CUDA kernels
// and (for CPU-only exhaustive checks) by csrc/hostcheck.cpp.
//
// Everything here is IEEE binary32 round-to-nearest, NO fused multiply-add, evaluated in
// the order the reference's Rust writes it, because the results feed truncating casts and
// a 1-ulp difference flips output bytes.  On the device every float op is
// an explicit __f*_rn intrinsic (never contracted by nvcc); on the host the file must be
// compiled with -ffp-contract=off.
//
// Reference lines restated (relative to /root/reference):
//   luma8      image 0.23.14 rgb_to_luma; call sites vision-cpu/src/lib.rs:154,224,242
//   hsv        util/src/image.rs:159-187
//   is_marker  vision-common/src/markers/mod.rs:17-19,40-54
#pragma once
#include <stdint.h>
#include <math.h>

#include "../../include/smh_consts.h"

#if defined(__CUDA_ARCH__)
#define SMH_HD __host__ __device__ __forceinline__
#define SMH_MUL(a, b) __fmul_rn((a), (b))
#define SMH_ADD(a, b) __fadd_rn((a), (b))
#define SMH_SUB(a, b) __fsub_rn((a), (b))
#define SMH_DIV(a, b) __fdiv_rn((a), (b))
#elif defined(__CUDACC__)
#define SMH_HD __host__ __device__ __forceinline__
#define SMH_MUL(a, b) ((a) * (b))
#define SMH_ADD(a, b) ((a) + (b))
#define SMH_SUB(a, b) ((a) - (b))
#define SMH_DIV(a, b) ((a) / (b))
#else
#define SMH_HD static inline
#define SMH_MUL(a, b) ((a) * (b))
#define SMH_ADD(a, b) ((a) + (b))
#define SMH_SUB(a, b) ((a) - (b))
#define SMH_DIV(a, b) ((a) / (b))
#endif

// Rust `f as u8/u16` for the value ranges that occur here (finite or NaN, < 2^16):
// truncate toward zero, negative/NaN -> 0, saturate above.
SMH_HD uint32_t smh_trunc_sat(float f, uint32_t hi) {
    if (!(f > 0.0f)) return 0u;
    if (f >= (float)hi) return hi;
    return (uint32_t)f;
}

SMH_HD uint32_t smh_luma8(uint32_t r, uint32_t g, uint32_t b) {
    float l = SMH_ADD(SMH_ADD(SMH_MUL(SMH_LUMA_R, (float)r), SMH_MUL(SMH_LUMA_G, (float)g)),
                      SMH_MUL(SMH_LUMA_B, (float)b));
    return smh_trunc_sat(l, 255u);
}

// Exact HSV -> (h 0..359, s 0..100, v 0..100)
SMH_HD void smh_hsv(uint32_t r8, uint32_t g8, uint32_t b8, uint32_t* ho, uint32_t* so, uint32_t* vo) {
    const float r = SMH_DIV((float)r8, 255.0f), g = SMH_DIV((float)g8, 255.0f), b = SMH_DIV((float)b8, 255.0f);
    const float mx = fmaxf(r, fmaxf(g, b));
    const float mn = fminf(r, fminf(g, b));
    const float d = SMH_SUB(mx, mn);
    float h;
    if (mx == mn) h = 0.0f;
    else if (mx == r) h = SMH_MUL(60.0f, fmodf(SMH_DIV(SMH_SUB(g, b), d), 6.0f));
    else if (mx == g) h = SMH_MUL(60.0f, SMH_ADD(SMH_DIV(SMH_SUB(b, r), d), 2.0f));
    else h = SMH_MUL(60.0f, SMH_ADD(SMH_DIV(SMH_SUB(r, g), d), 4.0f));
    const float s = SMH_DIV(SMH_MUL(100.0f, d), mx);  // 0/0 -> NaN -> 0
    const float v = SMH_MUL(100.0f, mx);
    float hm = fmodf(h, 360.0f);
    if (hm < 0.0f) hm = SMH_ADD(hm, 360.0f);
    *ho = smh_trunc_sat(hm, 65535u);
    *so = smh_trunc_sat(s, 255u);
    *vo = smh_trunc_sat(v, 255u);
}

SMH_HD bool smh_is_marker_exact(uint32_t r, uint32_t g, uint32_t b) {
    uint32_t h, s, v;
    smh_hsv(r, g, b, &h, &s, &v);
    if (s < SMH_MIN_SAT) return false;
    const int hi = (int)h, si = (int)s, vi = (int)v;
    const int MH[3] = {SMH_ALPHA_H, SMH_BRAVO_H, SMH_CHARLIE_H};
    const int MS[3] = {SMH_ALPHA_S, SMH_BRAVO_S, SMH_CHARLIE_S};
    const int MV[3] = {SMH_ALPHA_V, SMH_BRAVO_V, SMH_CHARLIE_V};
    bool ok = false;
#pragma unroll
    for (int t = 0; t < 3; t++) {
        const int dh = hi > MH[t] ? hi - MH[t] : MH[t] - hi;
        const int ds = si > MS[t] ? si - MS[t] : MS[t] - si;
        int da = si - (MS[t] - SMH_ARC_SAT);
        da = da < 0 ? -da : da;
        const int dv = vi > MV[t] ? vi - MV[t] : MV[t] - vi;
        ok = ok || (dh <= SMH_HUE_TOL && (ds <= SMH_SAT_TOL || da <= SMH_SAT_TOL) && dv <= SMH_VIB_TOL);
    }
    return ok;
}

// Integer fast-reject: a NECESSARY condition of smh_is_marker_exact, a few integer ops.
// Every marker colour has V >= min(mv)-15 = 70 and S >= 35; with V = trunc(100*max/255)
// and S = trunc(100*delta/max) that needs max8 >= 179 and 100*delta8 >= 35*max8 up to
// float rounding; the bounds below keep one unit of slack on each and are proven against
// the exact predicate over all 2^24 colours (tests/test_pixel_math.py, CPU; and on the GPU).
SMH_HD bool smh_marker_prefilter(uint32_t r, uint32_t g, uint32_t b) {
    const uint32_t mx = r > g ? (r > b ? r : b) : (g > b ? g : b);
    const uint32_t mn = r < g ? (r < b ? r : b) : (g < b ? g : b);
    return mx >= 178u && 100u * (mx - mn) >= 34u * mx;
}

SMH_HD bool smh_is_marker(uint32_t r, uint32_t g, uint32_t b) {
    return smh_marker_prefilter(r, g, b) && smh_is_marker_exact(r, g, b);
}

// OCR colour classes (vision-cpu/src/lib.rs:39-53,188-193): with mono = sum over all ordered
// channel pairs |ci-cj| = 4*(max-min):  mono <= 3 <=> max == min;  mono <= 48 <=> max-min <= 12.
SMH_HD bool smh_ocr_core(uint32_t r, uint32_t g, uint32_t b) {  // bright and monochrome
    return r == g && g == b && r >= SMH_OCR_BRIGHT;
}
SMH_HD bool smh_ocr_edge(uint32_t r, uint32_t g, uint32_t b) {  // candidate that needs a core neighbour
    const uint32_t mx = r > g ? (r > b ? r : b) : (g > b ? g : b);
    const uint32_t mn = r < g ? (r < b ? r : b) : (g < b ? g : b);
    return 4u * (mx - mn) <= SMH_OCR_MONO_EDGE && mn >= SMH_OCR_BRIGHT_EDGE;
}
SMH_HD bool smh_button_red(uint32_t r, uint32_t g, uint32_t b) {  // vision-cpu/src/lib.rs:120-128
    const int dr = (int)r - SMH_BTN_COLOR_R, dg = (int)g - SMH_BTN_COLOR_G, db = (int)b - SMH_BTN_COLOR_B;
    return (dr < 0 ? -dr : dr) <= SMH_BTN_TOLERANCE && (dg < 0 ? -dg : dg) <= SMH_BTN_TOLERANCE &&
           (db < 0 ? -db : db) <= SMH_BTN_TOLERANCE;
}
