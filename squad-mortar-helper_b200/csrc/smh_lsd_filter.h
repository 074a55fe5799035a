// smh_lsd_filter.h — exact ray filter of the line search, shared by k_lsd and (for the CPU-only
// soundness test) by csrc/hostcheck.cpp.
//
// lsd::find_lines (vision-common/src/lsd.rs:60-105) keeps the longest of 3600 rays of a seed only
// when its squared length exceeds 2500 (lsd.rs:95), so rays that provably stay <= 50 px can be
// skipped without changing find_lines' output.  The march (vision-cpu/src/lib.rs:403-424) samples
// P_n = seed + n*(cos,sin) (|P_n - exact| < 1e-3 for n < 4096), n = 0,1,2.., and stops at the
// K1-th consecutive non-white sample, K1 = ceil(max_gap) + 1; the returned end point is the last
// white sample before that run (or the last in-bounds sample when the ray leaves through the
// left/top border; rays leaving right/bottom return length 0 — D7).  Hence a ray of length > 50
// has NO run of K1 non-white samples within steps [0, 49], so every window of K1 consecutive
// steps
//     W_j = [j*K1, (j+1)*K1 - 1],  j = 1 .. nw,  j*K1 <= 50
// holds a white sample.  (A window reaching beyond step 49 is only implied for rays that cannot
// leave the image before its end; it is used when the seed is farther than (j+1)*K1 + 1 from the
// left and top borders.)  A white pixel whose centre is at distance D from the seed can hold
// sample n only if |n - D| <= 0.7072 + 1e-3, and only for ray directions within asin(0.7082/D)
// of the direction of its centre.  Marking, per window, the rays each white pixel of its ring can
// serve and AND-ing the windows gives a superset of the rays that can matter; the others are
// never marched.  tests/test_lsd_filter.py checks the superset property on the CPU.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/smh_consts.h"

#if defined(__CUDACC__)
#define SMH_LF_HD __host__ __device__ __forceinline__
#else
#define SMH_LF_HD static inline
#endif

#define SMH_LSD_FILTER_MAX_WINDOWS 3
#define SMH_LSD_BITMAP_WORDS ((SMH_LSD_RAYS + 31) / 32)  /* 113 */

typedef struct {
    int nw;        /* number of windows (0: filter off, march every ray) */
    int k1;        /* consecutive non-white samples that stop a ray */
    int radius;    /* Chebyshev half-size of the pixel window to scan around floor(seed) */
    float r2_lo[SMH_LSD_FILTER_MAX_WINDOWS], r2_hi[SMH_LSD_FILTER_MAX_WINDOWS]; /* ring of window j: centre distance^2 */
    int border[SMH_LSD_FILTER_MAX_WINDOWS]; /* 0: always valid; else seed must be >= this far from left/top */
} smh_lsd_filter;

/* sqrt(SMH_LSD_MIN_LEN_SQ): a ray that matters has a white sample at step >= 50 (or leaves
 * through the left/top border after step 50) */
#define SMH_LSD_MIN_STEPS 50

SMH_LF_HD void smh_lsd_filter_init(smh_lsd_filter* f, int gap_limit) {
    f->nw = 0; f->k1 = gap_limit + 1; f->radius = 0;
    if (gap_limit < 1 || gap_limit > 48) return;  /* K1 > 49: no window fits; K1 < 2: use the plain march */
    const int k1 = gap_limit + 1;
    int nw = SMH_LSD_MIN_STEPS / k1;
    if (nw > SMH_LSD_FILTER_MAX_WINDOWS) nw = SMH_LSD_FILTER_MAX_WINDOWS;
    for (int j = 1; j <= nw; j++) {
        const float lo = (float)(j * k1) - 0.75f, hi = (float)((j + 1) * k1 - 1) + 0.75f;
        f->r2_lo[j - 1] = lo * lo; f->r2_hi[j - 1] = hi * hi;
        f->border[j - 1] = ((j + 1) * k1 - 1 > SMH_LSD_MIN_STEPS - 1) ? (j + 1) * k1 + 2 : 0;
    }
    f->nw = nw;
    f->radius = (nw + 1) * k1 + 1;
}

/* Is window j (0-based) usable for a seed at (px, py)? */
SMH_LF_HD bool smh_lsd_filter_window_on(const smh_lsd_filter* f, int j, float px, float py) {
    return f->border[j] == 0 || (px >= (float)f->border[j] && py >= (float)f->border[j]);
}

/* Ray-index interval [lo, hi] (inclusive, lo may be negative and hi >= 3600: take modulo 3600)
 * of the rays that can sample a pixel whose centre lies at (cx, cy) from the seed, d2 = cx^2+cy^2 >= 1. */
SMH_LF_HD void smh_lsd_filter_span(float cx, float cy, float d2, int* lo, int* hi) {
    const float per_rad = 572.9577951f;                  /* ray indices per radian (0.1 degree steps) */
    const float c = atan2f(cy, cx) * per_rad;            /* direction of the centre */
    const float x = 0.7082f / sqrtf(d2);                 /* sin of the half-angle of the pixel's disc */
    const float half = x * (1.0f + 0.25f * x * x) * per_rad + 0.01f;  /* >= asin(x) for x <= 0.75 */
    *lo = (int)floorf(c - half) - 1;                     /* one index of slack each side: atan2f / */
    *hi = (int)ceilf(c + half) + 1;                      /* angle-table rounding are << 0.1 degree */
}


/* The same for a QUARTER: the white pixels of one pixel row between columns a and b (centres at (cxa, cy) ..
 * (cxb, cy) from the seed, cxa <= cxb, at most 8 pixels apart).  Along a row the direction of a pixel centre is
 * monotonic in x (cy != 0) and the half-angle asin(0.7082 / D) is largest at the row's NEAREST point, so
 * [min(dir_a, dir_b) - half(dmin), max(dir_a, dir_b) + half(dmin)] covers every pixel between them.  A quarter
 * that comes closer than one pixel to the seed (this includes cy == 0 with the seed between a and b, where the
 * direction jumps by 180 degrees) marks every ray.  dmin2 / dmax2: squared distance range of the quarter. */
SMH_LF_HD void smh_lsd_filter_quarter(float cy, float cxa, float cxb, float* dmin2, float* dmax2, int* lo, int* hi) {
    const float cy2 = cy * cy, da = cxa * cxa + cy2, db = cxb * cxb + cy2;
    *dmax2 = da > db ? da : db;
    *dmin2 = (cxa < 0.0f && cxb > 0.0f) ? cy2 : (da < db ? da : db);
    if (*dmin2 < 1.0f) { *lo = 0; *hi = SMH_LSD_RAYS - 1; return; }
    const float per_rad = 572.9577951f;
    const float aa = atan2f(cy, cxa) * per_rad, ab = atan2f(cy, cxb) * per_rad;
    const float x = 0.7082f / sqrtf(*dmin2);
    const float half = x * (1.0f + 0.25f * x * x) * per_rad + 0.01f;
    *lo = (int)floorf((aa < ab ? aa : ab) - half) - 1;
    *hi = (int)ceilf((aa > ab ? aa : ab) + half) + 1;
}

/* OR the ray interval [lo, hi] (mod 3600) into a 113-word bitmap; ATOMIC_OR(word_ptr, mask) */
#define SMH_LSD_FILTER_SET_BITS(bits, lo_, hi_, ATOMIC_OR)                                        \
    do {                                                                                          \
        int a__ = (lo_), b__ = (hi_);                                                             \
        if (b__ - a__ >= SMH_LSD_RAYS - 1) { a__ = 0; b__ = SMH_LSD_RAYS - 1; }                   \
        while (a__ < 0) { a__ += SMH_LSD_RAYS; b__ += SMH_LSD_RAYS; }                              \
        while (a__ >= SMH_LSD_RAYS) { a__ -= SMH_LSD_RAYS; b__ -= SMH_LSD_RAYS; }                  \
        for (int part__ = 0; part__ < 2; part__++) {                                              \
            int s__ = part__ == 0 ? a__ : 0;                                                      \
            int e__ = part__ == 0 ? (b__ < SMH_LSD_RAYS ? b__ : SMH_LSD_RAYS - 1) : b__ - SMH_LSD_RAYS; \
            if (e__ < s__) continue;                                                              \
            for (int w__ = s__ >> 5; w__ <= (e__ >> 5); w__++) {                                  \
                const int lb__ = w__ == (s__ >> 5) ? (s__ & 31) : 0;                              \
                const int hb__ = w__ == (e__ >> 5) ? (e__ & 31) : 31;                             \
                const uint32_t m__ = (0xFFFFFFFFu >> (31 - hb__)) & (0xFFFFFFFFu << lb__);        \
                ATOMIC_OR(&(bits)[w__], m__);                                                     \
            }                                                                                     \
        }                                                                                         \
    } while (0)
