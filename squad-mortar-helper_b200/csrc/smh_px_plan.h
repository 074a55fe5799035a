// smh_px_plan.h — how the fused pixel pass (k_pixels2, kernels_px.cuh) cuts a map into work items; shared by the
// kernel, the host code that launches it, and (for the CPU-only partition test, tests/test_px_plan.py) csrc/hostcheck.cpp.
//
// Columns are padded coordinates X = x + pad; a STRIP is 128 columns.  The strips [qs0, qs1) of the rows [qr0, qr1)
// (multiples of 32 around the OCR quadrant's rows [bqh, 2 bqh), qr1 clipped to h) own every pixel of the quadrant: the
// QUAD items, bands of `band` rows (a multiple of 8) of one strip each.  The plain items own the rest: every strip of the
// rows above (na bands per strip) and below (nc), the other strips beside (nb).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SMH_PP_HD __host__ __device__ __forceinline__
#else
#define SMH_PP_HD static inline
#endif

typedef struct {
    uint32_t nstrips, h;
    uint32_t qr0, qr1, qs0, qs1;
    uint32_t na, nb, nc, nq, band;
} smh_px_plan;

SMH_PP_HD void smh_px_plan_make(smh_px_plan* p, uint32_t pad, uint32_t w, uint32_t h, uint32_t bqw, uint32_t bqh, uint32_t band) {
    p->nstrips = (pad + w + 127u) / 128u; p->h = h; p->band = band;
    p->qs0 = (bqw + pad) / 128u; p->qs1 = (2u * bqw + pad + 127u) / 128u;
    if (p->qs1 > p->nstrips) p->qs1 = p->nstrips;
    p->qr0 = bqh & ~31u; p->qr1 = (2u * bqh + 31u) & ~31u;
    if (p->qr1 > h) p->qr1 = h;
    if (p->qs0 >= p->qs1 || p->qr0 >= p->qr1) { p->qs0 = p->qs1 = 0; p->qr0 = p->qr1 = 0; }
    p->na = (p->qr0 + band - 1u) / band; p->nb = (p->qr1 - p->qr0 + band - 1u) / band; p->nc = (h - p->qr1 + band - 1u) / band;
    p->nq = p->nb;
}

// items per frame (the plain count without the frame's button test)
SMH_PP_HD uint32_t smh_px_plan_quad_items(const smh_px_plan* p) { return p->nq * (p->qs1 - p->qs0); }
SMH_PP_HD uint32_t smh_px_plan_plain_items(const smh_px_plan* p) {
    return (p->na + p->nc) * p->nstrips + p->nb * (p->nstrips - (p->qs1 - p->qs0));
}

// QUAD item `sub` of a frame -> its strip and rows [r0, r1)
SMH_PP_HD void smh_px_plan_quad_item(const smh_px_plan* p, uint32_t sub, uint32_t* strip, int* r0, int* r1) {
    const uint32_t nqs = p->qs1 - p->qs0, bq = sub / nqs;
    *strip = p->qs0 + (sub - bq * nqs);
    *r0 = (int)(p->qr0 + bq * p->band);
    *r1 = *r0 + (int)p->band < (int)p->qr1 ? *r0 + (int)p->band : (int)p->qr1;
}

// plain item `sub` of a frame: the rows above the QUAD items', then the strips beside their own, then the rows below
SMH_PP_HD void smh_px_plan_plain_item(const smh_px_plan* p, uint32_t sub, uint32_t* strip, int* r0, int* r1) {
    const uint32_t nqs = p->qs1 - p->qs0, above = p->na * p->nstrips, beside = p->nb * (p->nstrips - nqs);
    uint32_t lo, hi, bq;
    if (sub < above) { bq = sub / p->nstrips; *strip = sub - bq * p->nstrips; lo = 0u; hi = p->qr0; }
    else if (sub < above + beside) {
        sub -= above;
        const uint32_t per = p->nstrips - nqs, k = sub % per;
        bq = sub / per; *strip = k < p->qs0 ? k : k + nqs; lo = p->qr0; hi = p->qr1;
    } else {
        sub -= above + beside;
        bq = sub / p->nstrips; *strip = sub - bq * p->nstrips; lo = p->qr1; hi = p->h;
    }
    *r0 = (int)(lo + bq * p->band);
    *r1 = *r0 + (int)p->band < (int)hi ? *r0 + (int)p->band : (int)hi;
}
