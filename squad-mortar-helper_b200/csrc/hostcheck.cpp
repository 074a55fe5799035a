// hostcheck.cpp — CPU-only build of the shared per-pixel arithmetic (smh_pixel.h) so that the
// exhaustive 2^24-colour checks (exact predicate, integer fast-reject, luma, OCR classes) can
// run without a GPU.  Test helper only; never loaded by the product path.
// Build: g++ -O2 -ffp-contract=off -fPIC -shared hostcheck.cpp -o ../libsmh_hostcheck.so
#include "smh_pixel.h"
#include "smh_lsd_filter.h"
#include "smh_px_plan.h"

#include <string.h>

extern "C" {
// which: 0 marker (prefilter && exact), 1 luma8, 2 button-red, 3 ocr class, 4 exact only, 5 prefilter only
void smh_hostcheck_table(uint32_t which, uint8_t* out) {
    for (uint32_t c = 0; c < (1u << 24); c++) {
        const uint32_t r = c >> 16, g = (c >> 8) & 0xFFu, b = c & 0xFFu;
        uint8_t v;
        switch (which) {
            case 0: v = smh_is_marker(r, g, b); break;
            case 1: v = (uint8_t)smh_luma8(r, g, b); break;
            case 2: v = smh_button_red(r, g, b); break;
            case 3: v = (uint8_t)((smh_ocr_core(r, g, b) ? 1 : 0) | ((!smh_ocr_core(r, g, b) && smh_ocr_edge(r, g, b)) ? 2 : 0)); break;
            case 4: v = smh_is_marker_exact(r, g, b); break;
            default: v = smh_marker_prefilter(r, g, b); break;
        }
        out[c] = v;
    }
}

// The ray filter of k_lsd (smh_lsd_filter.h) evaluated on the CPU for one seed of an L8 mask:
// marked[k] = 1 for the rays k_lsd would march.  Returns the number of windows used (0: filter
// off, everything marked).  Same window scan, ring test, span and AND as the kernel, minus the
// parallel plumbing.
static inline void or_bits(uint32_t* p, uint32_t m) { *p |= m; }
int smh_hostcheck_lsd_filter(const uint8_t* img, uint32_t w, uint32_t h, float px, float py, int gap_limit, uint8_t* marked) {
    smh_lsd_filter F;
    smh_lsd_filter_init(&F, gap_limit);
    if (F.nw == 0) { memset(marked, 1, SMH_LSD_RAYS); return 0; }
    static uint32_t bits[SMH_LSD_FILTER_MAX_WINDOWS][SMH_LSD_BITMAP_WORDS + 3];
    memset(bits, 0, sizeof bits);
    const int fx = (int)px, fy = (int)py, R = F.radius;
    const int x0 = fx - R < 0 ? 0 : fx - R, x1 = fx + R > (int)w - 1 ? (int)w - 1 : fx + R;
    const int y0 = fy - R < 0 ? 0 : fy - R, y1 = fy + R > (int)h - 1 ? (int)h - 1 : fy + R;
    for (int y = y0; y <= y1; y++)
        for (int x = x0; x <= x1; x++) {
            if (img[(size_t)y * w + x] != 255) continue;
            const float cx = ((float)x + 0.5f) - px, cy = ((float)y + 0.5f) - py;
            const float d2 = cx * cx + cy * cy;
            if (d2 < F.r2_lo[0] || d2 > F.r2_hi[F.nw - 1]) continue;
            int lo, hi;
            smh_lsd_filter_span(cx, cy, d2, &lo, &hi);
            for (int j = 0; j < F.nw; j++)
                if (d2 >= F.r2_lo[j] && d2 <= F.r2_hi[j]) SMH_LSD_FILTER_SET_BITS(bits[j], lo, hi, or_bits);
        }
    for (int k = 0; k < SMH_LSD_RAYS; k++) {
        int on = 1;
        for (int j = 0; j < F.nw; j++)
            if (smh_lsd_filter_window_on(&F, j, px, py) && !((bits[j][k >> 5] >> (k & 31)) & 1u)) on = 0;
        marked[k] = (uint8_t)on;
    }
    return F.nw;
}

// The QUARTER form of the filter, as k_lsd2 marks it: white pixels are taken 8 at a time (columns 8q .. 8q+7 of a
// pixel row in the bitmap's padded coordinates X = x + pad), each quarter marks the union span of its first and
// last white pixel in every ring its distance range meets.
int smh_hostcheck_lsd_filter_quarters(const uint8_t* img, uint32_t w, uint32_t h, uint32_t pad, float px, float py, int gap_limit,
                                      uint8_t* marked) {
    smh_lsd_filter F;
    smh_lsd_filter_init(&F, gap_limit);
    if (F.nw == 0) { memset(marked, 1, SMH_LSD_RAYS); return 0; }
    static uint32_t bits[SMH_LSD_FILTER_MAX_WINDOWS][SMH_LSD_BITMAP_WORDS + 3];
    memset(bits, 0, sizeof bits);
    const int fx = (int)px, fy = (int)py, R = F.radius;
    const float hx = px - (float)fx, hy = py - (float)fy;
    const int x0 = fx - R < 0 ? 0 : fx - R, x1 = fx + R > (int)w - 1 ? (int)w - 1 : fx + R;
    const int y0 = fy - R < 0 ? 0 : fy - R, y1 = fy + R > (int)h - 1 ? (int)h - 1 : fy + R;
    for (int y = y0; y <= y1; y++) {
        const float cy = (float)(y - fy) + 0.5f - hy;
        for (int Xq = ((x0 + (int)pad) & ~7); Xq <= x1 + (int)pad; Xq += 8) {
            int a = -1, b = -1;
            for (int e = 0; e < 8; e++) {
                const int x = Xq + e - (int)pad;
                if (x < x0 || x > x1 || img[(size_t)y * w + x] != 255) continue;
                if (a < 0) a = x;
                b = x;
            }
            if (a < 0) continue;
            float dmin2, dmax2; int lo, hi;
            smh_lsd_filter_quarter(cy, (float)(a - fx) + 0.5f - hx, (float)(b - fx) + 0.5f - hx, &dmin2, &dmax2, &lo, &hi);
            for (int j = 0; j < F.nw; j++)
                if (dmax2 >= F.r2_lo[j] && dmin2 <= F.r2_hi[j]) SMH_LSD_FILTER_SET_BITS(bits[j], lo, hi, or_bits);
        }
    }
    for (int k = 0; k < SMH_LSD_RAYS; k++) {
        int on = 1;
        for (int j = 0; j < F.nw; j++)
            if (smh_lsd_filter_window_on(&F, j, px, py) && !((bits[j][k >> 5] >> (k & 31)) & 1u)) on = 0;
        marked[k] = (uint8_t)on;
    }
    return F.nw;
}

// The pixel pass's cut of a map into work items (smh_px_plan.h), item by item: cover[(row / 8) * nstrips + strip] += 1 for
// every 8-row group of every item, quad[...] = 1 where a QUAD item owns the group.  Returns the number of strips.
uint32_t smh_hostcheck_px_plan(uint32_t pad, uint32_t w, uint32_t h, uint32_t bqw, uint32_t bqh, uint32_t band, uint32_t* cover, uint8_t* quad,
                               uint32_t* counts) {
    smh_px_plan p;
    smh_px_plan_make(&p, pad, w, h, bqw, bqh, band);
    const uint32_t nq = smh_px_plan_quad_items(&p), np = smh_px_plan_plain_items(&p);
    for (uint32_t k = 0; k < nq + np; k++) {
        uint32_t strip; int r0, r1;
        if (k < nq) smh_px_plan_quad_item(&p, k, &strip, &r0, &r1);
        else smh_px_plan_plain_item(&p, k - nq, &strip, &r0, &r1);
        if (r0 < 0 || r1 <= r0 || r1 > (int)h || (r0 & 7) || strip >= p.nstrips) return 0xFFFFFFFFu;  // an item must own rows, 8-row aligned
        for (int g = r0 >> 3; g <= (r1 - 1) >> 3; g++) {
            cover[(uint32_t)g * p.nstrips + strip]++;
            if (k < nq) quad[(uint32_t)g * p.nstrips + strip] = 1;
        }
    }
    if (counts) { counts[0] = nq; counts[1] = np; counts[2] = p.qr0; counts[3] = p.qr1; counts[4] = p.qs0; counts[5] = p.qs1; }
    return p.nstrips;
}
}
