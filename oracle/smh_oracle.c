/* smh_oracle.c — CPU ORACLE (test infrastructure; see smh_oracle.h header comment).
 *
 * PARITY STATUS: pinned against the reference's own compiled CUDA kernels (oracle/_ref; see smh_oracle.h and
 * DESIGN.md section 2); the vision-cpu-only branches listed there are reader-verified.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math -fopenmp -fPIC -shared
 * All float arithmetic is IEEE binary32, no FMA contraction, evaluated in the order the
 * Rust source writes it.  `as uN` casts follow Rust semantics (truncate, saturate, NaN->0).
 */
#include "smh_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/smh_consts.h"
#include "../squad-mortar-helper_b200/csrc/smh_synth.h"

/* ---------------------------------------------------------------- Rust cast semantics */

static inline uint32_t f32_as_u32(float f) {
    if (!(f > 0.0f)) return 0; /* NaN, negative, zero */
    if (f >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)f;
}
static inline uint16_t f32_as_u16(float f) {
    if (!(f > 0.0f)) return 0;
    if (f >= 65535.0f) return 65535;
    return (uint16_t)f;
}
static inline uint8_t f32_as_u8(float f) {
    if (!(f > 0.0f)) return 0;
    if (f >= 255.0f) return 255;
    return (uint8_t)f;
}
static inline uint32_t abs_diff_u(uint32_t a, uint32_t b) { return a > b ? a - b : b - a; }

static int g_threads = 1;
void smh_oracle_set_threads(int n) {
    g_threads = n < 1 ? 1 : n;
#ifdef _OPENMP
    omp_set_num_threads(g_threads);
#endif
}
int smh_oracle_get_threads(void) { return g_threads; }

/* ---------------------------------------------------------------- ROI geometry */

/* RelativeBound::ScreenH(f).into_absolute — vision-common/src/screen.rs:61-66 */
static inline uint32_t rel_h(double frac, uint32_t H) { return (uint32_t)round(frac * (double)H); }

int smh_oracle_roi(uint32_t W, uint32_t H, uint32_t out[12]) {
    /* CLOSE_DEPLOYMENT_BUTTON_BOUNDS: x Right, y Bottom — consts/mod.rs:14-19, screen.rs:8-30,41-46 */
    uint32_t bw = rel_h(SMH_BTN_W_FRAC, H), bh = rel_h(SMH_BTN_H_FRAC, H);
    int64_t bx = (int64_t)W - rel_h(SMH_BTN_X_RIGHT_FRAC, H) - bw;
    int64_t by = (int64_t)H - rel_h(SMH_BTN_Y_BOTTOM_FRAC, H) - bh;
    /* MAP_BOUNDS: x Left, y Bottom — consts/mod.rs:7-12 */
    uint32_t w0 = rel_h(SMH_MAP_W_FRAC, H), h = rel_h(SMH_MAP_H_FRAC, H);
    uint32_t x0 = rel_h(SMH_MAP_X_FRAC, H);
    int64_t y = (int64_t)H - rel_h(SMH_MAP_Y_BOTTOM_FRAC, H) - h;
    /* "Map fills remaining space" — vision-cpu/src/lib.rs:139-141 */
    int64_t w = (int64_t)W - w0;
    int64_t x = (int64_t)W - x0 - w;
    if (bx < 0 || by < 0 || y < 0 || w <= 0 || x < 0 || h == 0 || bw == 0 || bh == 0) return -1;
    /* par_iter_pixels! bounds — util/src/image.rs:9-16 */
    if (bx + bw > W || by + bh > H) return -1;
    /* par_crop_into panics when x+w >= W or y+h >= H — util/src/image.rs:67-73 (D12) */
    if (x + w >= W || y + h >= H) return -1;
    uint32_t bqw = (uint32_t)w / 2, bqh = h / 2; /* vision-cpu/src/lib.rs:143-145 */
    if (bqw < SMH_OCR_RADIUS || bqh < SMH_OCR_RADIUS) return -1; /* `w - RADIUS` would underflow */
    out[0] = (uint32_t)bx; out[1] = (uint32_t)by; out[2] = bw; out[3] = bh;
    out[4] = (uint32_t)x; out[5] = (uint32_t)y; out[6] = (uint32_t)w; out[7] = h;
    out[8] = (uint32_t)x + bqw; out[9] = (uint32_t)y + bqh; out[10] = bqw; out[11] = bqh;
    return 0;
}

/* ---------------------------------------------------------------- colour helpers */

uint8_t smh_oracle_luma8(uint8_t r, uint8_t g, uint8_t b) {
    float l = SMH_LUMA_R * (float)r + SMH_LUMA_G * (float)g + SMH_LUMA_B * (float)b;
    return f32_as_u8(l);
}

void smh_oracle_hsv(uint8_t r8, uint8_t g8, uint8_t b8, uint16_t *ho, uint8_t *so, uint8_t *vo) {
    /* util/src/image.rs:159-187 */
    float r = (float)r8 / 255.0f, g = (float)g8 / 255.0f, b = (float)b8 / 255.0f;
    float max = fmaxf(r, fmaxf(g, b));
    float min = fminf(r, fminf(g, b));
    float delta = max - min;
    float h;
    if (max == min) h = 0.0f;
    else if (max == r) h = 60.0f * fmodf((g - b) / delta, 6.0f);
    else if (max == g) h = 60.0f * (((b - r) / delta) + 2.0f);
    else h = 60.0f * (((r - g) / delta) + 4.0f);
    float s = 100.0f * delta / max; /* (100*delta)/max; 0/0 = NaN -> 0 */
    float v = 100.0f * max;
    float hm = fmodf(h, 360.0f);
    if (hm < 0.0f) hm = hm + 360.0f;
    *ho = f32_as_u16(hm);
    *so = f32_as_u8(s);
    *vo = f32_as_u8(v);
}

/* vision-common/src/markers/mod.rs:17-19 */
static inline int saturation_ok(uint8_t s, uint8_t ms) {
    if (abs_diff_u(ms, s) <= SMH_SAT_TOL) return 1;
    int16_t d = (int16_t)((int16_t)s - ((int16_t)ms - SMH_ARC_SAT));
    if (d < 0) d = (int16_t)-d;
    return (uint8_t)d <= SMH_SAT_TOL;
}

int smh_oracle_is_marker(uint8_t r, uint8_t g, uint8_t b) {
    /* vision-common/src/markers/mod.rs:40-54 */
    static const uint16_t MH[3] = {SMH_ALPHA_H, SMH_BRAVO_H, SMH_CHARLIE_H};
    static const uint8_t MS[3] = {SMH_ALPHA_S, SMH_BRAVO_S, SMH_CHARLIE_S};
    static const uint8_t MV[3] = {SMH_ALPHA_V, SMH_BRAVO_V, SMH_CHARLIE_V};
    uint16_t h; uint8_t s, v;
    smh_oracle_hsv(r, g, b, &h, &s, &v);
    if (s < SMH_MIN_SAT) return 0;
    for (int t = 0; t < 3; t++) {
        if (abs_diff_u(MH[t], h) <= SMH_HUE_TOL && saturation_ok(s, MS[t]) &&
            abs_diff_u(MV[t], v) <= SMH_VIB_TOL)
            return 1;
    }
    return 0;
}

void smh_oracle_table_marker(uint8_t *out) {
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < (1 << 24); c++)
        out[c] = (uint8_t)smh_oracle_is_marker((uint8_t)(c >> 16), (uint8_t)(c >> 8), (uint8_t)c);
}
void smh_oracle_table_luma(uint8_t *out) {
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < (1 << 24); c++)
        out[c] = smh_oracle_luma8((uint8_t)(c >> 16), (uint8_t)(c >> 8), (uint8_t)c);
}
void smh_oracle_table_hsv(uint16_t *h, uint8_t *s, uint8_t *v) {
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < (1 << 24); c++)
        smh_oracle_hsv((uint8_t)(c >> 16), (uint8_t)(c >> 8), (uint8_t)c, &h[c], &s[c], &v[c]);
}

void smh_oracle_ray_table(float *dxdy) {
    /* vision-cpu/src/lib.rs:437 `((theta as f32) / 10.0).to_radians()`; f32::to_radians is
     * `self * (PI_f32 / 180.0_f32)`; :398-399 `theta.cos()`, `theta.sin()` -> libm cosf/sinf */
    const float rads_per_deg = 3.14159265358979323846f / 180.0f;
    for (uint32_t k = 0; k < SMH_LSD_RAYS; k++) {
        float theta = ((float)k / 10.0f) * rads_per_deg;
        dxdy[2 * k] = cosf(theta);
        dxdy[2 * k + 1] = sinf(theta);
    }
}

/* ---------------------------------------------------------------- backend state */

struct smh_oracle {
    uint8_t *frame; /* BGRA, W*H*4 */
    uint32_t W, H;
    uint32_t roi[12];
    uint8_t *cropped_map;  /* RGB8 w*h */
    uint8_t *cropped_brq;  /* RGB8 bqw*bqh */
    uint8_t *ocr_out;      /* L8 bqw*bqh */
    uint8_t *scales;       /* L8 bqw*bqh */
    uint8_t *lsd_image;    /* L8 w*h */
    uint8_t *raw_mask;     /* L8 w*h (oracle-only view of the mask before dilation) */
    uint8_t *ui_map;       /* RGBA w*h, owned by the oracle, refreshed by crop_to_map */
    uint16_t *dist;        /* scratch for the distance transform */
    /* disabled template-match stage (A20): UnfilteredMarkers layout, n * size*size RGBA */
    uint8_t *markers; uint32_t n_markers, marker_size;
};

smh_oracle *smh_oracle_new(void) { return (smh_oracle *)calloc(1, sizeof(smh_oracle)); }

static void free_buffers(smh_oracle *o) {
    free(o->cropped_map); free(o->cropped_brq); free(o->ocr_out); free(o->scales);
    free(o->lsd_image); free(o->raw_mask); free(o->ui_map); free(o->dist);
    o->cropped_map = o->cropped_brq = o->ocr_out = o->scales = o->lsd_image = o->raw_mask = o->ui_map = NULL;
    o->dist = NULL;
}

void smh_oracle_free(smh_oracle *o) {
    if (!o) return;
    free_buffers(o);
    free(o->frame);
    free(o->markers);
    free(o);
}

int smh_oracle_load_frame(smh_oracle *o, const uint8_t *bgra, uint32_t W, uint32_t H) {
    /* vision-cpu/src/lib.rs:69-93: buffers are re-created (zeroed) only when the frame
     * dimensions change; otherwise they keep the previous frame's content. */
    if (!o || !bgra || W == 0 || H == 0) return -2;
    if (o->W != W || o->H != H || !o->frame) {
        uint32_t roi[12];
        if (smh_oracle_roi(W, H, roi) != 0) return -1;
        free_buffers(o);
        free(o->frame);
        o->frame = (uint8_t *)malloc((size_t)W * H * 4);
        memcpy(o->roi, roi, sizeof roi);
        size_t n = (size_t)roi[6] * roi[7], nq = (size_t)roi[10] * roi[11];
        o->cropped_map = (uint8_t *)calloc(n, 3);
        o->cropped_brq = (uint8_t *)calloc(nq, 3);
        o->ocr_out = (uint8_t *)calloc(nq, 1);
        o->scales = (uint8_t *)calloc(nq, 1);
        o->lsd_image = (uint8_t *)calloc(n, 1);
        o->raw_mask = (uint8_t *)calloc(n, 1);
        o->ui_map = (uint8_t *)calloc(n, 4);
        o->dist = (uint16_t *)calloc(n, sizeof(uint16_t));
        o->W = W; o->H = H;
    }
    memcpy(o->frame, bgra, (size_t)W * H * 4);
    return 0;
}

int smh_oracle_crop_to_map(smh_oracle *o, int grayscale, int *map_open, uint32_t xywh[4],
                           const uint8_t **ui_map_rgba, uint32_t *red_pixels) {
    if (!o || !o->frame) return -2;
    const uint32_t W = o->W;
    const uint8_t *f = o->frame;
    /* button test — vision-cpu/src/lib.rs:115-135 */
    const uint32_t bx = o->roi[0], by = o->roi[1], bw = o->roi[2], bh = o->roi[3];
    static const int16_t C[3] = {SMH_BTN_COLOR_R, SMH_BTN_COLOR_G, SMH_BTN_COLOR_B};
    uint32_t red = 0;
#pragma omp parallel for reduction(+ : red) schedule(static)
    for (uint32_t j = 0; j < bh; j++)
        for (uint32_t i = 0; i < bw; i++) {
            const uint8_t *p = f + ((size_t)(by + j) * W + (bx + i)) * 4;
            const uint8_t rgb[3] = {p[2], p[1], p[0]};
            int ok = 1;
            for (int c = 0; c < 3; c++) {
                int16_t d = (int16_t)(C[c] - (int16_t)rgb[c]);
                if (d < 0) d = (int16_t)-d;
                if ((uint16_t)d > SMH_BTN_TOLERANCE) { ok = 0; break; }
            }
            red += (uint32_t)ok;
        }
    if (red_pixels) *red_pixels = red;
    float ratio = (float)red / (float)(bw * bh);
    if (ratio < SMH_BTN_RED_RATIO_THRESHOLD) {
        if (map_open) *map_open = 0;
        return 0; /* Ok(None) */
    }
    /* crops — vision-cpu/src/lib.rs:137-170 */
    const uint32_t x = o->roi[4], y = o->roi[5], w = o->roi[6], h = o->roi[7];
    const uint32_t qx = o->roi[8], qy = o->roi[9], qw = o->roi[10], qh = o->roi[11];
#pragma omp parallel for schedule(static)
    for (uint32_t j = 0; j < h; j++)
        for (uint32_t i = 0; i < w; i++) {
            const uint8_t *p = f + ((size_t)(y + j) * W + (x + i)) * 4;
            uint8_t *u = o->ui_map + ((size_t)j * w + i) * 4;
            uint8_t *c = o->cropped_map + ((size_t)j * w + i) * 3;
            if (grayscale) {
                uint8_t l = smh_oracle_luma8(p[2], p[1], p[0]);
                u[0] = l; u[1] = l; u[2] = l; u[3] = 255;
            } else {
                u[0] = p[2]; u[1] = p[1]; u[2] = p[0]; u[3] = 255;
            }
            c[0] = p[2]; c[1] = p[1]; c[2] = p[0];
        }
#pragma omp parallel for schedule(static)
    for (uint32_t j = 0; j < qh; j++)
        for (uint32_t i = 0; i < qw; i++) {
            const uint8_t *p = f + ((size_t)(qy + j) * W + (qx + i)) * 4;
            uint8_t *c = o->cropped_brq + ((size_t)j * qw + i) * 3;
            c[0] = p[2]; c[1] = p[1]; c[2] = p[0];
        }
    if (map_open) *map_open = 1;
    if (xywh) { xywh[0] = x; xywh[1] = y; xywh[2] = w; xywh[3] = h; }
    if (ui_map_rgba) *ui_map_rgba = o->ui_map;
    return 0;
}

/* vision-cpu/src/lib.rs:39-53 */
static inline uint8_t ocr_brightness(const uint8_t *p) {
    uint8_t m = p[0] < p[1] ? p[0] : p[1];
    return m < p[2] ? m : p[2];
}
static inline uint16_t ocr_monochromaticy(const uint8_t *p) {
    uint16_t diff = 0;
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) diff = (uint16_t)(diff + abs_diff_u(p[a], p[b]));
    return diff;
}

int smh_oracle_ocr_preprocess(smh_oracle *o, const uint8_t **out, size_t *len) {
    /* vision-cpu/src/lib.rs:173-231 */
    if (!o || !o->frame) return -2;
    const uint32_t w = o->roi[10], h = o->roi[11];
    const uint8_t *q = o->cropped_brq;
    const uint32_t R = SMH_OCR_RADIUS;
#pragma omp parallel for schedule(static)
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) {
            const uint8_t *p = q + ((size_t)y * w + x) * 3;
            int keep = 0;
            uint16_t diff = ocr_monochromaticy(p);
            if (diff <= SMH_OCR_MONO && ocr_brightness(p) >= SMH_OCR_BRIGHT) {
                keep = 1;
            } else if (diff <= SMH_OCR_MONO_EDGE && ocr_brightness(p) >= SMH_OCR_BRIGHT_EDGE) {
                /* window: x.saturating_sub(R) ..= min(x+R, w-R) (lib.rs:201-205, D10) */
                uint32_t x_lo = x >= R ? x - R : 0, x_hi = (x + R < w - R) ? x + R : w - R;
                uint32_t y_lo = y >= R ? y - R : 0, y_hi = (y + R < h - R) ? y + R : h - R;
                for (uint32_t xx = x_lo; xx <= x_hi && !keep; xx++)
                    for (uint32_t yy = y_lo; yy <= y_hi; yy++) {
                        const uint8_t *n = q + ((size_t)yy * w + xx) * 3;
                        if (ocr_brightness(n) < SMH_OCR_BRIGHT) continue;
                        if (ocr_monochromaticy(n) <= SMH_OCR_MONO) { keep = 1; break; }
                    }
            }
            o->ocr_out[(size_t)y * w + x] =
                keep ? (uint8_t)(255 - smh_oracle_luma8(p[0], p[1], p[2])) : 255;
        }
    if (out) *out = o->ocr_out;
    if (len) *len = (size_t)w * h;
    return 0;
}

int smh_oracle_find_scales_preprocess(smh_oracle *o, uint32_t start_y, const uint8_t **out,
                                      uint32_t wh[2]) {
    /* vision-cpu/src/lib.rs:233-251.  Release-build behaviour for start_y >= h: `h - start_y`
     * wraps, par_iter_pixels' bound check `y + h' > height` wraps back to `h > h` = false and
     * the row range start_y..h is empty => no-op (D9). */
    if (!o || !o->frame) return -2;
    const uint32_t w = o->roi[10], h = o->roi[11];
    const uint8_t *q = o->cropped_brq;
    if (start_y < h) {
#pragma omp parallel for schedule(static)
        for (uint32_t y = start_y; y < h; y++)
            for (uint32_t x = 0; x < w; x++) {
                const uint8_t *p = q + ((size_t)y * w + x) * 3;
                o->scales[(size_t)y * w + x] = smh_oracle_luma8(p[0], p[1], p[2]) != 0 ? 255 : 0;
            }
    }
    if (out) *out = o->scales;
    if (wh) { wh[0] = w; wh[1] = h; }
    return 0;
}

int smh_oracle_isolate_map_markers(smh_oracle *o) {
    /* vision-cpu/src/lib.rs:253-280 */
    if (!o || !o->frame) return -2;
    const size_t n = (size_t)o->roi[6] * o->roi[7];
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        uint8_t *p = o->cropped_map + i * 3;
        if (!smh_oracle_is_marker(p[0], p[1], p[2])) { p[0] = 0; p[1] = 0; p[2] = 0; }
    }
    return 0;
}

/* imageproc 0.22.0 morphology::dilate_mut(image, Norm::L1, k):
 *   distance_transform_mut(image, L1)  — two-pass chamfer over the 4-neighbourhood,
 *     foreground (non-zero) -> 0, background -> 255, candidate+1 compared in u16;
 *   then every pixel <= k becomes 255, else 0. */
static void dilate_l1(uint8_t *img, uint16_t *d, uint32_t w, uint32_t h, uint8_t k) {
    const size_t n = (size_t)w * h;
    for (size_t i = 0; i < n; i++) d[i] = img[i] > 0 ? 0 : 255;
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) {
            uint16_t cur = d[(size_t)y * w + x];
            if (x > 0) { uint16_t c = (uint16_t)(d[(size_t)y * w + x - 1] + 1); if (c < cur) cur = c; }
            if (y > 0) { uint16_t c = (uint16_t)(d[(size_t)(y - 1) * w + x] + 1); if (c < cur) cur = c; }
            d[(size_t)y * w + x] = cur;
        }
    for (uint32_t y = h; y-- > 0;)
        for (uint32_t x = w; x-- > 0;) {
            uint16_t cur = d[(size_t)y * w + x];
            if (x + 1 < w) { uint16_t c = (uint16_t)(d[(size_t)y * w + x + 1] + 1); if (c < cur) cur = c; }
            if (y + 1 < h) { uint16_t c = (uint16_t)(d[(size_t)(y + 1) * w + x] + 1); if (c < cur) cur = c; }
            d[(size_t)y * w + x] = cur;
        }
    for (size_t i = 0; i < n; i++) img[i] = d[i] <= k ? 255 : 0;
}

int smh_oracle_mask_marker_lines(smh_oracle *o) {
    /* vision-cpu/src/lib.rs:357-375 */
    if (!o || !o->frame) return -2;
    const uint32_t w = o->roi[6], h = o->roi[7];
    const size_t n = (size_t)w * h;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        const uint8_t *p = o->cropped_map + i * 3;
        o->lsd_image[i] = smh_oracle_is_marker(p[0], p[1], p[2]) ? 255 : 0;
    }
    memcpy(o->raw_mask, o->lsd_image, n);
    dilate_l1(o->lsd_image, o->dist, w, h, 1);
    return 0;
}

/* ---------------------------------------------------------------- line search */

static _Thread_local uint64_t t_seeds, t_steps;
void smh_oracle_lsd_stats(uint64_t *seeds, uint64_t *ray_steps) {
    if (seeds) *seeds = t_seeds;
    if (ray_steps) *ray_steps = t_steps;
}

/* find_line_in_image — vision-cpu/src/lib.rs:388-432 */
static inline uint32_t march(const uint8_t *img, uint32_t w, uint32_t h, float px, float py,
                             float max_gap, float dx, float dy, float end[2]) {
    float x = px, y = py;
    float gap_n = 0.0f, gap_x = 0.0f, gap_y = 0.0f;
    float xo = 0.0f, yo = 0.0f;
    const float wf = (float)w, hf = (float)h;
    uint32_t steps = 0;
    while (x >= 0.0f && y >= 0.0f && x < wf && y < hf) {
        steps++;
        if (img[(size_t)(uint32_t)y * w + (uint32_t)x] == 255) {
            gap_n = 0.0f; gap_x = 0.0f; gap_y = 0.0f;
        } else if (gap_n >= max_gap) {
            x = gap_x; y = gap_y;
            break;
        } else if (gap_n == 0.0f) {
            gap_n = 1.0f; gap_x = x; gap_y = y;
        } else {
            gap_n += 1.0f;
        }
        xo += dx; yo += dy;
        x = xo + px; y = yo + py;
    }
    /* image.get_pixel_checked(x as u32, y as u32) == Some(Luma([0])) — lib.rs:426-429,
     * util/src/image.rs:113-124 (D7) */
    uint32_t xi = f32_as_u32(x), yi = f32_as_u32(y);
    if (xi < w && yi < h && img[(size_t)yi * w + xi] == 0) {
        end[0] = x - dx; end[1] = y - dy;
    } else {
        end[0] = px; end[1] = py;
    }
    return steps;
}

static const float *ray_table(void) {
    static float tab[2 * SMH_LSD_RAYS];
    static int ready = 0;
#pragma omp critical(smh_ray_table)
    {
        if (!ready) { smh_oracle_ray_table(tab); ready = 1; }
    }
    return tab;
}

int smh_oracle_find_longest_line_in(const uint8_t *img, uint32_t w, uint32_t h, float px, float py,
                                    float max_gap, float line[4], float *len_sq) {
    /* vision-cpu/src/lib.rs:434-448: map over theta in 0..3600, reduce(Default, keep a only if
     * a_len > b_len) => the LAST index among equal maxima wins; the Default identity only ever
     * sits on the left of the fold, so an all-zero search returns ray 3599's (pt,pt). */
    const float *tab = ray_table();
    float best_len = -1.0f; /* any len >= 0 replaces it at k=0, like (identity, item) -> item */
    float best_end[2] = {px, py};
    uint64_t steps_total = 0;
    if (g_threads > 1) {
        int best_k = -1;
#pragma omp parallel reduction(+ : steps_total)
        {
            float l_len = -1.0f, l_end[2] = {px, py};
            int l_k = -1;
#pragma omp for schedule(static) nowait
            for (int k = 0; k < SMH_LSD_RAYS; k++) {
                float e[2];
                steps_total += march(img, w, h, px, py, max_gap, tab[2 * k], tab[2 * k + 1], e);
                float len = (px - e[0]) * (px - e[0]) + (py - e[1]) * (py - e[1]);
                if (!(l_len > len)) { l_len = len; l_end[0] = e[0]; l_end[1] = e[1]; l_k = k; }
            }
            /* merge thread-local winners by the lexicographic (len, k) maximum: identical to
             * the sequential right-biased fold whatever the schedule */
#pragma omp critical(smh_reduce)
            {
                if (l_k >= 0 && (l_len > best_len || (l_len == best_len && l_k > best_k))) {
                    best_len = l_len; best_end[0] = l_end[0]; best_end[1] = l_end[1]; best_k = l_k;
                }
            }
        }
    } else {
        for (int k = 0; k < SMH_LSD_RAYS; k++) {
            float e[2];
            steps_total += march(img, w, h, px, py, max_gap, tab[2 * k], tab[2 * k + 1], e);
            /* distance_sqr — util/src/geometry.rs:62-68: (p0-p1)^2 summed, unfused */
            float len = (px - e[0]) * (px - e[0]) + (py - e[1]) * (py - e[1]);
            if (!(best_len > len)) { best_len = len; best_end[0] = e[0]; best_end[1] = e[1]; }
        }
    }
    t_steps += steps_total;
    line[0] = px; line[1] = py; line[2] = best_end[0]; line[3] = best_end[1];
    *len_sq = best_len;
    return 0;
}

/* Test helper: squared length of every one of the 3600 rays of one seed (same march as above). */
int smh_oracle_ray_lengths_in(const uint8_t *img, uint32_t w, uint32_t h, float px, float py, float max_gap,
                              float *len_sq) {
    const float *tab = ray_table();
#pragma omp parallel for schedule(static)
    for (int k = 0; k < SMH_LSD_RAYS; k++) {
        float e[2];
        march(img, w, h, px, py, max_gap, tab[2 * k], tab[2 * k + 1], e);
        len_sq[k] = (px - e[0]) * (px - e[0]) + (py - e[1]) * (py - e[1]);
    }
    return 0;
}

int smh_oracle_find_longest_line(smh_oracle *o, float x, float y, float max_gap, float line[4],
                                 float *len_sq) {
    if (!o || !o->frame) return -2;
    return smh_oracle_find_longest_line_in(o->lsd_image, o->roi[6], o->roi[7], x, y, max_gap, line, len_sq);
}

/* get_centre — vision-common/src/lsd.rs:5-44.  The reference indexes with unchecked
 * `unsafe_get_pixel`; an index == w or == h can only come from f32 rounding at the right /
 * bottom border (UB in the reference, D8) — clamp there. */
static inline uint8_t px_at(const uint8_t *img, uint32_t w, uint32_t h, float fx, float fy) {
    uint32_t x = f32_as_u32(fx), y = f32_as_u32(fy);
    if (x >= w) x = w - 1;
    if (y >= h) y = h - 1;
    return img[(size_t)y * w + x];
}
static void get_centre(const uint8_t *img, uint32_t w, uint32_t h, float ptx, float pty, float out[2]) {
    const float MAX_DIST = SMH_LSD_CENTRE_MAX_DIST;
    float left = ptx;
    while (left > 0.0f && fabsf(left - ptx) < MAX_DIST && px_at(img, w, h, left, pty) == 255) left -= 1.0f;
    float right = ptx;
    while (right < (float)(w - 1) && fabsf(right - ptx) < MAX_DIST && px_at(img, w, h, right, pty) == 255) right += 1.0f;
    float up = pty;
    while (up > 0.0f && fabsf(up - pty) < MAX_DIST && px_at(img, w, h, ptx, up) == 255) up -= 1.0f;
    float down = pty;
    while (down < (float)(h - 1) && fabsf(down - pty) < MAX_DIST && px_at(img, w, h, ptx, down) == 255) down += 1.0f;
    out[0] = (left + right) / 2.0f;
    out[1] = (up + down) / 2.0f;
}

/* nearest_point_on_line — vision-common/src/lsd.rs:46-58 (infinite line, D14) */
/* Test helper: get_centre on a caller-supplied image (the seed positions find_lines searches from). */
void smh_oracle_get_centre_in(const uint8_t *img, uint32_t w, uint32_t h, float x, float y, float out[2]) {
    get_centre(img, w, h, x, y, out);
}

static inline void nearest_point_on_line(float ptx, float pty, const float l[4], float out[2]) {
    float dx = l[2] - l[0], dy = l[3] - l[1];
    if (dx == 0.0f && dy == 0.0f) { out[0] = l[0]; out[1] = l[1]; return; }
    float u = ((ptx - l[0]) * dx + (pty - l[1]) * dy) / (dx * dx + dy * dy);
    out[0] = l[0] + u * dx;
    out[1] = l[1] + u * dy;
}

static int default_longest_cb(void *user, const uint8_t *img, uint32_t w, uint32_t h, float x, float y,
                             float max_gap, float line[4], float *len_sq) {
    (void)user;
    return smh_oracle_find_longest_line_in(img, w, h, x, y, max_gap, line, len_sq);
}

int smh_oracle_find_lines_cb(const uint8_t *img, uint32_t w, uint32_t h, uint32_t max_gap_u,
                             smh_oracle_longest_fn cb, void *user, float lines[32][4], uint32_t *n_out) {
    /* vision-common/src/lsd.rs:60-105 with N = 32 (vision-common/src/lib.rs:58); `cb` plays the
     * backend's find_longest_line closure (lsd.rs:63, vision-gpu/src/lib.rs:481-486) */
    const float max_gap = (float)max_gap_u;
    uint32_t n = 0;
    t_seeds = 0; t_steps = 0;
    for (uint32_t yi = 0; yi < h && n < SMH_LSD_MAX_LINES; yi++) {
        for (uint32_t xi = 0; xi < w; xi++) {
            if (img[(size_t)yi * w + xi] != 255) continue;
            const float x = (float)xi, y = (float)yi;
            int suppressed = 0;
            for (uint32_t l = 0; l < n; l++) {
                float np[2];
                nearest_point_on_line(x, y, lines[l], np);
                float ddx = x - np[0], ddy = y - np[1];
                if (ddx * ddx + ddy * ddy < SMH_LSD_SUPPRESS_DIST_SQ) { suppressed = 1; break; }
            }
            if (suppressed) continue;
            float c[2];
            get_centre(img, w, h, x, y, c);
            float longest[4], len;
            t_seeds++;
            int rc = cb(user, img, w, h, c[0], c[1], max_gap, longest, &len);
            if (rc) { *n_out = n; return rc; }
            if (len > SMH_LSD_MIN_LEN_SQ) {
                float c1[2];
                get_centre(img, w, h, longest[2], longest[3], c1);
                lines[n][0] = longest[0]; lines[n][1] = longest[1];
                lines[n][2] = c1[0]; lines[n][3] = c1[1];
                n++;
                if (n == SMH_LSD_MAX_LINES) break; /* break 'row */
            }
        }
    }
    *n_out = n;
    return 0;
}

int smh_oracle_find_lines_in(const uint8_t *img, uint32_t w, uint32_t h, uint32_t max_gap_u,
                             float lines[32][4], uint32_t *n_out) {
    return smh_oracle_find_lines_cb(img, w, h, max_gap_u, default_longest_cb, NULL, lines, n_out);
}

int smh_oracle_find_marker_lines(smh_oracle *o, uint32_t max_gap, float lines[32][4], uint32_t *n) {
    if (!o || !o->frame) return -2;
    return smh_oracle_find_lines_in(o->lsd_image, o->roi[6], o->roi[7], max_gap, lines, n);
}

/* ---------------------------------------------------------------- debug views / getters */

int smh_oracle_debug_view(smh_oracle *o, uint8_t which, uint8_t *rgba, uint32_t wh[2]) {
    /* vision-cpu/src/lib.rs:451-460; image's ConvertBuffer: Luma -> (l,l,l,255), Rgb -> (r,g,b,255) */
    if (!o || !o->frame) return -2;
    const uint8_t *src; uint32_t w, h; int ch;
    switch (which) {
        case 0: return 1;
        case 1: src = o->ocr_out; w = o->roi[10]; h = o->roi[11]; ch = 1; break;
        case 2: src = o->scales; w = o->roi[10]; h = o->roi[11]; ch = 1; break;
        case 3: src = o->cropped_map; w = o->roi[6]; h = o->roi[7]; ch = 3; break;
        case 4: src = o->lsd_image; w = o->roi[6]; h = o->roi[7]; ch = 1; break;
        case 5: src = o->cropped_brq; w = o->roi[10]; h = o->roi[11]; ch = 3; break;
        default: return -3;
    }
    if (wh) { wh[0] = w; wh[1] = h; }
    if (rgba) {
        const size_t n = (size_t)w * h;
        for (size_t i = 0; i < n; i++) {
            if (ch == 1) { rgba[4 * i] = rgba[4 * i + 1] = rgba[4 * i + 2] = src[i]; }
            else { rgba[4 * i] = src[3 * i]; rgba[4 * i + 1] = src[3 * i + 1]; rgba[4 * i + 2] = src[3 * i + 2]; }
            rgba[4 * i + 3] = 255;
        }
    }
    return 0;
}

/* ---------------------------------------------------------------- marker icon template match (A20 / N4)
 * The stage is commented out everywhere in the reference (F2); this restates the commented code. */

int smh_oracle_load_map_markers(smh_oracle *o, const uint8_t *rgba, uint32_t n, uint32_t size) {
    /* vision-cpu/src/lib.rs:101-108; the templates themselves come from
     * vision-common/src/markers/mod.rs:111-119 (host-side `image` crate work: TGA decode, Gaussian
     * resize, colour isolation, corner mask) and are passed in as data. */
    if (!o || !rgba || n == 0 || size == 0) return -2;
    free(o->markers);
    o->markers = (uint8_t *)malloc((size_t)n * size * size * 4);
    memcpy(o->markers, rgba, (size_t)n * size * size * 4);
    o->n_markers = n; o->marker_size = size;
    return 0;
}

int smh_oracle_filter_map_marker_icons(smh_oracle *o, int *found, uint32_t xy[2], uint32_t *sad_out) {
    /* vision-cpu/src/lib.rs:282-355 on cropped_map AFTER isolate_map_markers, with the candidate list
     * of lib.rs:270-277.  The reference's candidate order is whatever order its rayon workers' fetch_add
     * produced, and min_by_key keeps the first minimum of that order, so ties are nondeterministic
     * there; D17: candidates in raster order, markers in order, first minimum wins. */
    if (!o || !o->frame || !o->markers) return -2;
    const uint32_t w = o->roi[6], h = o->roi[7], size = o->marker_size;
    if (found) *found = 0;
    /* `if w >= map_marker_size && h >= map_marker_size` (src/vision/mod.rs:104); the u32 bounds
     * `w - size - 1` of lib.rs:272-273 must not wrap for the candidate rule to mean anything */
    if (w < size + 2 || h < size + 2) return 0;
    uint8_t *map = o->cropped_map;
    uint64_t best = UINT64_MAX;
#pragma omp parallel for schedule(dynamic, 8) reduction(min : best)
    for (uint32_t y = size; y < h - size - 1; y++)
        for (uint32_t x = size; x < w - size - 1; x++) {
            const uint8_t *c = map + ((size_t)y * w + x) * 3;
            /* a candidate is a pixel that survived isolate (lib.rs:269-277) */
            if (!smh_oracle_is_marker(c[0], c[1], c[2])) continue;
            for (uint32_t m = 0; m < o->n_markers; m++) {
                const uint8_t *tpl = o->markers + (size_t)m * size * size * 4;
                uint32_t sad = 0;
                for (uint32_t my = 0; my < size; my++)
                    for (uint32_t mx = 0; mx < size; mx++) {
                        const uint8_t *t = tpl + ((size_t)my * size + mx) * 4;
                        if (t[3] == 0) continue; /* FilteredMarkers drops alpha == 0 (markers/mod.rs:135) */
                        const uint8_t *q = map + ((size_t)(y + my) * w + (x + mx)) * 3;
                        uint32_t ad = 0;
                        for (int k = 0; k < 3; k++) ad += (uint32_t)(q[k] > t[k] ? q[k] - t[k] : t[k] - q[k]);
                        /* ((ad as f32) * (alpha as f32 / 255.0)) as u32 — lib.rs:323 */
                        const float a = (float)t[3] / 255.0f;
                        const float v = (float)ad * a;
                        sad += (uint32_t)v;
                    }
                const uint64_t key = ((uint64_t)sad << 32) | ((uint64_t)y << 16) | x;
                if (key < best) best = key;
            }
        }
    if (best == UINT64_MAX) return 0;
    const uint32_t x = (uint32_t)(best & 0xFFFFu), y = (uint32_t)((best >> 16) & 0xFFFFu);
    if (found) *found = 1;
    if (xy) { xy[0] = x; xy[1] = y; }
    if (sad_out) *sad_out = (uint32_t)(best >> 32);
    /* erase the icon — lib.rs:335-340 */
    const uint32_t ex = x + size < w - 1 ? x + size : w - 1, ey = y + size < h - 1 ? y + size : h - 1;
    for (uint32_t cx = x; cx < ex; cx++)
        for (uint32_t cy = y; cy < ey; cy++) { uint8_t *q = map + ((size_t)cy * w + cx) * 3; q[0] = q[1] = q[2] = 0; }
    /* leave a 4x4 square where the icon points — lib.rs:344-351 */
    const uint32_t px = x + size / 2, py = y + (uint32_t)roundf((float)size * SMH_MAP_MARKER_POI_LOCATION);
    for (uint32_t cx = px >= 2 ? px - 2 : 0; cx < px + 2; cx++)
        for (uint32_t cy = py >= 2 ? py - 2 : 0; cy < py + 2; cy++) {
            if (cx >= w || cy >= h) continue; /* put_pixel_fast is unchecked; unreachable for candidates of lib.rs:270-273 */
            uint8_t *q = map + ((size_t)cy * w + cx) * 3; q[0] = 0; q[1] = 255; q[2] = 0;
        }
    return 0;
}

#define GETTER(name, field, wi, hi)                                     \
    const uint8_t *smh_oracle_##name(smh_oracle *o, uint32_t wh[2]) {   \
        if (!o || !o->frame) return NULL;                               \
        if (wh) { wh[0] = o->roi[wi]; wh[1] = o->roi[hi]; }             \
        return o->field;                                                \
    }
GETTER(cropped_map, cropped_map, 6, 7)
GETTER(cropped_brq, cropped_brq, 10, 11)
GETTER(lsd_image, lsd_image, 6, 7)
GETTER(scales, scales, 10, 11)
GETTER(raw_mask, raw_mask, 6, 7)

/* ---------------------------------------------------------------- meters-per-pixel ratio (N2) */

/* image.get_pixel_fast(x, y).0[0] — unchecked in release builds (util/src/image.rs:136-140), a panic in
 * debug builds.  D16: a pixel outside the image reads as non-black, i.e. it never belongs to a bar. */
static inline uint8_t gray_at(const uint8_t *img, uint32_t w, uint32_t h, uint32_t x, uint32_t y) {
    return (x < w && y < h) ? img[(size_t)y * w + x] : 255;
}

/* "Make sure this is the vertical line upwards": `(y..y + 4).chain((y..y - 4).rev())` — the second
 * range is empty (end < start), so the test is the four pixels (x, y) .. (x, y + 3) —
 * src/vision/mpx_ratio.rs:24-29,45-50 */
static inline int vertical_bar(const uint8_t *img, uint32_t w, uint32_t h, uint32_t x, uint32_t y) {
    for (uint32_t yy = y; yy < y + SMH_SCALE_MIN_VBAR_HEIGHT; yy++)
        if (gray_at(img, w, h, x, yy) != 0) return 0;
    return 1;
}

/* find_scale_width — src/vision/mpx_ratio.rs:4-77.  Returns 1 (Some) and fills *ratio and
 * line = {left, right, y}, or 0 (None). */
static int find_scale_width(uint32_t meters, uint32_t x, uint32_t y0, const uint8_t *img, uint32_t w, uint32_t h,
                            double *ratio, uint32_t line[3]) {
    if (y0 < SMH_SCALE_MIN_VBAR_HEIGHT) return 0;
    if (x >= w) return 0; /* D16: the reference reads out of bounds here */
    const uint32_t max_off = (uint32_t)llround((20.0 / 640.0) * (double)w);
    const uint64_t y_end64 = (uint64_t)y0 + max_off;
    const uint32_t y_end = y_end64 < h ? (uint32_t)y_end64 : h;
    for (uint32_t y = y0; y < y_end; y++) {
        if (img[(size_t)y * w + x] != 0) continue;
        /* go right */
        uint32_t right = 0;
        for (uint32_t xr = x; xr < w; xr++)
            if (vertical_bar(img, w, h, xr, y)) { right = xr; break; }
        if (right == 0) continue;
        right -= 1;
        /* go left */
        uint32_t left = 0;
        for (uint32_t xl = x; xl-- > 0;)
            if (vertical_bar(img, w, h, xl, y)) { left = xl; break; }
        if (left == 0) continue;
        left += 1;
        const uint32_t width = right - left; /* wraps like the release build when right < left */
        if (width < SMH_SCALE_MIN_WIDTH) continue;
        line[0] = left; line[1] = right; line[2] = y;
        *ratio = (double)meters / (double)width;
        return 1;
    }
    return 0;
}

int smh_oracle_calc_meters_to_px_ratio(const uint8_t *img, uint32_t w, uint32_t h, const uint32_t (*scales)[3],
                                       uint32_t n_scales, double *ratio, uint32_t debug[3][4], uint32_t *n_debug) {
    /* src/vision/mpx_ratio.rs:79-133: the mean of the scales that were found, summed in order */
    if (n_debug) *n_debug = 0;
    if (n_scales == 0 || n_scales > 3) return 0;
    double sum = 0.0;
    uint32_t found = 0;
    for (uint32_t k = 0; k < n_scales; k++) {
        double r;
        uint32_t line[3];
        if (find_scale_width(scales[k][0], scales[k][1], scales[k][2], img, w, h, &r, line)) {
            sum = found ? sum + r : r;
            if (debug && n_debug) {
                debug[*n_debug][0] = scales[k][0]; debug[*n_debug][1] = line[0]; debug[*n_debug][2] = line[1]; debug[*n_debug][3] = line[2];
                (*n_debug)++;
            }
            found++;
        }
    }
    if (!found) return 0;
    *ratio = found == 1 ? sum : sum / (double)found;
    return 1;
}

/* ---------------------------------------------------------------- find_minimap (N3) */

/* get_edginess — src/vision/find_minimap.rs:8-46: max over the 8 neighbours of the summed absolute
 * B,G,R differences, as f32 / 765.0.  (x, y) relative to the view; needs a 1-px margin. */
static float edginess(const uint8_t *bgra, uint32_t pitch_px, uint32_t vx, uint32_t vy, uint32_t x, uint32_t y) {
    static const int off[8][2] = {{-1, -1}, {0, -1}, {1, -1}, {-1, 1}, {0, 1}, {1, 1}, {-1, 0}, {1, 0}};
    const uint8_t *p = bgra + ((size_t)(vy + y) * pitch_px + (vx + x)) * 4;
    uint16_t mx = 0;
    for (int k = 0; k < 8; k++) {
        const uint8_t *q = bgra + ((size_t)(vy + y + off[k][1]) * pitch_px + (vx + x + off[k][0])) * 4;
        uint16_t s = 0;
        for (int c = 0; c < 3; c++) s += (uint16_t)abs((int)p[c] - (int)q[c]);
        if (s > mx) mx = s;
    }
    return (float)mx / 765.0f;
}

/* find_edge — src/vision/find_minimap.rs:65-129.  dir: 0 Left, 1 Right, 2 Up, 3 Down. */
static uint32_t find_edge(const uint8_t *bgra, uint32_t pitch_px, uint32_t vx, uint32_t vy, uint32_t w, uint32_t h,
                          uint32_t x, uint32_t y, int dir) {
    const float THRESHOLD = SMH_MINIMAP_EDGINESS_THRESHOLD;
    uint32_t xy[2] = {x, y};
    const int c = dir < 2 ? 0 : 1, oc = 1 - c;
    const int cod = (dir & 1) ? 1 : -1;
    uint32_t c_max = c == 0 ? w : h, oc_max = c == 0 ? h : w;
    const uint32_t d = oc_max > xy[oc] ? oc_max - xy[oc] : xy[oc] - oc_max;
    const uint32_t min_line_length = d / 2 - 1;
    c_max -= 3; oc_max -= 3;
    for (;;) {
        xy[c] = (uint32_t)((int32_t)xy[c] + cod);
        if (xy[c] > c_max) return c_max + 2;
        else if (xy[c] < 3) return 0;
        if (edginess(bgra, pitch_px, vx, vy, xy[0], xy[1]) <= THRESHOLD) {
            const uint32_t ret = xy[c];
            uint32_t q[2] = {xy[0], xy[1]};
            uint32_t left = min_line_length;
            int ok = 1;
            while (left > 0) {
                q[oc] = (uint32_t)((int32_t)q[oc] - cod);
                if (q[oc] < 3 || q[oc] > oc_max) { ok = 0; break; }
                if (edginess(bgra, pitch_px, vx, vy, q[0], q[1]) <= THRESHOLD) left--;
                else { ok = 0; break; }
            }
            if (ok) return (uint32_t)((int32_t)ret - cod);
        }
    }
}

int smh_oracle_find_minimap_in(const uint8_t *bgra, uint32_t pitch_px, uint32_t vx, uint32_t vy, uint32_t w, uint32_t h,
                               uint32_t rect[4]) {
    /* src/vision/find_minimap.rs:51-146 -> Rect { left, right, top, bottom } */
    if (w < 3 || h < 3) return 0;
    const uint32_t x = w / 2, y = h / 2;
    for (int dir = 0; dir < 4; dir++) rect[dir] = find_edge(bgra, pitch_px, vx, vy, w, h, x, y, dir);
    return 1;
}

int smh_oracle_find_minimap(smh_oracle *o, uint32_t rect[4]) {
    /* src/vision/mod.rs:85: find_minimap(vision.get_cpu_frame().view(x, y, w, h)) */
    if (!o || !o->frame) return -2;
    return smh_oracle_find_minimap_in(o->frame, o->W, o->roi[4], o->roi[5], o->roi[6], o->roi[7], rect);
}

int smh_oracle_process_frame(smh_oracle *o, const uint8_t *bgra, uint32_t W, uint32_t H,
                             int grayscale, uint32_t start_y, uint32_t max_gap,
                             float lines[32][4], uint32_t *n_lines) {
    /* call order of VisionState::process — src/vision/mod.rs:66-223 */
    int rc = smh_oracle_load_frame(o, bgra, W, H);
    if (rc) return rc;
    int open = 0;
    rc = smh_oracle_crop_to_map(o, grayscale, &open, NULL, NULL, NULL);
    if (rc) return rc;
    *n_lines = 0;
    if (!open) return 0;
    smh_oracle_isolate_map_markers(o);
    smh_oracle_mask_marker_lines(o);
    smh_oracle_find_marker_lines(o, max_gap, lines, n_lines);
    smh_oracle_ocr_preprocess(o, NULL, NULL);
    smh_oracle_find_scales_preprocess(o, start_y, NULL, NULL);
    return 1;
}

/* ---------------------------------------------------------------- vision-gpu flavour (see smh_oracle.h)
 * Restates the reference's CUDA kernels so that their object code can pin this file bit for bit. */

/* CUDA `(uint32_t)f`: cvt.rzi.u32.f32 — truncates, saturates, NaN -> 0: the same as Rust's `as u32` */
uint32_t smh_oracle_gpuflavour_ray(const uint8_t *img, uint32_t w, uint32_t h, float px, float py, float max_gap,
                                   float dx, float dy, float end[2]) {
    /* vision-gpu/cuda/cuda.cu:656-722 */
    float x = px, y = py;
    uint32_t gap_count = 0, gap = 0, d5 = 0;
    float gap_x = 0.0f, gap_y = 0.0f, xo = 0.0f, yo = 0.0f;
    const float wf = (float)w, hf = (float)h;
    while (x >= 0.0f && y >= 0.0f && x < wf && y < hf) {
        const uint8_t v = img[(size_t)f32_as_u32(y) * w + f32_as_u32(x)];
        if (gap == 0) {
            if (v != 255) { gap = 1; gap_x = x; gap_y = y; gap_count++; }
        } else {
            if (v == 255) { gap = 0; gap_x = 0.0f; gap_y = 0.0f; }
            else if ((float)gap >= max_gap || (float)gap_count >= max_gap) {
                if (!((float)gap >= max_gap)) d5 = 1;
                x = gap_x; y = gap_y;
                break;
            } else gap++;
        }
        xo += dx; yo += dy;
        x = xo + px; y = yo + py;
    }
    end[0] = px; end[1] = py;
    const uint32_t xi = f32_as_u32(x), yi = f32_as_u32(y);
    if (xi < w && yi < h && img[(size_t)yi * w + xi] == 0) { end[0] = x - dx; end[1] = y - dy; }
    return d5;
}

int smh_oracle_gpuflavour_rays_in(const uint8_t *img, uint32_t w, uint32_t h, float px, float py, float max_gap,
                                  const float *dxdy, float *ends, uint8_t *d5) {
#pragma omp parallel for schedule(static)
    for (int k = 0; k < SMH_LSD_RAYS; k++) {
        const uint32_t f = smh_oracle_gpuflavour_ray(img, w, h, px, py, max_gap, dxdy[2 * k], dxdy[2 * k + 1], ends + 2 * k);
        if (d5) d5[k] = (uint8_t)f;
    }
    return 0;
}

int smh_oracle_rays_in_dirs(const uint8_t *img, uint32_t w, uint32_t h, float px, float py, float max_gap,
                            const float *dxdy, float *ends) {
#pragma omp parallel for schedule(static)
    for (int k = 0; k < SMH_LSD_RAYS; k++) march(img, w, h, px, py, max_gap, dxdy[2 * k], dxdy[2 * k + 1], ends + 2 * k);
    return 0;
}

void smh_oracle_gpuflavour_dilate(const uint8_t *raw, uint32_t w, uint32_t h, uint8_t *out) {
    /* dilate.cpp:12-24: src/dst advanced to (1,1), ROI (w-2, h-2), 2x2 mask of 255s anchored (1,1):
     * dst(x,y) = max over mask(i,j) != 0 of src(x + i - 1, y + j - 1) */
    if (w < 3 || h < 3) return;
    for (uint32_t y = 1; y + 1 < h; y++)
        for (uint32_t x = 1; x + 1 < w; x++) {
            uint8_t m = raw[(size_t)y * w + x];
            const uint8_t a = raw[(size_t)(y - 1) * w + x - 1], b = raw[(size_t)(y - 1) * w + x], c = raw[(size_t)y * w + x - 1];
            if (a > m) m = a;
            if (b > m) m = b;
            if (c > m) m = c;
            out[(size_t)y * w + x] = m;
        }
}

uint16_t smh_oracle_gpuflavour_icon_sad(const uint8_t *map_rgb, uint32_t stride, uint32_t x, uint32_t y,
                                        const uint8_t *tpl_rgba, uint32_t size) {
    /* cuda.cu:571-583 */
    uint16_t sad = 0;
    for (uint32_t my = 0; my < size; my++)
        for (uint32_t mx = 0; mx < size; mx++) {
            const uint8_t *t = tpl_rgba + ((size_t)my * size + mx) * 4;
            const uint8_t *q = map_rgb + ((size_t)(y + my) * stride + (x + mx)) * 3;
            uint16_t ad = 0;
            for (int k = 0; k < 3; k++) ad = (uint16_t)(ad + (q[k] > t[k] ? q[k] - t[k] : t[k] - q[k]));
            const double v = (double)(float)ad * ((double)(float)t[3] / 255.0); /* `(float)ad * ((float)a / 255.0)` */
            ad = (uint16_t)v; /* cvt.rzi.u16.f64; 0 <= v <= 765 */
            sad = (uint16_t)(sad + ad);
        }
    return sad;
}

void smh_oracle_gpuflavour_icon_clear(uint8_t *map_rgb, uint32_t w, uint32_t h, uint32_t rx, uint32_t ry, uint32_t size) {
    /* cuda.cu:594-613 under a grid of ceil(size/8)^2 blocks of (8,8) threads */
    const uint32_t ext = (size + 7u) / 8u * 8u;
    const uint32_t sq_x = rx + size / 2u, sq_y = ry + (uint32_t)roundf((float)size * SMH_MAP_MARKER_POI_LOCATION);
    for (uint32_t j = 0; j < ext; j++)
        for (uint32_t i = 0; i < ext; i++) {
            const uint32_t x = rx + i, y = ry + j;
            if (x >= w || y >= h) continue;
            uint8_t *q = map_rgb + ((size_t)y * w + x) * 3;
            q[0] = q[1] = q[2] = 0;
            if (x >= sq_x - 2u && x <= sq_x + 2u && y >= sq_y - 2u && y <= sq_y + 2u) { q[0] = 0; q[1] = 255; q[2] = 0; }
        }
}

int smh_oracle_gpuflavour_is_marker(uint8_t r8, uint8_t g8, uint8_t b8) {
    /* cuda.cu:191-239 (to_hsv), 277-309 (is_any_map_marker_color); CUDA's (uintN_t)f saturates like Rust's `as` */
    const float r = (float)r8 / 255.0f, g = (float)g8 / 255.0f, b = (float)b8 / 255.0f;
    const float mx = fmaxf(r, fmaxf(g, b)), mn = fminf(r, fminf(g, b)), diff = mx - mn;
    float h = 0.0f, s = 0.0f;
    const float v = mx;
    if (v != 0.0f) {
        s = diff / v;
        if (!(diff < 0.001f)) {
            if (mx == r) { h = 60.0f * (g - b) / diff; if (h < 0.0f) h += 360.0f; }
            else if (mx == g) h = 60.0f * (2 + (b - r) / diff);
            else h = 60.0f * (4 + (r - g) / diff);
        }
    }
    const int H = (int)f32_as_u16(h), S = (int)f32_as_u8(s * 100.f), V = (int)f32_as_u8(v * 100.f);
    static const int M[3][3] = {{105, 100, 100}, {285, 46, 85}, {158, 60, 91}};  /* consts.toml:33-43 */
    if (S < SMH_MIN_SAT) return 0;
    for (int t = 0; t < 3; t++) {
        const int hue_ok = abs(H - M[t][0]) <= 15, sat_ok = abs(S - M[t][1]) <= 15, light_ok = abs(S - (M[t][1] - 50)) <= 15;
        const int vib_ok = abs(V - M[t][2]) <= 15;
        if (hue_ok && (sat_ok || light_ok) && vib_ok) return 1;
    }
    return 0;
}

void smh_oracle_gpuflavour_table_marker(uint8_t *out) {
#pragma omp parallel for schedule(static)
    for (uint32_t c = 0; c < (1u << 24); c++) out[c] = (uint8_t)smh_oracle_gpuflavour_is_marker((uint8_t)(c >> 16), (uint8_t)(c >> 8), (uint8_t)c);
}

/* ---------------------------------------------------------------- synthetic captures (bench workload) */
int smh_oracle_synth_frames(const void *params, uint32_t first, uint32_t n, uint8_t *out) {
    const smh_synth_params *p = (const smh_synth_params *)params;
    uint32_t roi[12];
    if (!p || !out || p->n_labels > SMH_SYNTH_MAX_LABELS || smh_oracle_roi(p->W, p->H, roi)) return -1;
    smh_synth_segment *segs = (smh_synth_segment *)malloc(sizeof(smh_synth_segment) * (p->n_segments + 1));
    smh_synth_blip *blips = (smh_synth_blip *)malloc(sizeof(smh_synth_blip) * (p->n_blips + 1));
    for (uint32_t f = 0; f < n; f++) {
        smh_synth_frame_desc d;
        smh_synth_scene_frame(p, roi, first + f, &d, segs, blips);
        uint32_t *dst = (uint32_t *)(out + (size_t)f * p->W * p->H * 4);
#pragma omp parallel for schedule(static)
        for (uint32_t y = 0; y < p->H; y++)
            for (uint32_t x = 0; x < p->W; x++) dst[(size_t)y * p->W + x] = smh_synth_pixel(&d, segs, blips, (int32_t)x, (int32_t)y);
    }
    free(segs); free(blips);
    return 0;
}
