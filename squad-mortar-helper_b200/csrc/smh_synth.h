// smh_synth.h — deterministic synthetic map screenshots (SURVEY.md 8(d) generator).
//
// Stateless and integer-only: a pixel is a pure function of (scene, x, y), so the CPU
// rasteriser (oracle inputs, tests) and the CUDA rasteriser (device-resident bench batches)
// produce byte-identical frames.  Not part of the reference — the reference's inputs are
// OS screen grabs (src/capture.rs), which the benchmarks replace with these frames.
#pragma once
#include <stdint.h>
#include <string.h>
#ifndef __cplusplus
#include <stdbool.h>
#endif

#if defined(__CUDACC__)
#define SYN_HD __host__ __device__ __forceinline__
#else
#define SYN_HD static inline
#endif

typedef struct {
    int32_t x0, y0, x1, y1;  // frame coordinates
    uint32_t team;           // 0 alpha, 1 bravo, 2 charlie
} smh_synth_segment;

typedef struct {
    int32_t x, y;   // frame coordinates of the top-left corner
    uint32_t size;  // 2..4
    uint32_t team;
} smh_synth_blip;

typedef struct {
    int32_t x, y;  // frame coordinates of the label box (48x16) top-left; bar sits below
} smh_synth_label;

#define SMH_SYNTH_MAX_LABELS 4
#define SMH_SYNTH_LABEL_W 48
#define SMH_SYNTH_LABEL_H 16
#define SMH_SYNTH_BAR_W 60

typedef struct {
    uint64_t seed;  // per-frame seed = base seed + frame index
    uint32_t W, H;
    uint32_t btn[4], map[4];  // ROIs (x,y,w,h) in frame coordinates
    uint32_t map_open;        // paint the red "close deployment" button
    uint32_t n_segments, n_blips, n_labels;
    smh_synth_label labels[SMH_SYNTH_MAX_LABELS];
} smh_synth_frame_desc;

SYN_HD uint64_t smh_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// team colours = RGB of the three HSV centres in consts.toml:33-43
SYN_HD void smh_synth_team_rgb(uint32_t team, uint32_t jit, uint32_t* r, uint32_t* g, uint32_t* b) {
    const int jr = (int)(jit % 5u) - 2, jg = (int)((jit / 5u) % 5u) - 2, jb = (int)((jit / 25u) % 5u) - 2;
    int R, G, B;
    if (team == 0) { R = 64; G = 253; B = 2; }          // ~HSV(105,100,100), kept inside 0..255 under jitter
    else if (team == 1) { R = 192; G = 117; B = 217; }  // ~HSV(285,46,85)
    else { R = 93; G = 232; B = 181; }                  // ~HSV(158,60,91)
    *r = (uint32_t)(R + jr); *g = (uint32_t)(G + jg); *b = (uint32_t)(B + jb);
}

// 4*dist^2(P, segment AB) <= 9  (half-width 1.5 px), exact in int64
SYN_HD bool smh_synth_on_segment(int32_t px, int32_t py, const smh_synth_segment* s) {
    const int64_t ax = s->x0, ay = s->y0, bx = s->x1, by = s->y1;
    const int64_t dx = bx - ax, dy = by - ay, qx = px - ax, qy = py - ay;
    const int64_t L2 = dx * dx + dy * dy;
    const int64_t t = qx * dx + qy * dy;
    if (t <= 0) return 4 * (qx * qx + qy * qy) <= 9;
    if (t >= L2) { const int64_t ex = px - bx, ey = py - by; return 4 * (ex * ex + ey * ey) <= 9; }
    const int64_t cr = qx * dy - qy * dx;
    return 4 * cr * cr <= 9 * L2;
}

// BGRA little-endian word: b | g<<8 | r<<16 | a<<24
SYN_HD uint32_t smh_synth_pixel(const smh_synth_frame_desc* d, const smh_synth_segment* segs,
                                const smh_synth_blip* blips, int32_t x, int32_t y) {
    const uint64_t hsh = smh_splitmix64(d->seed * 0x100000001B3ull + (uint64_t)y * d->W + (uint64_t)x);
    uint32_t b = 20u + (uint32_t)(hsh & 0xFFu) % 101u;
    uint32_t g = 20u + (uint32_t)((hsh >> 8) & 0xFFu) % 101u;
    uint32_t r = 20u + (uint32_t)((hsh >> 16) & 0xFFu) % 101u;
    const uint32_t jit = (uint32_t)((hsh >> 24) & 0xFFFFu);

    // red button (>= 80 % of its pixels inside the +-25 tolerance of (217,67,49))
    if (d->map_open && (uint32_t)x - d->btn[0] < d->btn[2] && (uint32_t)y - d->btn[1] < d->btn[3]) {
        if (jit % 100u < 85u) {
            r = 217u - 10u + (jit >> 3) % 21u;
            g = 67u - 10u + (jit >> 5) % 21u;
            b = 49u - 10u + (jit >> 7) % 21u;
        }
        return b | (g << 8) | (r << 16) | 0xFF000000u;
    }
    const uint32_t mx = (uint32_t)x - d->map[0], my = (uint32_t)y - d->map[1];
    if (mx < d->map[2] && my < d->map[3]) {
        // scale labels + bars (bottom-right quadrant text for the OCR / scales stages)
        for (uint32_t k = 0; k < d->n_labels; k++) {
            const int32_t lx = x - d->labels[k].x, ly = y - d->labels[k].y;
            if (lx >= 0 && lx < SMH_SYNTH_LABEL_W && ly >= 0 && ly < SMH_SYNTH_LABEL_H) {
                // glyph core: 3x3 checker inside a 5-px margin; the rest is anti-aliased fringe
                const bool inner = lx >= 5 && lx < SMH_SYNTH_LABEL_W - 5 && ly >= 5 && ly < SMH_SYNTH_LABEL_H - 5;
                if (inner && (((lx - 5) / 3 + (ly - 5) / 3) & 1) == 0) {
                    const uint32_t v = 200u + jit % 56u;  // bright, r == g == b
                    return v | (v << 8) | (v << 16) | 0xFF000000u;
                }
                const uint32_t v = 130u + jit % 64u;  // fringe: >= 130, channel spread <= 6
                const uint32_t fr = v + (jit >> 6) % 7u, fg = v + (jit >> 9) % 7u, fb = v + (jit >> 12) % 7u;
                return fb | (fg << 8) | (fr << 16) | 0xFF000000u;
            }
            // black scale bar  I-----I  three pixels below the label
            const int32_t by = ly - (SMH_SYNTH_LABEL_H + 3);
            if (lx >= 0 && lx < SMH_SYNTH_BAR_W) {
                if (by == 2 || ((lx == 0 || lx == SMH_SYNTH_BAR_W - 1) && by >= 0 && by <= 4))
                    return 0xFF000000u;  // luma 0
            }
        }
        // marker segments (+ a radius-6 icon disc at the start point) and blips; last one wins
        int hit_team = -1;
        for (uint32_t k = 0; k < d->n_segments; k++) {
            const smh_synth_segment* s = &segs[k];
            const int32_t lo_x = (s->x0 < s->x1 ? s->x0 : s->x1) - 7, hi_x = (s->x0 > s->x1 ? s->x0 : s->x1) + 7;
            const int32_t lo_y = (s->y0 < s->y1 ? s->y0 : s->y1) - 7, hi_y = (s->y0 > s->y1 ? s->y0 : s->y1) + 7;
            if (x < lo_x || x > hi_x || y < lo_y || y > hi_y) continue;
            const int32_t ex = x - s->x0, ey = y - s->y0;
            if (smh_synth_on_segment(x, y, s) || ex * ex + ey * ey <= 36) hit_team = (int)s->team;
        }
        for (uint32_t k = 0; k < d->n_blips; k++) {
            const smh_synth_blip* q = &blips[k];
            if ((uint32_t)(x - q->x) < q->size && (uint32_t)(y - q->y) < q->size) hit_team = (int)q->team;
        }
        if (hit_team >= 0) smh_synth_team_rgb((uint32_t)hit_team, jit, &r, &g, &b);
    }
    return b | (g << 8) | (r << 16) | 0xFF000000u;
}


// ---- scene of one frame (host side; shared by the library's rasterisers and the oracle's, so that the
//      reference arm of bench.py can make the workload's frames without loading the product library) ----------
typedef struct {
    uint64_t seed;        // frame k uses seed + k
    uint32_t W, H;
    uint32_t n_segments, n_blips, n_labels, map_open, blip_band_pct, reserved;
} smh_synth_params;       // = smh_b200_synth_params (include/smh_b200.h)

static inline uint32_t smh_synth_below(uint64_t* s, uint32_t n) {
    *s = smh_splitmix64(*s);
    return n ? (uint32_t)(*s % n) : 0u;
}

// roi = {button xywh, map xywh, brq xywh} of the frame size (smh_b200_roi / smh_oracle_roi).  segs: n_segments entries,
// blips: n_blips entries.
static inline void smh_synth_scene_frame(const smh_synth_params* p, const uint32_t roi[12], uint32_t frame_index,
                                         smh_synth_frame_desc* d, smh_synth_segment* segs, smh_synth_blip* blips) {
    memset(d, 0, sizeof *d);
    d->seed = p->seed + frame_index;
    d->W = p->W; d->H = p->H;
    memcpy(d->btn, roi, 16); memcpy(d->map, roi + 4, 16);
    d->map_open = p->map_open; d->n_segments = p->n_segments; d->n_blips = p->n_blips; d->n_labels = p->n_labels;
    const int32_t mx = (int32_t)roi[4], my = (int32_t)roi[5], mw = (int32_t)roi[6], mh = (int32_t)roi[7];
    uint64_t rng = smh_splitmix64(d->seed ^ 0x534D4800ull);
    for (uint32_t k = 0; k < p->n_segments; k++) {
        smh_synth_segment* s = &segs[k];
        for (;;) {  // random end points inside the map, 60..600 px apart
            s->x0 = mx + 10 + (int32_t)smh_synth_below(&rng, (uint32_t)mw - 20); s->y0 = my + 10 + (int32_t)smh_synth_below(&rng, (uint32_t)mh - 20);
            s->x1 = mx + 10 + (int32_t)smh_synth_below(&rng, (uint32_t)mw - 20); s->y1 = my + 10 + (int32_t)smh_synth_below(&rng, (uint32_t)mh - 20);
            const int64_t dx = s->x1 - s->x0, dy = s->y1 - s->y0, l2 = dx * dx + dy * dy;
            if (l2 >= 60 * 60 && l2 <= 600 * 600) break;
        }
        s->team = smh_synth_below(&rng, 3);
    }
    const uint32_t band = p->blip_band_pct == 0 || p->blip_band_pct > 100 ? 100 : p->blip_band_pct;
    const uint32_t bh = (uint32_t)((uint64_t)mh * band / 100u);
    for (uint32_t k = 0; k < p->n_blips; k++) {
        smh_synth_blip* q = &blips[k];
        q->size = 2 + smh_synth_below(&rng, 3);
        q->x = mx + 6 + (int32_t)smh_synth_below(&rng, (uint32_t)mw - 12);
        q->y = my + 6 + (int32_t)smh_synth_below(&rng, (bh > 12 ? bh : 13) - 12);
        q->team = smh_synth_below(&rng, 3);
    }
    // labels inside the bottom-right quadrant; the last one hugs the right/bottom edge (D10 clip)
    const int32_t qx = (int32_t)roi[8], qy = (int32_t)roi[9], qw = (int32_t)roi[10], qh = (int32_t)roi[11];
    for (uint32_t k = 0; k < p->n_labels; k++) {
        const int32_t spanx = qw - SMH_SYNTH_BAR_W, spany = qh - (SMH_SYNTH_LABEL_H + 8);
        if (k + 1 == p->n_labels && k > 0) {
            d->labels[k].x = qx + qw - SMH_SYNTH_LABEL_W + 2;           // crosses the right edge of the map
            d->labels[k].y = qy + qh - SMH_SYNTH_LABEL_H + 3;           // crosses the bottom edge
        } else {
            d->labels[k].x = qx + (spanx > 1 ? (int32_t)smh_synth_below(&rng, (uint32_t)spanx) : 0);
            d->labels[k].y = qy + (spany > 1 ? (int32_t)smh_synth_below(&rng, (uint32_t)spany) : 0);
        }
    }
}
