// kernels_lsd.cuh — k_lsd2: lsd::find_lines for whole frames, one small CTA (NW warps) per frame in flight.
//
// Replaces the reference's K10 + host loop (vision-common/src/lsd.rs:60-105,
// vision-cpu/src/lib.rs:387-449, vision-gpu/src/lib.rs:477-538: one launch + copy + sync per seed).
//
// find_lines is a strictly sequential chain per frame (every accepted line suppresses later seeds), so the
// parallelism of a batch is ACROSS frames first; within a frame only the bulk phases of a seed are worth
// spreading.  Warp 0 of the CTA is the frame's MASTER: it alone runs the scalar chain (seed scan, the
// reference's suppression test, dead clusters, get_centre, accepting a line), so a uniform instruction is
// issued once per frame — the round-1 kernel issued it once per warp of a 256-thread CTA and spent most of its
// 1.9 M warp instructions per frame on that.  The other warps sleep on the barrier until a seed needs bulk
// work, which all NW warps share:
//   * the window around the seed (<= 6 x 6 tiles of the tile-major bitmap, one coalesced load per tile, held
//     row-major with an odd pitch in shared memory),
//   * the exact ray filter (smh_lsd_filter.h): a warp lists the white 8-pixel quarters of its tiles and marks
//     the rays each can serve, one quarter per lane, into the per-ring bitmaps,
//   * the march: the AND of the rings is taken word by word (32 consecutive rays = one warp-wide group, lane =
//     ray, first 64 samples out of the window); survivors go to a shared list and continue one ray per warp,
//     32 samples per step on the global bitmap (ray_block_*); last-index argmax (rayon's right-biased reduce, D6).
// Dead clusters as in smh_lsd_filter.h: a seed whose first ring is empty proves a disc of seeds hopeless.
// Persistent CTAs pull frames from an atomic counter.  Preconditions (launch_lsd): 1 <= gap_limit <= 32.
#pragma once
#include "kernels.cuh"

namespace smh {

constexpr int L2_SURV_CAP = 512;   // phase-1 survivors of one seed (more: they are marched right away by their warp)
constexpr int L2_GRP_CAP = 1024;   // white quarters of a seed's window over all warps (more: every ray is marched)
constexpr int L2_WIN_T = 6;        // window tiles per axis: ceil((2 * (70 + 3) + 1 + 31) / 32)
constexpr int L2_WIN_PITCH = 7;    // words per window row (odd: lane = row stores are conflict-free)
constexpr int L2_DEAD = 16;
// Rays of a 32-ray group that keep it marching one ray per lane beyond the window, on the global bitmap.  Measured on
// B200 (1024 dense 1440p frames, profiles/r3w_sweep2.txt): 10 -> 3.96 ms, 16 -> 3.83, 24 -> 3.73, never (33) -> 3.47: the
// scattered bitmap loads of 32 diverging rays cost more than the ray-per-warp blocks they save, so it is off.
#ifndef L2_LANES_MIN_N
#define L2_LANES_MIN_N 33
#endif
constexpr int L2_LANES_MIN = L2_LANES_MIN_N;

enum { L2_JOB_SEED = 0, L2_JOB_ROW_DONE = 1 };

template <int NW>
struct Lsd2Shared {
    float4 lines[SMH_LSD_MAX_LINES];
    float4 dead[L2_DEAD];  // (x, y, rho^2, raw^2)
    uint32_t ring[SMH_LSD_FILTER_MAX_WINDOWS][SMH_LSD_BITMAP_WORDS + 3];
    uint32_t grp[L2_GRP_CAP];  // white quarters of the window: m8 | (x offset of its first column + 160) << 8 | (row offset + 128) << 17
    uint32_t surv[L2_SURV_CAP];         // l2_pack(ray, gap, next block)
    uint32_t win[L2_WIN_T * 32 * L2_WIN_PITCH];
    unsigned long long bkey[NW];
    float2 bend[NW];
    unsigned long long steps;
    struct Job { int job, isx, isy, skip; float px, py; } jobs[2];  // double-buffered: the master writes the next one while the last is read
    int n;
    uint32_t n_dead, n_surv, surv_end, n_grp, near2, far2, frame;
};

// The window of a seed, row-major: rows [y0, ..) x words [tx0, tx0 + ntx) of the bitmap.
struct Win2 {
    const uint32_t* p;
    int tx0, y0;
    uint32_t pad, w, h;
    __device__ __forceinline__ bool white(uint32_t x, uint32_t y) const {
        const uint32_t X = x + pad;
        return (p[((int)y - y0) * L2_WIN_PITCH + (int)(X >> 5) - tx0] >> (X & 31u)) & 1u;
    }
    // the same for ANY (x, y): a word of the window's array is read whatever the coordinates (the caller masks the
    // result of a sample outside the image), so that the load needs no branch around it
    __device__ __forceinline__ uint32_t white_anywhere(uint32_t x, uint32_t y) const {
        const uint32_t X = x + pad;
        const int ry = min(max((int)y - y0, 0), L2_WIN_T * 32 - 1), rt = min(max((int)(X >> 5) - tx0, 0), L2_WIN_T - 1);
        return (p[ry * L2_WIN_PITCH + rt] >> (X & 31u)) & 1u;
    }
};

// Clear columns [xlo, xhi] of row r of the tile-row in shared memory (pitch 33).
template <bool ATOMIC>
__device__ __forceinline__ void l2_clear(uint32_t* chunk, uint32_t pad, int r, int xlo, int xhi) {
    const int Xlo = xlo + (int)pad, Xhi = xhi + (int)pad;
    for (int t = Xlo >> 5; t <= (Xhi >> 5); t++) {
        uint32_t m = 0xFFFFFFFFu;
        if (t == (Xlo >> 5)) m &= 0xFFFFFFFFu << (Xlo & 31);
        if (t == (Xhi >> 5)) m &= 0xFFFFFFFFu >> (31 - (Xhi & 31));
        if (ATOMIC) atomicAnd(&chunk[t * 33 + r], ~m);
        else chunk[t * 33 + r] &= ~m;
    }
}
// lane = row: what `line` suppresses for sure in rows r >= r_from of the tile-row starting at row0
template <bool ATOMIC>
__device__ __forceinline__ void l2_strike_line(uint32_t* chunk, uint32_t pad, const float4& line, int row0, int r_from, int W) {
    const int r = threadIdx.x & 31;
    if (r >= r_from) {
        const Band b = line_band(line, (float)(row0 + r), W);
        if (b.ilo <= b.ihi) l2_clear<ATOMIC>(chunk, pad, r, b.ilo, b.ihi);
    }
}
template <bool ATOMIC>
__device__ __forceinline__ void l2_strike_disc(uint32_t* chunk, uint32_t pad, float cx, float cy, float raw2, int row0, int W) {
    const int r = threadIdx.x & 31;
    const float dy = (float)(row0 + r) - cy, r2 = raw2 - dy * dy;
    if (r2 > 0.0f) {
        const float hw = sqrtf(r2) - 0.01f;
        const int xlo = max((int)ceilf(cx - hw), 0), xhi = min((int)floorf(cx + hw), W - 1);
        if (xlo <= xhi) l2_clear<ATOMIC>(chunk, pad, r, xlo, xhi);
    }
}

// OR the ray interval [lo, hi] (any integers) into a ring bitmap.
__device__ __forceinline__ void l2_mark(uint32_t* bits, int lo, int hi) {
    if (hi - lo >= SMH_LSD_RAYS - 1) { lo = 0; hi = SMH_LSD_RAYS - 1; }
    else {
        if (lo < 0) { lo += SMH_LSD_RAYS; hi += SMH_LSD_RAYS; }
        if (lo >= SMH_LSD_RAYS) { lo -= SMH_LSD_RAYS; hi -= SMH_LSD_RAYS; }
    }
    const int e0 = min(hi, SMH_LSD_RAYS - 1);
    for (int w = lo >> 5; w <= (e0 >> 5); w++) or_if_new(&bits[w], range_mask(w, lo, e0));
    if (hi >= SMH_LSD_RAYS) {
        const int e1 = hi - SMH_LSD_RAYS;
        for (int w = 0; w <= (e1 >> 5); w++) or_if_new(&bits[w], range_mask(w, 0, e1));
    }
}

// `nsteps` more samples (a multiple of 8) of one ray per LANE, continuing from the accumulated offsets (xo, yo) and the
// gap state (gap, gx, gy); read from the seed's window (IMG = Win2: all of them lie inside it) or from the bitmap
// (IMG = Bits).  Returns 0: still marching (gap = current run of non-white samples), 1: stopped on a gap that began
// at (gx, gy), 2: (gx, gy) is the first sample outside the image.  Same arithmetic as march_fast (kernels.cuh); a
// batch of 8 white samples — the body of a marker line — skips the automaton.
template <bool STATS, typename IMG>
__device__ __forceinline__ int l2_march_lanes(const IMG& wv, int nsteps, float px, float py, int gap_limit, float dx, float dy,
                                              float& xo, float& yo, float& gx, float& gy, int& gap, uint32_t& steps) {
    float x = __fadd_rn(xo, px), y = __fadd_rn(yo, py);
    int state = 0;
    for (int n0 = 0; n0 < nsteps && state == 0; n0 += 8) {
        float xs[8], ys[8];
        uint32_t wh = 0, inb = 0;  // bit u: sample u is white / inside the image
#pragma unroll
        for (int u = 0; u < 8; u++) {
            xs[u] = x; ys[u] = y;
            const uint32_t ix = (uint32_t)(__float_as_int(__fadd_rd(x, 12582912.0f)) - 0x4B400000);
            const uint32_t iy = (uint32_t)(__float_as_int(__fadd_rd(y, 12582912.0f)) - 0x4B400000);
            // (no branch around the load: eight dependent branch-and-load rounds per batch were most of a window march)
            const uint32_t in = (ix < wv.w && iy < wv.h) ? 1u : 0u;
            inb |= in << u;
            wh |= (wv.white_anywhere(ix, iy) & in) << u;
            xo = __fadd_rn(xo, dx); yo = __fadd_rn(yo, dy);
            x = __fadd_rn(xo, px); y = __fadd_rn(yo, py);
        }
        if (wh == 0xFFu) { gap = 0; if (STATS) steps += 8u; continue; }
#pragma unroll
        for (int u = 0; u < 8; u++) {
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

struct L2Best { unsigned long long key; float ex, ey; };

// A ray handed over from the lane-per-ray march: ray index | current gap << 12 | next block (32 samples) << 18
__device__ __forceinline__ uint32_t l2_pack(int k, int gap, int c) { return (uint32_t)k | ((uint32_t)gap << 12) | ((uint32_t)c << 18); }

// One ray per WARP from its block on (samples 32 c ..) on the global bitmap; folds the result into `best`
// (the same value in every lane).  DEEP: the wide CTAs that serve one frame's latency (more loads in flight per ray:
// 0.90 -> 0.87 ms for a 1080p frame); a full batch is bound by throughput, and there the speculative loads cost 10 %.
template <bool STATS, bool DEEP>
__device__ __forceinline__ void l2_march_long(const Bits& im, const LsdArgs& a, uint32_t e, float px, float py, L2Best& best,
                                              uint32_t& steps_u) {
    const int lane = threadIdx.x & 31, sub = lane & (LSD_CKPT - 1);
    const int k2 = (int)(e & 0xFFFu), c0 = (int)(e >> 18);
    uint32_t gap_in = (e >> 12) & 0x3Fu;
    const float2 d = __ldg(&a.rays[k2]);
    const float2* cp = a.ckpt + (uint64_t)k2 * a.n_ckpt + (lane / LSD_CKPT) + c0 * (32 / LSD_CKPT);
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
    float gx, gy;
    if (DEEP) {
        // Four blocks at a time: their table values, then their bitmap words, are four independent loads each — a block at a
        // time the ray paid two dependent global-memory latencies per 32 samples.  (A ray rarely ends in the first block of
        // a group: it came this far because the line it follows is long.)  The table holds n_ckpt values per ray, two
        // blocks more than any ray inside the mask can need; a group that reaches beyond them repeats the last ones.
        const float2* const cend = a.ckpt + (uint64_t)(k2 + 1) * a.n_ckpt - 1;
        while (true) {
            float2 ck[4];
#pragma unroll
            for (int i = 1; i < 4; i++) { const float2* q = cp + i * (32 / LSD_CKPT); ck[i] = __ldg(q <= cend ? q : cend); }
            ck[0] = c1;
            RayBlock B[4];
#pragma unroll
            for (int i = 0; i < 4; i++) B[i] = ray_block_fetch(im, ck[i], sub, d.x, d.y, px, py);
            bool done = false;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (!done) {
                    done = ray_block_decide<STATS>(B[i], xprev, yprev, gap_in, a.gap_limit, gx, gy, steps_u);
                    xprev = B[i].x; yprev = B[i].y;
                }
            }
            if (done) break;
            cp += 4 * (32 / LSD_CKPT);
            c1 = __ldg(cp <= cend ? cp : cend);
        }
    } else {
        // two blocks in flight: block c+1 is fetched (table value loaded one iteration earlier) before block c is decided
        cp += 32 / LSD_CKPT;
        float2 c2 = __ldg(cp);
        RayBlock b = ray_block_fetch(im, c1, sub, d.x, d.y, px, py);
        while (true) {
            cp += 32 / LSD_CKPT;
            const float2 c3 = __ldg(cp);
            const RayBlock bn = ray_block_fetch(im, c2, sub, d.x, d.y, px, py);
            if (ray_block_decide<STATS>(b, xprev, yprev, gap_in, a.gap_limit, gx, gy, steps_u)) break;
            xprev = b.x; yprev = b.y; b = bn; c2 = c3;
        }
    }
    float ex, ey;
    const float len = ray_finish(im, gx, gy, px, py, d.x, d.y, ex, ey);
    const unsigned long long key = ((unsigned long long)__float_as_uint(len) << 32) | (uint32_t)k2;
    if (key >= best.key) { best.key = key; best.ex = ex; best.ey = ey; }
}

#ifndef L2_MIN_CTAS
#define L2_MIN_CTAS 5
#endif
// (registers: the 4-warp form is held to what five resident CTAs allow, 94: left to itself the compiler takes 123, four
// CTAs fill the register file, and the next batch's pixel pass finds no room beside the line search.  Measured with
// batches in flight, ms per 1024-frame step: 123 registers 4.66, 94 -> 4.57 - 4.69, 80 -> 4.88, 72 -> 4.81)
template <int NW, bool STATS>
__global__ void __launch_bounds__(NW * 32, NW <= 4 ? L2_MIN_CTAS : 1) k_lsd2(const LsdArgs a) {
    extern __shared__ __align__(16) uint32_t l2_smem[];
    using Shared = Lsd2Shared<NW>;
    Shared& sh = *reinterpret_cast<Shared*>(l2_smem);
    uint32_t* chunk = l2_smem + (sizeof(Shared) + 3) / 4;  // [tx][33] current tile-row: pixels that can still become seeds
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool master = warp == 0;
    const uint32_t FULL = 0xFFFFFFFFu, LT = (1u << lane) - 1u;
    const int TX = (int)a.g.tx, W = (int)a.g.w, H = (int)a.g.h;
    const uint32_t PAD = a.g.pad;
    smh_lsd_filter F;
    smh_lsd_filter_init(&F, a.gap_limit);
    const bool dead_ok = F.nw >= 1 && F.border[0] == 0;  // "ring 1 empty => no line" holds for every seed
    const int win_steps = F.radius >= 65 ? 64 : (F.radius >= 33 ? 32 : 0);  // samples that stay inside the window
    const int WR = F.radius + LSD_WIN_EXTRA;  // window radius around the raw seed pixel

    for (int i = threadIdx.x; i < SMH_LSD_FILTER_MAX_WINDOWS * (SMH_LSD_BITMAP_WORDS + 3); i += NW * 32) (&sh.ring[0][0])[i] = 0u;

    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) {
            sh.frame = atomicAdd(a.counter, 1u);
            sh.n = 0; sh.n_dead = 0u; sh.n_surv = 0u; sh.surv_end = (uint32_t)L2_SURV_CAP; sh.n_grp = 0u;
            sh.near2 = 0u; sh.far2 = __float_as_uint(3.0e38f); sh.steps = 0ull;
        }
        __syncthreads();
        const uint32_t f = sh.frame;
        if (f >= a.n_frames) break;
        float* out_lines = a.lines + (uint64_t)f * SMH_LSD_MAX_LINES * 4;
        if (a.open && a.open[f] == 0) {  // map closed: the reference returns before any line search
            if (threadIdx.x == 0) { a.n_lines[f] = 0; if (STATS) { a.stats[3 * f] = 0; a.stats[3 * f + 1] = 0; a.stats[3 * f + 2] = 0; } }
            continue;
        }
        const Bits im{a.bits + (uint64_t)f * a.bits_stride, a.g.tx, PAD, a.g.w, a.g.h};
        uint32_t steps = 0, steps_u = 0;          // per lane / per warp sample counters (STATS)
        uint32_t seeds = 0, pruned = 0, n_dead = 0;  // master only
        int n = 0;                                 // master's copy of sh.n
        int par = 0;                               // job slot of the next seed (uniform)

        for (uint32_t tyi = 0; tyi < a.g.ty; tyi++) {
            if (sh.n >= SMH_LSD_MAX_LINES) break;  // uniform: written before the last barrier
            // ---- the tile-row (contiguous in memory) into shared memory, tiles split over the warps ------
            const uint32_t* src = im.p + (uint64_t)tyi * (32u * a.g.tx);
            uint32_t any = 0;
#pragma unroll 4
            for (int t = warp; t < TX; t += NW) {
                const uint32_t w = __ldg(src + t * 32 + lane);
                chunk[t * 33 + lane] = w;
                any |= w;
            }
            if (tyi + 1 < a.g.ty)  // pull the next tile-row towards the SM while this one is searched
                for (int t = threadIdx.x; t < TX; t += NW * 32) asm volatile("prefetch.global.L2 [%0];" :: "l"(src + 32 * TX + t * 32));
            if (!__syncthreads_or(any != 0u)) continue;
            const int row0 = (int)(tyi * 32u);
            {
                const int nl = sh.n;
                for (int l = warp; l < nl; l += NW) l2_strike_line<(NW > 1)>(chunk, PAD, sh.lines[l], row0, 0, W);
                if (!STATS) {
                    const int nd = (int)min(sh.n_dead, (uint32_t)L2_DEAD);
                    for (int d = warp; d < nd; d += NW) l2_strike_disc<(NW > 1)>(chunk, PAD, sh.dead[d].x, sh.dead[d].y, sh.dead[d].w, row0, W);
                }
            }
            __syncthreads();

            // ---- seeds of this tile-row, in raster order ---------------------------------------------
            int r_cur = 0;         // master: rows before it hold no candidate any more
            uint32_t rowmask = 0;  // master: rows that may still hold one
            if (master) {
                uint32_t nzr = 0;
                for (int t = 0; t < TX; t++) nzr |= chunk[t * 33 + lane];
                rowmask = __ballot_sync(FULL, nzr != 0u);
            }
            while (true) {
                // ======== master: the scalar chain up to the next seed that needs bulk work ================
                if (master) {
                    int job = L2_JOB_ROW_DONE, isx = -1, isy = 0;
                    float sx = 0.0f, sy = 0.0f;
                    while (n < SMH_LSD_MAX_LINES) {
                        isx = -1;
                        while (rowmask) {
                            r_cur = __ffs(rowmask) - 1;
                            for (int t0 = 0; t0 < TX && isx < 0; t0 += 32) {
                                const int t = t0 + lane;
                                const uint32_t w = t < TX ? chunk[t * 33 + r_cur] : 0u;
                                const uint32_t nz = __ballot_sync(FULL, w != 0u);
                                if (nz) {
                                    const int src_lane = __ffs(nz) - 1;
                                    const uint32_t word = __shfl_sync(FULL, w, src_lane);
                                    const int bit = __ffs(word) - 1;
                                    if (lane == src_lane) chunk[t * 33 + r_cur] = word & (word - 1u);  // the candidate is taken off
                                    isx = (t0 + src_lane) * 32 + bit - (int)PAD;
                                }
                            }
                            if (isx >= 0) break;
                            rowmask &= rowmask - 1u;  // the row is exhausted
                        }
                        if (isx < 0) break;  // tile-row exhausted
                        __syncwarp();
                        isy = row0 + r_cur;
                        sx = (float)isx; sy = (float)isy;
                        // the reference's suppression test against every accepted line (lsd.rs:84-89), lane = line
                        if (__any_sync(FULL, lane < n && suppressed_by(sh.lines[lane], sx, sy))) continue;
                        seeds++;
                        if (dead_ok) {
                            // deep inside a remembered dead cluster?  get_centre moves a seed by at most 2.5 px per axis, so
                            // the raw pixel decides when it is sqrt(12.5) closer.  Lane l tests entry l.
                            bool in = false;
                            if (lane < (int)min(n_dead, (uint32_t)L2_DEAD)) {
                                const float4 d = sh.dead[lane];
                                const float ex = sx - d.x, ey = sy - d.y;
                                in = ex * ex + ey * ey < d.w;
                            }
                            if (__any_sync(FULL, in)) { pruned++; continue; }
                        }
                        job = L2_JOB_SEED;
                        break;
                    }
                    if (lane == 0) { sh.jobs[par].job = job; sh.jobs[par].isx = isx; sh.jobs[par].isy = isy; sh.jobs[par].skip = 0; }
                    if (job == L2_JOB_SEED) {
                        // get_centre straight from the bitmap (the window is being loaded meanwhile) + second dead test
                        float px, py;
                        get_centre_warp(im, sx, sy, px, py);
                        bool in = false;
                        if (dead_ok && lane < (int)min(n_dead, (uint32_t)L2_DEAD)) {
                            const float4 d = sh.dead[lane];
                            const float ex = px - d.x, ey = py - d.y;
                            in = ex * ex + ey * ey < d.z;
                        }
                        const bool skip = __any_sync(FULL, in);
                        if (skip) pruned++;
                        if (lane == 0) { sh.jobs[par].px = px; sh.jobs[par].py = py; sh.jobs[par].skip = skip ? 1 : 0; }
                    }
                }
                __syncthreads();
                const typename Shared::Job jb = sh.jobs[par];  // (the master's next job goes to the other slot)
                par ^= 1;
                if (jb.job == L2_JOB_ROW_DONE) break;
                if (jb.skip) continue;
                const int isx = jb.isx, isy = jb.isy;
                const float px = jb.px, py = jb.py;
                // ---- the seed's neighbourhood: tiles of the bitmap, one coalesced load each ---------------
                Win2 wv;
                int ntx, nty;
                {
                    const int wx0 = max(isx - WR, 0), wx1 = min(isx + WR, W - 1), wy0 = max(isy - WR, 0), wy1 = min(isy + WR, H - 1);
                    wv.p = sh.win; wv.pad = PAD; wv.w = a.g.w; wv.h = a.g.h;
                    wv.tx0 = (wx0 + (int)PAD) >> 5; ntx = ((wx1 + (int)PAD) >> 5) - wv.tx0 + 1;
                    const int ty0 = wy0 >> 5;
                    wv.y0 = ty0 << 5; nty = (wy1 >> 5) - ty0 + 1;
                    for (int tyy = warp; tyy < nty; tyy += NW) {
                        const uint32_t* trow = im.p + ((uint64_t)(ty0 + tyy) * a.g.tx + (uint32_t)wv.tx0) * 32u + lane;
                        uint32_t tw[L2_WIN_T];
#pragma unroll
                        for (int txx = 0; txx < L2_WIN_T; txx++) tw[txx] = txx < ntx ? __ldg(trow + txx * 32) : 0u;
#pragma unroll
                        for (int txx = 0; txx < L2_WIN_T; txx++)
                            if (txx < ntx) sh.win[(tyy * 32 + lane) * L2_WIN_PITCH + txx] = tw[txx];
                    }
                }
                __syncthreads();
                // ---- exact ray filter: the warps list the white quarters of the (2R+1)^2 window around the centre, tile
                //      by tile, into one shared list (slots by warp-aggregated atomics: the order does not matter) ... -----
                const int fx = (int)px, fy = (int)py, R = F.radius;  // px, py >= 0
                const float hx = px - (float)fx, hy = py - (float)fy;  // 0 or 0.5
                {
                    const int x0 = max(fx - R, 0), x1 = min(fx + R, W - 1), y0 = max(fy - R, 0), y1 = min(fy + R, H - 1);
                    const int X0 = x0 + (int)PAD, X1 = x1 + (int)PAD;
                    const int ta = (y0 - wv.y0) >> 5, tb = (y1 - wv.y0) >> 5, ua = (X0 >> 5) - wv.tx0, ub = (X1 >> 5) - wv.tx0;
                    const int nu = ub - ua + 1, ntile = (tb - ta + 1) * nu;
                    const uint32_t rdiv = (uint32_t)((256 + nu - 1) / nu);
                    for (int ti = warp; ti < ntile; ti += NW) {
                        const int tq = (int)(((uint32_t)ti * rdiv) >> 8);  // ti / nu for ti < 64, nu <= 8
                        const int tyy = ta + tq, txx = ua + ti - tq * nu;
                        const int y = wv.y0 + (tyy << 5) + lane;
                        uint32_t w = (y >= y0 && y <= y1) ? sh.win[(tyy * 32 + lane) * L2_WIN_PITCH + txx] : 0u;
                        const int Xw = (wv.tx0 + txx) << 5;
                        if (Xw < X0) w &= 0xFFFFFFFFu << (X0 - Xw);
                        if (Xw + 31 > X1) w &= 0xFFFFFFFFu >> (Xw + 31 - X1);
                        // the tile's quarters, by byte position: 4 ballots, one atomic for the tile
                        uint32_t have[4], cnt = 0;
#pragma unroll
                        for (int q = 0; q < 4; q++) { have[q] = __ballot_sync(FULL, ((w >> (8 * q)) & 0xFFu) != 0u); cnt += (uint32_t)__popc(have[q]); }
                        if (cnt == 0u) continue;
                        uint32_t base = 0;
                        if (lane == 0) base = atomicAdd(&sh.n_grp, cnt);
                        base = __shfl_sync(FULL, base, 0);
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const uint32_t m8 = (w >> (8 * q)) & 0xFFu;
                            const uint32_t slot = base + (uint32_t)__popc(have[q] & LT);
                            if (m8 && slot < (uint32_t)L2_GRP_CAP)
                                sh.grp[slot] = m8 | ((uint32_t)(Xw + 8 * q - (int)PAD - fx + 160) << 8) | ((uint32_t)(y - fy + 128) << 17);
                            base += (uint32_t)__popc(have[q]);
                        }
                    }
                }
                __syncthreads();
                // ---- ... and mark the rays every quarter can serve, ring by ring: one quarter per lane, all warps ----------
                const uint32_t n_grp = sh.n_grp;
                const bool everything = n_grp > (uint32_t)L2_GRP_CAP;  // too many white pixels around: every ray is marched
                if (!everything) {
                    float near2 = 0.0f, far2 = 3.0e38f;
                    const float a2 = F.r2_lo[0];
                    for (uint32_t i = threadIdx.x; i < n_grp; i += NW * 32) {
                        const uint32_t e = sh.grp[i], m8 = e & 0xFFu;
                        const int ox = (int)((e >> 8) & 0x1FFu) - 160, oy = (int)(e >> 17) - 128;
                        const float cy = (float)oy + 0.5f - hy;
                        const float cxa = (float)(ox + __ffs(m8) - 1) + 0.5f - hx, cxb = (float)(ox + 31 - __clz((int)m8)) + 0.5f - hx;
                        float dmin2, dmax2;
                        int lo, hi;
                        smh_lsd_filter_quarter(cy, cxa, cxb, &dmin2, &dmax2, &lo, &hi);
                        if (dmax2 < a2) near2 = fmaxf(near2, dmax2);
                        else {
                            if (dmin2 < a2) near2 = fmaxf(near2, a2);  // straddles the inner disc's rim: no dead cluster from this seed
                            far2 = fminf(far2, fmaxf(dmin2, a2));
#pragma unroll
                            for (int j = 0; j < SMH_LSD_FILTER_MAX_WINDOWS; j++)
                                if (j < F.nw && dmax2 >= F.r2_lo[j] && dmin2 <= F.r2_hi[j]) l2_mark(sh.ring[j], lo, hi);
                        }
                    }
                    const uint32_t n2 = __reduce_max_sync(FULL, __float_as_uint(near2));  // positive floats order like their bits
                    const uint32_t f2 = __reduce_min_sync(FULL, __float_as_uint(far2));
                    if (lane == 0) {
                        if (n2) atomicMax(&sh.near2, n2);
                        if (f2 != __float_as_uint(3.0e38f)) atomicMin(&sh.far2, f2);
                    }
                }
                __syncthreads();
                // ---- AND of the usable rings, word by word = 32 consecutive rays = one lane-per-ray group -------
                L2Best best{0ull, px, py};
                bool any_ray = false;
                const int c0 = win_steps / 32;
                // (every lane takes one word: one pass over the rings, then only the non-empty words are visited.  A marker line
                // through the seed leaves rays in two opposite directions, 56 words apart: the words of the second half go to
                // the warps half a CTA further on, so that the four words of such a seed fall to four different warps)
                uint32_t vmine = 0u;
                constexpr int L2_HALF = SMH_LSD_BITMAP_WORDS / 2;  // 56
                static_assert(NW * 16 >= SMH_LSD_BITMAP_WORDS - L2_HALF, "sixteen lanes of every warp cover half the words");
                const int wmine = lane < 16 ? (warp + NW * lane < L2_HALF ? warp + NW * lane : SMH_LSD_BITMAP_WORDS)
                                            : L2_HALF + (warp + NW - NW / 2) % NW + NW * (lane - 16);
                {
                    if (wmine < SMH_LSD_BITMAP_WORDS) {
                        vmine = 0xFFFFFFFFu;
#pragma unroll
                        for (int j = 0; j < SMH_LSD_FILTER_MAX_WINDOWS; j++) {
                            if (!everything && j < F.nw && smh_lsd_filter_window_on(&F, j, px, py)) vmine &= sh.ring[j][wmine];
                            sh.ring[j][wmine] = 0u;  // cleared for the next seed
                        }
                        if (wmine == SMH_LSD_BITMAP_WORDS - 1) vmine &= (1u << (SMH_LSD_RAYS & 31)) - 1u;  // 3600 = 112 * 32 + 16
                    }
                }
                for (uint32_t nzw = __ballot_sync(FULL, vmine != 0u); nzw; nzw &= nzw - 1u) {
                    const int wl = __ffs(nzw) - 1, w = __shfl_sync(FULL, wmine, wl);
                    const uint32_t v = __shfl_sync(FULL, vmine, wl);
                    if (v == 0u) continue;
                    any_ray = true;
                    // lane = ray: the window's samples first, then 32 samples at a time on the bitmap while enough rays of the
                    // group live (a lane-per-ray step costs ~18 instructions per group, a ray-per-warp step ~2 per ray)
                    const int k = w * 32 + lane;
                    bool alive = ((v >> lane) & 1u) != 0u;
                    int gap = 0, c = c0;
                    float xo = 0.0f, yo = 0.0f, gx = px, gy = py;
                    float2 d = make_float2(0.0f, 0.0f);
                    if (alive) d = __ldg(&a.rays[k]);
                    auto finish = [&](int state) {  // the ray ended: its end point and length into the lane's best
                        float ex, ey, len;
                        if (state == 1) {  // the gap's first sample is inside the image and not white: lib.rs:426-429 holds
                            ex = __fsub_rn(gx, d.x); ey = __fsub_rn(gy, d.y);
                            const float ddx = __fsub_rn(px, ex), ddy = __fsub_rn(py, ey);
                            len = __fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy));
                        } else len = ray_finish(im, gx, gy, px, py, d.x, d.y, ex, ey);
                        const unsigned long long key = ((unsigned long long)__float_as_uint(len) << 32) | (uint32_t)k;
                        if (key >= best.key) { best.key = key; best.ex = ex; best.ey = ey; }
                    };
                    if (alive && win_steps > 0) {
                        const int state = l2_march_lanes<STATS>(wv, win_steps, px, py, a.gap_limit, d.x, d.y, xo, yo, gx, gy, gap, steps);
                        if (state) { finish(state); alive = false; }
                    }
                    uint32_t am = __ballot_sync(FULL, alive);
                    while (win_steps > 0 && __popc(am) >= L2_LANES_MIN) {
                        if (alive) {
                            const int state = l2_march_lanes<STATS>(im, 32, px, py, a.gap_limit, d.x, d.y, xo, yo, gx, gy, gap, steps);
                            if (state) { finish(state); alive = false; }
                        }
                        c++;
                        am = __ballot_sync(FULL, alive);
                    }
                    if (am) {  // survivors continue one ray per warp: through the shared list, so that all warps share them
                        uint32_t slot = 0;
                        if (lane == 0) slot = atomicAdd(&sh.n_surv, (uint32_t)__popc(am));
                        slot = __shfl_sync(FULL, slot, 0);
                        const uint32_t mine = slot + (uint32_t)__popc(am & LT);
                        const bool fits = slot + (uint32_t)__popc(am) <= (uint32_t)L2_SURV_CAP;  // uniform
                        if (fits) { if (alive) sh.surv[mine] = l2_pack(k, gap, c); }
                        else {
                            if (lane == 0 && slot <= (uint32_t)L2_SURV_CAP) sh.surv_end = slot;  // the list is whole up to here
                            for (uint32_t todo = am; todo; todo &= todo - 1u) {  // list full: march them here and now
                                const int sl = __ffs(todo) - 1;
                                const uint32_t e = __shfl_sync(FULL, l2_pack(k, gap, c), sl);
                                l2_march_long<STATS, (NW >= 8)>(im, a, e, px, py, best, steps_u);
                            }
                        }
                    }
                }
                const bool some_ray = __syncthreads_or(any_ray);
                if (!some_ray) {  // no ray can exceed the minimum length
                    if (master) {
                        pruned++;
                        // every white pixel is within sqrt(near2) of the seed or beyond dfar: a seed at distance t < rho sees
                        // none of them in its first ring either
                        const float near2 = __uint_as_float(sh.near2), far2 = everything ? 0.0f : __uint_as_float(sh.far2);
                        const float k1 = (float)F.k1;
                        const float dfar = fminf(sqrtf(far2), (float)(F.radius - 1));
                        const float rho = fminf(k1 - 0.75f - sqrtf(near2), dfar - (2.0f * k1 - 0.25f)) - 0.05f;
                        if (dead_ok && rho > 0.5f) {
                            const float raw = fmaxf(rho - 3.6f, 0.0f);  // radius for raw (un-centred) seed pixels
                            if (lane == 0) sh.dead[n_dead % (uint32_t)L2_DEAD] = make_float4(px, py, rho * rho, raw * raw);
                            n_dead++;
                            __syncwarp();
                            if (!STATS && raw > 0.0f) {  // off this tile-row right away (only the master touches the chunk here)
                                l2_strike_disc<false>(chunk, PAD, px, py, raw * raw, row0, W);
                                __syncwarp();
                                uint32_t nzr = 0;
                                for (int t = 0; t < TX; t++) nzr |= chunk[t * 33 + lane];
                                rowmask &= __ballot_sync(FULL, nzr != 0u);
                            }
                        }
                        __syncwarp();  // every lane has read near2 / far2
                        if (lane == 0) { sh.near2 = 0u; sh.far2 = __float_as_uint(3.0e38f); sh.n_dead = n_dead; sh.n_grp = 0u; }
                    }
                    continue;  // the master's writes are published by the next barrier
                }
                // ---- survivors: one ray per warp, shared over the warps -----------------------------------
                {
                    const uint32_t ns = min(sh.n_surv, sh.surv_end);
                    for (uint32_t j = warp; j < ns; j += NW) l2_march_long<STATS, (NW >= 8)>(im, a, sh.surv[j], px, py, best, steps_u);
                }
                // warp argmax of (len, k): the LAST index among the longest rays (D6)
                {
                    unsigned long long wkey = best.key;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const unsigned long long other = __shfl_xor_sync(FULL, wkey, o);
                        wkey = other > wkey ? other : wkey;
                    }
                    const int srcl = __ffs(__ballot_sync(FULL, best.key == wkey)) - 1;
                    const float bex = __shfl_sync(FULL, best.ex, srcl), bey = __shfl_sync(FULL, best.ey, srcl);
                    if (lane == 0) { sh.bkey[warp] = wkey; sh.bend[warp] = make_float2(bex, bey); }
                }
                __syncthreads();
                if (master) {
                    int bw = 0;
                    unsigned long long bk = sh.bkey[0];
#pragma unroll
                    for (int k2 = 1; k2 < NW; k2++) { const unsigned long long o = sh.bkey[k2]; if (o > bk) { bk = o; bw = k2; } }
                    const float blen = __uint_as_float((uint32_t)(bk >> 32));
                    if (blen > SMH_LSD_MIN_LEN_SQ) {
                        float c1x, c1y;
                        get_centre_warp(im, sh.bend[bw].x, sh.bend[bw].y, c1x, c1y);
                        const float4 line = make_float4(px, py, c1x, c1y);
                        if (lane == 0) {
                            sh.lines[n] = line;
                            out_lines[4 * n + 0] = px; out_lines[4 * n + 1] = py;
                            out_lines[4 * n + 2] = c1x; out_lines[4 * n + 3] = c1y;
                            sh.n = n + 1;
                        }
                        n++;
                        __syncwarp();
                        if (n < SMH_LSD_MAX_LINES) {  // suppressed for good: strike the band off the rest of this tile-row
                            l2_strike_line<false>(chunk, PAD, line, row0, r_cur, W);
                            __syncwarp();
                            uint32_t nzr = 0;
                            for (int t = 0; t < TX; t++) nzr |= chunk[t * 33 + lane];
                            rowmask &= __ballot_sync(FULL, nzr != 0u);
                        }
                    }
                    if (lane == 0) { sh.n_surv = 0u; sh.surv_end = (uint32_t)L2_SURV_CAP; sh.near2 = 0u; sh.far2 = __float_as_uint(3.0e38f); sh.n_grp = 0u; }
                }
                // no barrier here: the next thing the workers read (sh.job) is published by the barrier at the loop's head
            }
            __syncthreads();  // the master's last writes (sh.n, dead clusters) before the next tile-row's strike-off
        }
        if (STATS) {
            steps = __reduce_add_sync(FULL, steps) + steps_u;
            if (lane == 0) atomicAdd(&sh.steps, (unsigned long long)steps);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            a.n_lines[f] = (uint32_t)sh.n;
            if (STATS) { a.stats[3 * f] = seeds; a.stats[3 * f + 1] = sh.steps; a.stats[3 * f + 2] = pruned; }
        }
    }
}

}  // namespace smh
