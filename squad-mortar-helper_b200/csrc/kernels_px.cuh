// kernels_px.cuh — k_pixels2: the fused pixel pass with TMA tile movement (sm_100a).
//
// One read of the captured BGRA map ROI produces ui_map, the dilated marker mask (lsd_image) AND its 1-bit
// tile-major copy for the line search, ocr_out and scales, plus the close-deployment-button test.  Replaces
// reference kernels K1-K9 (vision-gpu/cuda/cuda.cu:312-546,616-635 + NPP dilate) and restates the vision-cpu
// arithmetic (vision-cpu/src/lib.rs:110-280,357-375) bit for bit (vec_plain / vec_quadrant in kernels.cuh).
//
// Warp-autonomous, no block barrier.  A work item is (frame, band of 64 map rows — 16 when the batch is too small to
// fill the SMs —, strip of 128 columns = 4 bitmap tiles) or a frame's button test.  Columns are PADDED coordinates X = x + pad,
// pad = map_x & 3: the map ROI starts at an arbitrary pixel column of the capture (1219 at 1440p) and a TMA box
// must start on a 16-byte boundary (measured: tools/tma_probe.cu — an odd pixel start raises an illegal
// instruction), so every plane carries the capture's 16-byte phase as a left pad of 0..3 pixels and strip,
// vector and bitmap-word boundaries coincide in all of them.  Lane l owns the 4 pixels X = 128 s + 4 l .. of
// every row: 512 contiguous, 128-byte aligned bytes of ui_map and 128 of lsd_image per warp row.  An elected lane
// issues one cp.async.bulk.tensor per chunk of two rows (box = 128 pixels x 2 rows: whole 128-byte lines when the
// map starts on one, pixels outside the map's columns / rows are not fetched and arrive as zeros) into a
// three-stage ring of the warp's own shared memory and the lanes wait on the stage's mbarrier; nobody computes
// an address or commits / waits for a copy of its own.
//   phase 1 (row jn)      classify the lane's 4 pixels, store ui_map / provisional ocr_out / scales, shift the
//                         row's flag nibbles (marker, OCR core, OCR edge) into window registers.
//   phase 2 (row jn - lag) 3x3-cross dilation (up/down from the window, left/right by shuffle) -> lsd_image bytes
//                         and, gathered by a 3-step butterfly, the four 32-pixel words of the bitmap; the 7x7
//                         neighbour search of the rare OCR edge pixels.  lag = 1 outside the OCR quadrant, 3 in it.
// No warp reads a pixel of another strip.  What the 3x3 cross and the 7x7 window need from across a seam between two
// strips travels as one flag byte per (strip, side, row), stored only where nonzero (Pixel2Args::edge), and
// k_px_seams — one thread per seam and four rows, almost all of which read two zero words — finishes the pixels
// either side.  Round 2 measured the alternatives on the 1024-frame batch (profiles/README.md): a box of 136 pixels
// (128 + a 4-pixel halo either side, 3.02 ms), second and third 4-pixel boxes (3.04 ms), cp.async halos (3.15 ms),
// halos out of the neighbouring warp's ring with the warps of a CTA in lockstep (3.3 - 3.6 ms); this form 2.88 ms.
#pragma once
#include <cuda.h>

#include "kernels.cuh"
#include "smh_px_plan.h"

namespace smh {

struct Pixel2Args {
    const uint8_t* frames;  // captured BGRA frames; ROI coordinates below are relative to this origin
    uint64_t frame_stride;
    uint32_t pitch;
    uint32_t map_x, map_y, w, h, bqw, bqh;
    uint32_t pad;           // map_x & 3: plane column X = x + pad
    uint32_t btn_x, btn_y, btn_w, btn_h;
    uint8_t* ui; uint64_t ui_stride; uint32_t ui_pitch;
    uint8_t* lsd; uint64_t lsd_stride; uint32_t lsd_pitch;
    uint8_t* ocr; uint8_t* scales; uint64_t q_stride; uint32_t q_pitch;
    uint32_t q_pad;         // left pad of the quadrant planes: (pad + bqw) & 127 (a strip's quadrant bytes start 128-byte aligned)
    const uint32_t* start_y;  // device [n] or nullptr
    uint32_t start_y_uniform;
    uint32_t* red; uint8_t* open;
    uint32_t n_frames;
    smh_px_plan plan;       // how the map is cut into items (smh_px_plan.h)
    uint32_t blocks_q;      // CTAs of QUAD items in a merged launch
    const uint32_t* lut;    // marker predicate, 1 bit per colour (r<<16|g<<8|b)
    uint32_t* bits; uint64_t bits_stride; uint32_t bits_tx;  // tile-major 1-bit copy of the dilated mask
    // What a strip's outermost columns tell the strip next door, [frame][strip][side 0 = first columns, 1 = last][edge_hp
    // rows] bytes, cleared before the pass and stored only where a warp found something; k_px_seams reads them:
    //   bit 0      the first / last pixel of the row is a marker pixel
    //   bits 1-3   OCR core flags of the first / last three pixels (in column order)
    //   bits 4-6   OCR edge pixels among them whose 7x7 window found no core pixel inside the strip
    uint8_t* edge; uint64_t edge_stride; uint32_t edge_hp;
};

#ifndef PX2_WARPS_N
#define PX2_WARPS_N 4
#endif
constexpr int PX2_WARPS = PX2_WARPS_N;           // warps (= items) per CTA; 32 resident warps per SM either way
constexpr int PX2_THREADS = PX2_WARPS * 32;
#ifndef PX2_MIN_CTAS
#define PX2_MIN_CTAS (32 / PX2_WARPS)   // resident CTAs per SM the register allocation aims at
#endif
#ifndef PX2_STAGES_N
#define PX2_STAGES_N 3
#endif
#ifndef PX2_ROWS_N
#define PX2_ROWS_N 2
#endif
constexpr int PX2_STAGES = PX2_STAGES_N;
constexpr int PX2_ROWS = PX2_ROWS_N;             // rows per TMA box
constexpr int PX2_BOX_W = 128;                   // pixels per row of a strip box: 512 bytes, whole 128-byte lines when the
constexpr int PX2_ROW_BYTES = PX2_BOX_W * 4;     //   capture's map starts on one
constexpr int PX2_STAGE_BYTES = PX2_ROWS * PX2_ROW_BYTES;
constexpr int PX2_STAGE_TX = PX2_STAGE_BYTES;    // bytes the bulk copy of a stage delivers
constexpr int PX2_WARP_BYTES = PX2_STAGES * PX2_STAGE_BYTES + 128;  // ring + mbarriers
constexpr int PX2_SMEM_BYTES = PX2_WARPS * PX2_WARP_BYTES;

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int x, int y, int z, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}

// one lane of the (converged) warp
__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(p));
    return p != 0u;
}

// nibble -> 4 bytes of 0/255
__device__ __forceinline__ uint32_t nibble_bytes(uint32_t d4) { return ((d4 * 0x00204081u) & 0x01010101u) * 255u; }

template <bool B> struct PxTag { static constexpr bool value = B; };

template <bool GRAY, bool QUAD>
__device__ __forceinline__ void px2_item(const CUtensorMap* tmap, const Pixel2Args& a, uint32_t f, int r0, int r1, uint32_t strip,
                                         uint32_t ring, uint32_t bars) {
    constexpr int LAG = QUAD ? SMH_OCR_RADIUS : 1;   // rows phase 2 runs behind phase 1 (= rows of halo above / below)
    const int lane = threadIdx.x & 31;
    const uint32_t FULL = 0xFFFFFFFFu;
    const int w = (int)a.w, h = (int)a.h, bqw = (int)a.bqw, bqh = (int)a.bqh;
    const int x_base = (int)strip * 128 - (int)a.pad;  // map column of the strip's first pixel (X = 128 strip)
    const int i_lo = x_base + 4 * lane;              // map column of the lane's first pixel
    const bool active = i_lo < w;                     // the lane's vector starts left of the map's right edge: it stores
    const uint8_t* frame = a.frames + (uint64_t)f * a.frame_stride;
    const uint32_t* __restrict__ lut = a.lut;
    const uint32_t alpha = opaque_alpha();
    const uint32_t start_y = a.start_y ? a.start_y[f] : a.start_y_uniform;

    // ---- per-lane column masks: pixel in the image / may carry an OCR core bit / an OCR edge bit -----------
    uint32_t vm4 = 0, cm4 = 0, em4 = 0;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const int i = i_lo + e;
        const bool valid = i >= 0 && i < w;
        const bool inq = QUAD && valid && i >= bqw && i < 2 * bqw;
        vm4 |= valid ? (1u << e) : 0u;
        em4 |= inq ? (1u << e) : 0u;
        // an OCR core pixel is visible to a neighbour's search only below the clip min(x+3, w-3) (D10)
        cm4 |= (inq && i - bqw <= bqw - SMH_OCR_RADIUS) ? (1u << e) : 0u;
    }
    // Nobody reads a pixel of another strip: the first / last columns' flags go to the neighbour through Pixel2Args::edge
    // (rare stores) and k_px_seams finishes the pixels either side of a seam.  A halo read of the neighbour's line —
    // as a wider box, a second box, cp.async or a plain load — cost the pass a quarter of its time (measured).
    const bool edge = lane == 0 || lane == 31;
    const uint32_t own_off = (uint32_t)lane * 16u;  // the lane's vector in a strip row
    const bool col_brq = em4 != 0u;
    const uint32_t qpos = (uint32_t)(i_lo - bqw + (int)a.q_pad);  // byte position in a quadrant row (used when col_brq)

    // ---- rows: read [j_first, j_end), owned (stored) [r0, r1) -----------------------------------------------
    const int j_first = r0 - LAG, j_end = r1 + LAG;
    const int n_chunks = (j_end - j_first + PX2_ROWS - 1) / PX2_ROWS;
    const int tx = (int)a.map_x + x_base, ty = (int)a.map_y + j_first;

    const uint32_t X_lo = strip * 128u + 4u * (uint32_t)lane;  // plane column of the lane's first pixel
    uint8_t* ui_p = a.ui + (uint64_t)f * a.ui_stride + (int64_t)j_first * (int64_t)a.ui_pitch + X_lo * 4u;       // row jn
    uint8_t* lsd_p = a.lsd + (uint64_t)f * a.lsd_stride + (int64_t)(j_first - LAG) * (int64_t)a.lsd_pitch + X_lo;  // row jn - LAG
    uint8_t* ocr_f = a.ocr + (uint64_t)f * a.q_stride;
    // the lane's word of quadrant row jn - bqh in ocr_out (the same offset from a.scales is its word in scales); outside
    // the quadrant's rows / columns the pointer is never dereferenced
    uint8_t* ocr_p = ocr_f + ((int64_t)(j_first - bqh) * (int64_t)a.q_pitch + (int64_t)(int32_t)qpos);
    const int64_t sc_delta = a.scales - a.ocr;
    // bitmap: word of tile (jc >> 5, 4 strip + lane / 8), row jc & 31.  The eight lanes of a group all hold the group's
    // word of a row; lane 8 g + k keeps row (jc & ~7) + k's and every eighth row the warp stores four whole 32-byte
    // sectors (a 4-byte store per row and tile cost the memory system far more than its bytes: tools/tile_probe2.cu)
    const uint32_t btile = 4u * strip + (uint32_t)(lane >> 3);
    uint32_t* const bits_f = a.bits + (uint64_t)f * a.bits_stride;
    uint32_t bits_keep = 0;

    uint32_t mwin = 0, cwin = 0, ewin = 0;  // flag nibbles of rows jn, jn-1, ... (newest in bits 0-3)
    // lanes 0 / 31: this strip's flag bytes for its neighbour (side 0 / 1), row 0
    uint8_t* const edge_base = a.edge + (uint64_t)f * a.edge_stride + (uint64_t)(2u * strip + (lane == 31 ? 1u : 0u)) * a.edge_hp;
    // the window bits of row jc that the neighbour needs: lane 0's first three pixels, lane 31's last three
    constexpr uint32_t NIB = 4 * (QUAD ? SMH_OCR_RADIUS : 1);
    const uint32_t seam_mask = lane == 0 ? (0x7u << NIB) : (lane == 31 ? (0xEu << NIB) : 0u);

    // ---- prologue: the first PX2_STAGES chunks are requested (every operand of the copy is warp-uniform) --------
    if (elect_one()) {
#pragma unroll
        for (int c = 0; c < PX2_STAGES; c++) {
            if (c < n_chunks) {
                mbar_expect_tx(bars + 8u * c, PX2_STAGE_TX);
                tma_load_3d(ring + (uint32_t)c * PX2_STAGE_BYTES, tmap, tx, ty + PX2_ROWS * c, (int)f, bars + 8u * c);
            }
        }
    }

    // FAST rows (the bulk): row jn and row jn - LAG are both owned (so inside the map), and in a QUAD item both are rows
    // of the quadrant; the generic form tests all of that per row.
    auto row = [&](auto tag, uint32_t srow, int jn) {
        constexpr bool FAST = decltype(tag)::value;
        // ---------------- phase 1: row jn ---------------------------------------------------------------------
        uint32_t m4 = 0, c4 = 0, e4 = 0;
        if (FAST || (jn >= 0 && jn < h)) {  // uniform
            const uint4 cur = lds_v4(srow + own_off);
            const bool owned = FAST || (jn >= r0 && jn < r1);
            const int jq = jn - bqh;
            if (QUAD && (FAST || (uint32_t)jq < (uint32_t)bqh)) {  // uniform: a row of the OCR quadrant
                // Every lane tests its own vector's OR against the cheap necessary conditions (a marker colour has a channel
                // >= 178; both OCR classes need every channel >= 130, i.e. bit 7 of all three bytes; scales needs a pixel of
                // luma 0) and classifies pixel by pixel only then: almost every vector of a map passes straight through.
                const uint32_t any = or4(cur);
                uint32_t lb[4] = {BIAS + 1u, BIAS + 1u, BIAS + 1u, BIAS + 1u};  // biased luma words 0x4B0000LL (owned rows)
                if (owned) { lb[0] = luma_biased_i2f(cur.x); lb[1] = luma_biased_i2f(cur.y); lb[2] = luma_biased_i2f(cur.z); lb[3] = luma_biased_i2f(cur.w); }
                uint32_t ocrw = 0xFFFFFFFFu, scw = 0xFFFFFFFFu;
                if (bright_bytes(any)) m4 = marker_nibble(cur, lut);  // rare
                if (col_brq) {
                    if ((any & 0x00808080u) == 0x00808080u) {  // rare: possibly a bright near-grey pixel (vision-cpu/src/lib.rs:39-53,188-193)
                        // (the vector is read again: a volatile load keeps the byte extraction below inside this branch)
                        const uint4 again = lds_v4(srow + own_off);
                        const uint32_t wd[4] = {again.x, again.y, again.z, again.w};
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const uint32_t cb = wd[e] & 0xFFu, cg = (wd[e] >> 8) & 0xFFu, cr = (wd[e] >> 16) & 0xFFu;
                            const uint32_t mx = max(cr, max(cg, cb)), mn = min(cr, min(cg, cb));
                            if (mx == mn && mn >= SMH_OCR_BRIGHT) {
                                c4 |= 1u << e;
                                ocrw ^= (lb[e] & 0xFFu) << (8 * e);  // 255 - luma
                            } else if (4u * (mx - mn) <= SMH_OCR_MONO_EDGE && mn >= SMH_OCR_BRIGHT_EDGE) e4 |= 1u << e;
                        }
                        c4 = jq <= bqh - SMH_OCR_RADIUS ? c4 & cm4 : 0u;  // rows above the vertical clip only (D10)
                        e4 &= em4;
                    }
                    if (min(min(lb[0], lb[1]), min(lb[2], lb[3])) == BIAS) {  // rare: a pixel with luma 0
#pragma unroll
                        for (int e = 0; e < 4; e++)
                            if (lb[e] == BIAS) scw &= ~(0xFFu << (8 * e));
                    }
                }
                if (owned && active) {
                    uint4 ui;
                    if (GRAY) {
                        ui.x = lb[0] * 0x010101u + 0xB4000000u; ui.y = lb[1] * 0x010101u + 0xB4000000u;
                        ui.z = lb[2] * 0x010101u + 0xB4000000u; ui.w = lb[3] * 0x010101u + 0xB4000000u;
                    } else {
                        ui.x = __byte_perm(cur.x, alpha, 0x7012); ui.y = __byte_perm(cur.y, alpha, 0x7012);
                        ui.z = __byte_perm(cur.z, alpha, 0x7012); ui.w = __byte_perm(cur.w, alpha, 0x7012);
                    }
                    stg_v4(ui_p, ui);
                    if (col_brq) {
                        stg_u32(ocr_p, ocrw);
                        stg_u32(ocr_p + sc_delta, (uint32_t)jq >= start_y ? scw : 0u);  // rows above start_y: zero (D9, a fresh buffer)
                    }
                }
            } else {
                if (owned) {
                    uint4 ui;
                    if (GRAY) {  // (L, L, L, 255) = 0x4B0000LL * 0x010101 + 0xB4000000 (mod 2^32): one IMAD per pixel
                        ui.x = luma_biased_i2f(cur.x) * 0x010101u + 0xB4000000u; ui.y = luma_biased_i2f(cur.y) * 0x010101u + 0xB4000000u;
                        ui.z = luma_biased_i2f(cur.z) * 0x010101u + 0xB4000000u; ui.w = luma_biased_i2f(cur.w) * 0x010101u + 0xB4000000u;
                    } else {
                        ui.x = __byte_perm(cur.x, alpha, 0x7012); ui.y = __byte_perm(cur.y, alpha, 0x7012);
                        ui.z = __byte_perm(cur.z, alpha, 0x7012); ui.w = __byte_perm(cur.w, alpha, 0x7012);
                    }
                    if (active) stg_v4(ui_p, ui);
                }
                if (bright_bytes(or4(cur))) m4 = marker_nibble(cur, lut);  // rare
            }
            m4 &= vm4;
        }
        mwin = (mwin << 4) | m4;
        if (QUAD) { cwin = (cwin << 4) | c4; ewin = (ewin << 4) | e4; }
        ui_p += a.ui_pitch;
        if (QUAD) ocr_p += a.q_pitch;

        // ---------------- phase 2: row jc = jn - LAG ----------------------------------------------------------
        const int jc = jn - LAG;
        uint8_t* const lsd_row = lsd_p;  // row jc
        lsd_p += a.lsd_pitch;
        if (!FAST && (jc < r0 || jc >= r1)) return;  // uniform
        // rows jc-1, jc, jc+1 sit at nibbles LAG+1, LAG, LAG-1 of the windows
        constexpr uint32_t ROWS3 = 0xFFFu << (4 * (LAG - 1));
        uint32_t fl = 0;  // lanes 0 / 31: what the neighbouring strip must know about this row (Pixel2Args::edge)
        uint32_t v = 0;   // the bitmap word of the lane's group of eight
        if (__any_sync(FULL, (mwin & ROWS3) != 0u)) {  // (almost every row of a map has no marker pixel nearby)
            const uint32_t mc = (mwin >> (4 * LAG)) & 0xFu;
            uint32_t l = __shfl_up_sync(FULL, mc, 1), rr = __shfl_down_sync(FULL, mc, 1);
            if (lane == 0) l = 0u;    // the pixel across the seam: k_px_seams
            if (lane == 31) rr = 0u;
            const uint32_t d4 = (mc | (mc << 1) | (mc >> 1) | (l >> 3) | (rr << 3) | (mwin >> (4 * (LAG + 1))) | (mwin >> (4 * (LAG - 1)))) & vm4;
            v = d4 << (4 * (lane & 7));
            v |= __shfl_xor_sync(FULL, v, 1);
            v |= __shfl_xor_sync(FULL, v, 2);
            v |= __shfl_xor_sync(FULL, v, 4);
            if (active) stg_u32(lsd_row, nibble_bytes(d4));
            if (edge) fl = lane == 0 ? (mc & 1u) : (mc >> 3);
        } else {
            if (active) stg_u32(lsd_row, 0u);
        }
        if ((lane & 7) == (jc & 7)) bits_keep = v;
        if ((jc & 7) == 7 || jc == r1 - 1) {  // uniform: the rows (jc & ~7) .. jc of the four tiles
            const int j = (jc & ~7) + (lane & 7);
            if (j <= jc && btile < a.bits_tx) bits_f[((((uint32_t)j >> 5) * a.bits_tx + btile) << 5) + ((uint32_t)j & 31u)] = bits_keep;
        }

        const int jqc = jc - bqh;
        if (QUAD && (FAST || (uint32_t)jqc < (uint32_t)bqh)) {  // uniform: a row of the OCR quadrant
            constexpr uint32_t ROW1 = 0xFu << (4 * LAG);
            if (__any_sync(FULL, ((ewin & ROW1) | (cwin & seam_mask)) != 0u)) {
                // rare: an OCR edge pixel keeps its value iff a core pixel lies in its 7x7 window
                // (vision-cpu/src/lib.rs:196-216); rows/columns outside the quadrant carry no core bit.  The part of a
                // window that lies in the neighbouring strip is k_px_seams' business: it gets this strip's outermost
                // core bits and the edge pixels still looking for one.
                const uint32_t eb = (ewin >> (4 * LAG)) & 0xFu;
                uint32_t lwin = __shfl_up_sync(FULL, cwin, 1), rwin = __shfl_down_sync(FULL, cwin, 1);
                if (lane == 0) lwin = 0u;
                if (lane == 31) rwin = 0u;
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
                if (edge) {
                    const uint32_t cc = (cwin >> (4 * LAG)) & 0xFu, pend = eb & ~found;
                    fl |= lane == 0 ? ((cc & 7u) << 1) | ((pend & 7u) << 4) : ((cc >> 1) << 1) | ((pend >> 1) << 4);
                }
            }
        }
        if (fl) edge_base[jc] = (uint8_t)fl;
    };

    // ---- main loop over the chunks of PX2_ROWS rows -------------------------------------------------------------
    uint32_t stage = 0, phase = 0;
#pragma unroll 1
    for (int c = 0; c < n_chunks; c++) {
        mbar_wait(bars + 8u * stage, phase);
        const uint32_t sbase = ring + stage * PX2_STAGE_BYTES;
        const int jn = j_first + PX2_ROWS * c;
        // every row of the chunk and the rows LAG above them owned; in a QUAD item all of them rows of the quadrant
        const bool fast = jn - LAG >= r0 && jn + PX2_ROWS - 1 < r1 && (!QUAD || (jn - LAG >= bqh && jn + PX2_ROWS - 1 < 2 * bqh));
        if (fast) {
#pragma unroll 1
            for (int rr = 0; rr < PX2_ROWS; rr++)
                row(PxTag<true>(), sbase + (uint32_t)rr * PX2_ROW_BYTES, jn + rr);
        } else {
#pragma unroll 1
            for (int rr = 0; rr < PX2_ROWS; rr++)
                if (jn + rr < j_end)
                    row(PxTag<false>(), sbase + (uint32_t)rr * PX2_ROW_BYTES, jn + rr);
        }
        __syncwarp();  // every lane has read the stage: it may be refilled
        if (c + PX2_STAGES < n_chunks && elect_one()) {
            mbar_expect_tx(bars + 8u * stage, PX2_STAGE_TX);
            tma_load_3d(sbase, tmap, tx, ty + PX2_ROWS * (c + PX2_STAGES), (int)f, bars + 8u * stage);
        }
        if (++stage == PX2_STAGES) { stage = 0; phase ^= 1u; }
    }
}

// One launch: the items that own pixels of the OCR quadrant (QUAD: three rows of halo, OCR classes, the 7x7 search —
// bound by instruction issue) are spread evenly among the plain ones (bound by HBM) and the button tests, CTA by CTA,
// so that every SM holds both kinds all the time (MODE 2; 2.72 ms against 2.82 ms as two launches and 2.77 ms as one
// launch with the QUAD CTAs first).  MODE 0 / 1: one kind per launch (SMH_B200_PX_MERGED=0).
template <bool GRAY, int MODE>
__global__ void __launch_bounds__(PX2_THREADS, PX2_MIN_CTAS) k_pixels2(const __grid_constant__ CUtensorMap tmap, const Pixel2Args a) {
    extern __shared__ __align__(128) uint8_t px2_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bool quad = MODE == 1;
    uint32_t blk = blockIdx.x;
    if (MODE == 2) {  // block b is the q-th QUAD CTA if the count of QUAD CTAs below b + 1 exceeds the count below b
        const uint32_t q0 = (uint32_t)((uint64_t)blockIdx.x * a.blocks_q / gridDim.x), q1 = (uint32_t)(((uint64_t)blockIdx.x + 1u) * a.blocks_q / gridDim.x);
        quad = q1 > q0;
        blk = quad ? q0 : blockIdx.x - q0;
    }
    const uint32_t item = blk * PX2_WARPS + warp;
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(px2_smem) + (uint32_t)warp * PX2_WARP_BYTES;
    const uint32_t bars = ring + PX2_STAGES * PX2_STAGE_BYTES;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < PX2_STAGES; s++) mbar_init(bars + 8u * s, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t f, strip;
    int r0, r1;
    if (quad) {
        const uint32_t per_frame = smh_px_plan_quad_items(&a.plan);
        f = item / per_frame;
        if (f >= a.n_frames) return;
        smh_px_plan_quad_item(&a.plan, item - f * per_frame, &strip, &r0, &r1);
        if (MODE != 0) px2_item<GRAY, true>(&tmap, a, f, r0, r1, strip, ring, bars);
    } else {
        const uint32_t n_plain = smh_px_plan_plain_items(&a.plan), per_frame = n_plain + 1u;
        f = item / per_frame;
        if (f >= a.n_frames) return;
        const uint32_t sub = item - f * per_frame;
        if (sub == n_plain) {
            // ---- close-deployment button test (vision-cpu/src/lib.rs:115-135) -----------------
            const uint8_t* frame = a.frames + (uint64_t)f * a.frame_stride;
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
        smh_px_plan_plain_item(&a.plan, sub, &strip, &r0, &r1);
        if (MODE != 1) px2_item<GRAY, false>(&tmap, a, f, r0, r1, strip, ring, bars);
    }
}

// The pixels either side of a seam between two strips, from the flags the strips left in Pixel2Args::edge: the 3x3
// cross of a marker pixel reaches one column across (dilated mask + its bitmap), the 7x7 window of an OCR edge pixel
// three (vision-cpu/src/lib.rs:196-216).  One thread per (frame, seam, four rows); almost all of them read two zero
// words and leave.
__global__ void __launch_bounds__(256) k_px_seams(const Pixel2Args a) {
    const uint32_t nrq = a.edge_hp >> 2, nseams = a.plan.nstrips - 1u;
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t rq = (uint32_t)(idx % nrq);
    const uint64_t t = idx / nrq;
    const uint32_t b = 1u + (uint32_t)(t % nseams), f = (uint32_t)(t / nseams);  // the seam left of strip b
    if (f >= a.n_frames) return;
    const uint8_t* ef = a.edge + (uint64_t)f * a.edge_stride;
    const uint8_t* L = ef + (uint64_t)(2u * b - 1u) * a.edge_hp;  // strip b - 1, its last columns
    const uint8_t* R = ef + (uint64_t)(2u * b) * a.edge_hp;       // strip b, its first columns
    const uint32_t lw = *reinterpret_cast<const uint32_t*>(L + 4u * rq), rw = *reinterpret_cast<const uint32_t*>(R + 4u * rq);
    if (!(lw | rw)) return;
    const uint32_t X0 = 128u * b;  // plane column of strip b's first pixel
    const int bqw = (int)a.bqw, bqh = (int)a.bqh, h = (int)a.h, w = (int)a.w;
    for (int k = 0; k < 4; k++) {
        const int j = (int)(4u * rq) + k;
        if (j >= h) break;
        const uint32_t lb = (lw >> (8 * k)) & 0xFFu, rb = (rw >> (8 * k)) & 0xFFu;
        // ---- dilation across the seam: a marker pixel in one strip's outermost column sets its neighbour's
#pragma unroll
        for (int side = 0; side < 2; side++) {
            if (!((side ? rb : lb) & 1u)) continue;
            const uint32_t X = side ? X0 - 1u : X0;
            const int i = (int)X - (int)a.pad;
            if (i < 0 || i >= w) continue;
            a.lsd[(uint64_t)f * a.lsd_stride + (uint64_t)j * a.lsd_pitch + X] = 255;
            atomicOr(a.bits + (uint64_t)f * a.bits_stride + ((((uint64_t)(j >> 5) * a.bits_tx + (X >> 5)) << 5) + (uint32_t)(j & 31)), 1u << (X & 31u));
        }
        // ---- OCR edge pixels that found no core pixel inside their own strip: the other strip's part of the window
        const uint32_t pl = (lb >> 4) & 7u, pr = (rb >> 4) & 7u;  // pixels X0 - 3 .. X0 - 1 / X0 .. X0 + 2
        if (!(pl | pr)) continue;
        uint32_t cl = 0, cr = 0;  // core flags of those six columns over the window's rows (quadrant rows only carry any)
        for (int jj = max(j - SMH_OCR_RADIUS, bqh); jj <= min(j + SMH_OCR_RADIUS, min(2 * bqh, h) - 1); jj++) {
            cl |= ((uint32_t)L[jj] >> 1) & 7u;
            cr |= ((uint32_t)R[jj] >> 1) & 7u;
        }
        uint32_t fix = 0;  // bit c: pixel X0 - 3 + c keeps its value
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (((pl >> c) & 1u) && (cr & ((2u << c) - 1u))) fix |= 1u << c;               // sees strip b's columns 0 .. c
            if (((pr >> c) & 1u) && (cl & (7u & ~((1u << c) - 1u)))) fix |= 8u << c;      // sees strip b - 1's last 3 - c
        }
        for (int c = 0; c < 6; c++) {
            if (!((fix >> c) & 1u)) continue;
            const int i = (int)X0 - 3 + c - (int)a.pad;  // map column (inside the quadrant: the pixel was classified there)
            const uint32_t word = __ldg(reinterpret_cast<const uint32_t*>(
                a.frames + (uint64_t)f * a.frame_stride + (uint64_t)(a.map_y + (uint32_t)j) * a.pitch + (uint64_t)(a.map_x + (uint32_t)i) * 4u));
            a.ocr[(uint64_t)f * a.q_stride + (uint64_t)(j - bqh) * a.q_pitch + (uint32_t)(i - bqw + (int)a.q_pad)] = (uint8_t)(255u - luma8_bgra(word));
        }
    }
}

}  // namespace smh
