// tile_probe2.cu — tile_probe.cu's 128-column tile kernel with the pixel pass's features switched on one at a time, to
// find which of them costs HBM throughput.  FEAT bits: 1 = ld.global.nc.L1::no_allocate loads, 2 = two extra rows read per
// band (the rows above / below), 4 = a tile-major 1-bit plane written as 4-byte words by lanes 0, 8, 16, 24, 8 = output
// rows padded to 128-byte multiples, 16 = quadrant planes (two more 1-byte planes over a quarter of the ROI),
// 32 = 10.5 strips (ROI 1344 wide: last strip half empty) instead of 1280.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>
constexpr int W = 2560, H = 1440, RX = 1216, RY = 100, RH = 1152, BAND = 64;
__device__ __forceinline__ uint4 ld_na(const void* p) {
    uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r;
}
template <int FEAT>
__global__ void k(const uint32_t* __restrict__ in, uint32_t* __restrict__ ui, uint8_t* __restrict__ lsd, uint32_t* __restrict__ bits,
                  uint8_t* __restrict__ q1, uint8_t* __restrict__ q2, int n) {
    constexpr int RW = (FEAT & 32) ? 1344 : 1280, NS = (RW + 127) / 128, NB = RH / BAND;
    constexpr int UP = (FEAT & 8) ? (RW + 31) / 32 * 32 : RW;  // output pitch in pixels
    constexpr int LAGR = (FEAT & 2) ? 1 : 0;
    const int lane = threadIdx.x & 31;
    const long item = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int f = item / (NS * NB);
    if (f >= n) return;
    const int sub = item - (long)f * NS * NB, b = sub / NS, s = sub - b * NS;
    const bool act = s * 128 + lane * 4 < RW;
    const uint32_t* src = in + ((long)f * H + RY + b * BAND - LAGR) * W + RX + s * 128 + lane * 4;
    uint32_t* du = ui + ((long)f * RH + b * BAND) * UP + s * 128 + lane * 4;
    uint8_t* dl = lsd + ((long)f * RH + b * BAND) * UP + s * 128 + lane * 4;
    uint32_t* db = bits + (long)f * (RH / 32) * (NS * 4) * 32 + ((long)(b * 2) * (NS * 4) + s * 4 + (lane >> 3)) * 32;
    const bool inq = (FEAT & 16) && b >= NB / 2 && s >= NS / 2;
    uint8_t* dq1 = q1 + ((long)f * RH + b * BAND) * UP + s * 128 + lane * 4;
    uint8_t* dq2 = q2 + ((long)f * RH + b * BAND) * UP + s * 128 + lane * 4;
    uint32_t acc = 0;
#pragma unroll 1
    for (int r = -LAGR; r < BAND + LAGR; r += 4) {
        uint4 v[4];
#pragma unroll
        for (int k2 = 0; k2 < 4; k2++) {
            const void* p = src + (long)(r + LAGR + k2) * W;
            v[k2] = (act && r + k2 < BAND + LAGR) ? ((FEAT & 1) ? ld_na(p) : __ldg(reinterpret_cast<const uint4*>(p))) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int k2 = 0; k2 < 4; k2++) {
            const int j = r + k2;
            uint4 x = v[k2];
            x.x = x.x * 0x010101u + 0xB4000000u; x.y ^= 0xFF000000u; x.z += 3u; x.w |= 0xFF000000u;
            acc ^= x.x;
            if (j < 0 || j >= BAND) continue;
            if (act) {
                *reinterpret_cast<uint4*>(du + (long)j * UP) = x;
                *reinterpret_cast<uint32_t*>(dl + (long)j * UP) = (x.x ^ x.y ^ x.z ^ x.w) & 0x01010101u;
                if (inq) { *reinterpret_cast<uint32_t*>(dq1 + (long)j * UP) = x.y; *reinterpret_cast<uint32_t*>(dq2 + (long)j * UP) = x.z; }
            }
            if ((FEAT & 4) && (lane & 7) == 0) db[(j >> 5) * (NS * 4) * 32 + (j & 31)] = acc;
        }
    }
}
template <int FEAT> void run(const uint32_t* in, uint32_t* ui, uint8_t* lsd, uint32_t* bits, uint8_t* q1, uint8_t* q2, int n, const char* what) {
    constexpr int RW = (FEAT & 32) ? 1344 : 1280, NS = (RW + 127) / 128, NB = RH / BAND;
    const long items = (long)n * NS * NB;
    cudaFuncSetAttribute(k<FEAT>, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int it = 0; it < 5; it++) {
        cudaEventRecord(a);
        k<FEAT><<<(unsigned)((items + 3) / 4), 128>>>(in, ui, lsd, bits, q1, q2, n);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (it && ms < best) best = ms;
    }
    double bytes = (double)n * RW * RH * 9.0;
    if (FEAT & 2) bytes += (double)n * RW * RH * 4.0 * 2 / 64;
    if (FEAT & 4) bytes += (double)n * RW * RH / 8;
    if (FEAT & 16) bytes += (double)n * RW * RH * 2 / 4;
    printf("%-58s %.3f ms  %.0f GB/s (bytes counted %.2f GB)\n", what, best, bytes / best / 1e6, bytes / 1e9);
}

// ---- the same tiles through a per-warp ring of TMA boxes (128 pixels x ROWS rows), as k_pixels2 moves them -------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int x, int y, int z, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one() { uint32_t p; asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(p)); return p != 0u; }
__device__ __forceinline__ uint4 lds_v4(uint32_t a) { uint4 r; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a)); return r; }
template <int FEAT, int STAGES, int ROWS>
__global__ void __launch_bounds__(128, 8) ktma(const __grid_constant__ CUtensorMap tmap, uint32_t* __restrict__ ui, uint8_t* __restrict__ lsd, uint32_t* __restrict__ bits,
                     uint8_t* __restrict__ q1, uint8_t* __restrict__ q2, int n) {
    extern __shared__ __align__(128) uint8_t sm[];
    constexpr int RW = (FEAT & 32) ? 1344 : 1280, NS = (RW + 127) / 128, NB = RH / BAND;
    constexpr int UP = (FEAT & 8) ? (RW + 31) / 32 * 32 : RW;
    constexpr int LAGR = (FEAT & 2) ? 1 : 0;
    constexpr int WB = STAGES * ROWS * 512 + 128;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long item = (long)blockIdx.x * 4 + warp;
    const int f = item / (NS * NB);
    if (f >= n) return;
    const int sub = item - (long)f * NS * NB, b = sub / NS, s = sub - b * NS;
    const bool act = s * 128 + lane * 4 < RW;
    uint32_t* du = ui + ((long)f * RH + b * BAND) * UP + s * 128 + lane * 4;
    uint8_t* dl = lsd + ((long)f * RH + b * BAND) * UP + s * 128 + lane * 4;
    uint32_t* db = bits + (long)f * (RH / 32) * (NS * 4) * 32 + ((long)(b * 2) * (NS * 4) + s * 4 + (lane >> 3)) * 32;
    const bool inq = (FEAT & 16) && b >= NB / 2 && s >= NS / 2;
    uint8_t* dq1 = q1 + ((long)f * RH + b * BAND) * UP + s * 128 + lane * 4;
    uint8_t* dq2 = q2 + ((long)f * RH + b * BAND) * UP + s * 128 + lane * 4;
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(sm) + warp * WB, bars = ring + STAGES * ROWS * 512;
    if (lane == 0) { for (int i = 0; i < STAGES; i++) mbar_init(bars + 8 * i, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    const int nrows = BAND + 2 * LAGR, nch = (nrows + ROWS - 1) / ROWS;
    const int tx = RX + s * 128, ty = RY + b * BAND - LAGR;
    if (elect_one()) for (int c = 0; c < STAGES && c < nch; c++) { mbar_expect_tx(bars + 8 * c, ROWS * 512); tma_load_3d(ring + c * ROWS * 512, &tmap, tx, ty + c * ROWS, f, bars + 8 * c); }
    uint32_t stage = 0, phase = 0, acc = 0;
#pragma unroll 1
    for (int c = 0; c < nch; c++) {
        mbar_wait(bars + 8 * stage, phase);
#pragma unroll
        for (int k2 = 0; k2 < ROWS; k2++) {
            const int j = c * ROWS + k2 - LAGR;
            uint4 x = lds_v4(ring + stage * ROWS * 512 + k2 * 512 + lane * 16);
            x.x = x.x * 0x010101u + 0xB4000000u; x.y ^= 0xFF000000u; x.z += 3u; x.w |= 0xFF000000u;
            acc ^= x.x;
            if (j < 0 || j >= BAND) continue;
            if (act) {
                *reinterpret_cast<uint4*>(du + (long)j * UP) = x;
                *reinterpret_cast<uint32_t*>(dl + (long)j * UP) = (x.x ^ x.y ^ x.z ^ x.w) & 0x01010101u;
                if (inq) { *reinterpret_cast<uint32_t*>(dq1 + (long)j * UP) = x.y; *reinterpret_cast<uint32_t*>(dq2 + (long)j * UP) = x.z; }
            }
            if ((FEAT & 4) && (lane & 7) == 0) db[(j >> 5) * (NS * 4) * 32 + (j & 31)] = acc;
        }
        __syncwarp();
        if (c + STAGES < nch && elect_one()) { mbar_expect_tx(bars + 8 * stage, ROWS * 512); tma_load_3d(ring + stage * ROWS * 512, &tmap, tx, ty + (c + STAGES) * ROWS, f, bars + 8 * stage); }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int FEAT, int STAGES, int ROWS> void run_tma(const uint32_t* in, uint32_t* ui, uint8_t* lsd, uint32_t* bits, uint8_t* q1, uint8_t* q2, int n, const char* what) {
    constexpr int RW = (FEAT & 32) ? 1344 : 1280, NS = (RW + 127) / 128, NB = RH / BAND;
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap tmap;
    const cuuint64_t dims[3] = {W, H, (cuuint64_t)n}; const cuuint64_t strides[2] = {W * 4, (cuuint64_t)W * H * 4};
    const cuuint32_t box[3] = {128, ROWS, 1}, estr[3] = {1, 1, 1};
    CUresult r = ((EncodeTiledFn)p)(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void*)in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return; }
    const long items = (long)n * NS * NB;
    const int smem = 4 * (STAGES * ROWS * 512 + 128);
    cudaFuncSetAttribute(ktma<FEAT, STAGES, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int it = 0; it < 5; it++) {
        cudaEventRecord(a);
        ktma<FEAT, STAGES, ROWS><<<(unsigned)((items + 3) / 4), 128, smem>>>(tmap, ui, lsd, bits, q1, q2, n);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (it && ms < best) best = ms;
    }
    double bytes = (double)n * RW * RH * 9.0;
    if (FEAT & 2) bytes += (double)n * RW * RH * 4.0 * 2 / 64;
    if (FEAT & 4) bytes += (double)n * RW * RH / 8;
    if (FEAT & 16) bytes += (double)n * RW * RH * 2 / 4;
    printf("%-58s %.3f ms  %.0f GB/s (bytes counted %.2f GB)\n", what, best, bytes / best / 1e6, bytes / 1e9);
}
int main() {
    const int n = 1024;
    uint32_t *in, *ui, *bits; uint8_t *lsd, *q1, *q2;
    cudaMalloc(&in, (size_t)n * W * H * 4); cudaMalloc(&ui, (size_t)n * 1408 * RH * 4); cudaMalloc(&lsd, (size_t)n * 1408 * RH);
    cudaMalloc(&q1, (size_t)n * 1408 * RH); cudaMalloc(&q2, (size_t)n * 1408 * RH); cudaMalloc(&bits, (size_t)n * 1408 * RH / 8 + 4096);
    cudaMemset(in, 1, (size_t)n * W * H * 4);
    run<0>(in, ui, lsd, bits, q1, q2, n, "base: 10 strips, __ldg, r + w4 + w1");
    run<1>(in, ui, lsd, bits, q1, q2, n, "+ L1::no_allocate loads");
    run<2>(in, ui, lsd, bits, q1, q2, n, "+ 2 halo rows per band");
    run<4>(in, ui, lsd, bits, q1, q2, n, "+ tile-major bitmap words");
    run<8>(in, ui, lsd, bits, q1, q2, n, "+ 128-byte padded output rows");
    run<16>(in, ui, lsd, bits, q1, q2, n, "+ quadrant planes");
    run<32>(in, ui, lsd, bits, q1, q2, n, "+ 10.5 strips");
    run<63>(in, ui, lsd, bits, q1, q2, n, "all of them");
    run<62>(in, ui, lsd, bits, q1, q2, n, "all but no_allocate");
    run<59>(in, ui, lsd, bits, q1, q2, n, "all but bitmap");
    run_tma<0, 3, 2>(in, ui, lsd, bits, q1, q2, n, "TMA ring 3 x 2 rows: base");
    run_tma<0, 6, 2>(in, ui, lsd, bits, q1, q2, n, "TMA ring 6 x 2 rows: base");
    run_tma<0, 4, 4>(in, ui, lsd, bits, q1, q2, n, "TMA ring 4 x 4 rows: base");
    run_tma<58, 3, 2>(in, ui, lsd, bits, q1, q2, n, "TMA ring 3 x 2 rows: all but bitmap");
    run_tma<58, 6, 2>(in, ui, lsd, bits, q1, q2, n, "TMA ring 6 x 2 rows: all but bitmap");
    run_tma<62, 6, 2>(in, ui, lsd, bits, q1, q2, n, "TMA ring 6 x 2 rows: all");
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
