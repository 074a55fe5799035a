// tile_probe2.cu — tile_probe.cu's 128-column tile kernel with the pixel pass's features switched on one at a time, to
// find which of them costs HBM throughput.  FEAT bits: 1 = ld.global.nc.L1::no_allocate loads, 2 = two extra rows read per
// band (the rows above / below), 4 = a tile-major 1-bit plane written as 4-byte words by lanes 0, 8, 16, 24, 8 = output
// rows padded to 128-byte multiples, 16 = quadrant planes (two more 1-byte planes over a quarter of the ROI),
// 32 = 10.5 strips (ROI 1344 wide: last strip half empty) instead of 1280.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
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
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
