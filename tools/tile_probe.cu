// tile_probe.cu — what does HBM give a 2-D tiled read-modify-write of the pixel pass's shape, as a function of the strip
// width?  n frames of 2560x1440 BGRA; ROI 1536x1152 at (1024, 100) is read, a 4-byte and a 1-byte plane of the ROI's
// size are written.  One warp per (frame, band of 64 rows, strip of SW pixels); 2 KB of loads in flight per warp.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tile_probe tile_probe.cu ; run: ./tile_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int W = 2560, H = 1440, RX = 1216, RY = 100, RW = 1344, RH = 1152, BAND = 64;
template <int SW, int MODE>  // MODE 0: read + both writes, 1: read + 4-byte plane only, 2: linear order (frame-major rows)
__global__ void k(const uint32_t* __restrict__ in, uint32_t* __restrict__ ui, uint8_t* __restrict__ lsd, int n) {
    const int lane = threadIdx.x & 31;
    const long item = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    constexpr int NS = (RW + SW - 1) / SW, NB = RH / BAND, V = SW / 128;  // vectors per lane per row
    const int f = item / (NS * NB);
    if (f >= n) return;
    const int sub = item - (long)f * NS * NB, b = sub / NS, s = sub - b * NS;
    const uint32_t* src = in + ((long)f * H + RY + b * BAND) * W + RX + s * SW;
    uint32_t* du = ui + ((long)f * RH + b * BAND) * RW + s * SW;
    uint8_t* dl = lsd + ((long)f * RH + b * BAND) * RW + s * SW;
#pragma unroll 1
    constexpr int RU = 4 / V;
    for (int r = 0; r < BAND; r += RU) {
        uint4 v[RU][V];
#pragma unroll
        for (int k2 = 0; k2 < RU; k2++)
#pragma unroll
            for (int q = 0; q < V; q++) if (s * SW + q * 128 + lane * 4 < RW) v[k2][q] = __ldg(reinterpret_cast<const uint4*>(src + (long)(r + k2) * W + q * 128 + lane * 4));
#pragma unroll
        for (int k2 = 0; k2 < RU; k2++)
#pragma unroll
            for (int q = 0; q < V; q++) {
                if (s * SW + q * 128 + lane * 4 >= RW) continue;
                uint4 x = v[k2][q];
                x.x = x.x * 0x010101u + 0xB4000000u; x.y ^= 0xFF000000u; x.z += 3u; x.w |= 0xFF000000u;
                *reinterpret_cast<uint4*>(du + (long)(r + k2) * RW + q * 128 + lane * 4) = x;
                if (MODE == 0) *reinterpret_cast<uint32_t*>(dl + (long)(r + k2) * RW + q * 128 + lane * 4) = (x.x ^ x.y ^ x.z ^ x.w) & 0x01010101u;
            }
    }
}
template <int SW, int MODE> float run(const uint32_t* in, uint32_t* ui, uint8_t* lsd, int n, int smem = 0, int threads = 128) {
    constexpr int NS = (RW + SW - 1) / SW, NB = RH / BAND;
    const long items = (long)n * NS * NB;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int it = 0; it < 5; it++) {
        cudaEventRecord(a);
        k<SW, MODE><<<(unsigned)((items + threads / 32 - 1) / (threads / 32)), threads, smem>>>(in, ui, lsd, n);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (it && ms < best) best = ms;
    }
    return best;
}
__global__ void copyk(const uint4* __restrict__ a, uint4* __restrict__ b, long nv) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long)gridDim.x * blockDim.x) b[i] = a[i];
}
int main() {
    const int n = 1024;
    uint32_t *in, *ui; uint8_t* lsd;
    cudaMalloc(&in, (size_t)n * W * H * 4); cudaMalloc(&ui, (size_t)n * RW * RH * 4); cudaMalloc(&lsd, (size_t)n * RW * RH);
    cudaMemset(in, 1, (size_t)n * W * H * 4);
    const double roi = (double)n * RW * RH;
    {   // linear copy of the same number of bytes for reference
        const long nv = (long)(roi * 4 / 16);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); float best = 1e9;
        for (int it = 0; it < 5; it++) { cudaEventRecord(a); copyk<<<148 * 16, 256>>>((const uint4*)in, (uint4*)ui, nv); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); if (it && ms < best) best = ms; }
        printf("linear copy          %.3f ms  %.0f GB/s\n", best, roi * 8 / best / 1e6);
    }
    float t;
    t = run<128, 0>(in, ui, lsd, n); printf("strip 128 r+w4+w1    %.3f ms  %.0f GB/s\n", t, roi * 9 / t / 1e6);
    t = run<256, 0>(in, ui, lsd, n); printf("strip 256 r+w4+w1    %.3f ms  %.0f GB/s\n", t, roi * 9 / t / 1e6);
    t = run<512, 0>(in, ui, lsd, n); printf("strip 512 r+w4+w1    %.3f ms  %.0f GB/s\n", t, roi * 9 / t / 1e6);
    t = run<128, 1>(in, ui, lsd, n); printf("strip 128 r+w4       %.3f ms  %.0f GB/s\n", t, roi * 8 / t / 1e6);
    t = run<256, 1>(in, ui, lsd, n); printf("strip 256 r+w4       %.3f ms  %.0f GB/s\n", t, roi * 8 / t / 1e6);
    t = run<512, 1>(in, ui, lsd, n); printf("strip 512 r+w4       %.3f ms  %.0f GB/s\n", t, roi * 8 / t / 1e6);
    for (int co : {0, 25, 50, 75, 100}) {  // shared-memory carve-out preference (percent) with no shared memory in use: only the L1 shrinks
        cudaFuncSetAttribute(k<128, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, co);
        t = run<128, 0>(in, ui, lsd, n, 0); printf("strip 128 r+w4+w1, carve-out %3d %%          %.3f ms  %.0f GB/s\n", co, t, roi * 9 / t / 1e6);
    }
    cudaFuncSetAttribute(k<128, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, -1);
    for (int kb : {1, 4, 8, 12, 20, 27, 36}) {  // dynamic shared memory caps the CTAs per SM: 227 KB / (kb + 1) -> 10, 8, 6, 4 CTAs of 4 warps
        t = run<128, 0>(in, ui, lsd, n, kb * 1024); printf("strip 128 r+w4+w1, %2d KB smem per CTA  %.3f ms  %.0f GB/s\n", kb, t, roi * 9 / t / 1e6);
    }
    // resident warps per SM capped by unused dynamic shared memory while the carve-out stays below its maximum (70 % = 164 KB)
    cudaFuncSetAttribute(k<128, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, 70);
    for (int kb : {9, 12, 15, 19, 25, 39}) {
        int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k<128, 0>, 128, kb * 1024);
        t = run<128, 0>(in, ui, lsd, n, kb * 1024); printf("strip 128 r+w4+w1, carve-out 70 %%, %2d KB per CTA (%2d CTAs = %2d warps per SM)  %.3f ms  %.0f GB/s\n", kb, nb, nb * 4, t, roi * 9 / t / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
