// TMA probe: which box widths / start coordinates does cp.async.bulk.tensor.3d accept for a u32 tensor?
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tmap, int x, int y, int z, int bytes, uint32_t* out, int nwords, int mode) {
    extern __shared__ __align__(128) uint8_t sm[];
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm);
    const uint32_t bar = dst + 8192;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (mode == 0 ? threadIdx.x == 0 : (threadIdx.x & 31) == 0 && threadIdx.x == 32) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     :: "r"(dst), "l"(&tmap), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" :: "r"(bar) : "memory");
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) out[i] = ((uint32_t*)sm)[i];
}
int main() {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    const int W = 1280, H = 64, N = 2;
    std::vector<uint32_t> h((size_t)W * H * N);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint32_t)i;
    uint32_t *d, *o; cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 8192);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    const int boxes[][2] = {{136, 2}, {128, 2}, {132, 2}, {64, 2}, {136, 1}, {160, 2}, {192, 2}, {256, 2}};
    for (auto& b : boxes) {
        CUtensorMap tm;
        cuuint64_t dims[3] = {W, H, N}; cuuint64_t str[2] = {W * 4, (cuuint64_t)W * H * 4}; cuuint32_t box[3] = {(cuuint32_t)b[0], (cuuint32_t)b[1], 1}; cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("box %dx%d encode rc=%d\n", b[0], b[1], (int)r);
        if (r) continue;
        for (int x : {0, 4, 1, 3, 646, -4, 1200}) for (int mode : {0, 1}) {
            const int nw = b[0] * b[1];
            k<<<1, 64, 8192 + 64>>>(tm, x, 5, 1, nw * 4, o, nw, mode);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<uint32_t> got(nw);
            if (e == cudaSuccess) cudaMemcpy(got.data(), o, nw * 4, cudaMemcpyDeviceToHost);
            bool ok = e == cudaSuccess;
            if (ok) for (int r2 = 0; r2 < b[1]; r2++) for (int c = 0; c < b[0]; c++) {
                const int gx = x + c; const uint32_t want = (gx < 0 || gx >= W) ? 0u : (uint32_t)((size_t)1 * W * H + (size_t)(5 + r2) * W + gx);
                if (got[r2 * b[0] + c] != want) ok = false;
            }
            printf("  x=%5d mode=%d: %s %s\n", x, mode, cudaGetErrorString(e), ok ? "data ok" : "DATA MISMATCH");
            if (e != cudaSuccess) { printf("sticky error, stop\n"); return 1; }
        }
    }
    return 0;
}
