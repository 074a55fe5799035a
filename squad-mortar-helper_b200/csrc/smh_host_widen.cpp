// smh_host_widen.cpp — host side of the grayscale ui_map transfer of smh_b200_process_host_batch: only the L
// byte of every (L, L, L, 255) pixel crosses the bus; this widens it into the caller's RGBA buffer
// (`image`'s Luma -> Rgba conversion, vision-cpu/src/lib.rs:152-158).  Streaming stores: the destination is
// written once and not read here, so it should not be pulled into the cache first.
#include <stddef.h>
#include <stdint.h>

#if defined(__x86_64__)
#include <immintrin.h>

__attribute__((target("avx2"))) static void widen_avx2(const uint8_t* src, uint32_t* dst, size_t n) {
    size_t i = 0;
    for (; i < n && ((uintptr_t)(dst + i) & 31u); i++) dst[i] = 0xFF000000u | ((uint32_t)src[i] * 0x00010101u);
    const __m256i mul = _mm256_set1_epi32(0x00010101), alpha = _mm256_set1_epi32((int)0xFF000000u);
    for (; i + 8 <= n; i += 8) {
        const __m256i l = _mm256_cvtepu8_epi32(_mm_loadl_epi64(reinterpret_cast<const __m128i*>(src + i)));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), _mm256_or_si256(_mm256_mullo_epi32(l, mul), alpha));
    }
    for (; i < n; i++) dst[i] = 0xFF000000u | ((uint32_t)src[i] * 0x00010101u);
    _mm_sfence();
}
#endif

extern "C" void smh_host_widen_l8_to_rgba(const uint8_t* src, uint32_t* dst, size_t n) {
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) { widen_avx2(src, dst, n); return; }
#endif
    for (size_t i = 0; i < n; i++) dst[i] = 0xFF000000u | ((uint32_t)src[i] * 0x00010101u);  // (L, L, L, 255)
}
