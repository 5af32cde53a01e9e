// libnrb200: host-side conversion of the uint16 wire format into the caller's int32 matrix (run.cuh, drain of a
// host-bound count matrix).  Plain C++ (compiled by the host compiler, not by nvcc's front end) so that it can carry
// AVX-512 / AVX2 variants chosen at run time; every variant writes with streaming stores: the destination is not read
// again by the converting thread and must not evict the staging data other threads still need.
#include <cstddef>
#include <cstdint>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define NRB_X86 1
#endif

namespace nrb {

#ifdef NRB_X86
__attribute__((target("avx512f,avx512bw"))) static void widen_avx512(const uint16_t *src, int32_t *dst, size_t n) {
    size_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 63)) {
        dst[i] = src[i];
        ++i;
    }
    for (; i + 16 <= n; i += 16) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
        _mm512_stream_si512(reinterpret_cast<__m512i *>(dst + i), _mm512_cvtepu16_epi32(v));
    }
    _mm_sfence();
    for (; i < n; ++i) dst[i] = src[i];
}

__attribute__((target("avx2"))) static void widen_avx2(const uint16_t *src, int32_t *dst, size_t n) {
    size_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31)) {
        dst[i] = src[i];
        ++i;
    }
    for (; i + 16 <= n; i += 16) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 8));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), _mm256_cvtepu16_epi32(a));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 8), _mm256_cvtepu16_epi32(b));
    }
    _mm_sfence();
    for (; i < n; ++i) dst[i] = src[i];
}

static void widen_sse2(const uint16_t *src, int32_t *dst, size_t n) {
    size_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 15)) {
        dst[i] = src[i];
        ++i;
    }
    const __m128i zero = _mm_setzero_si128();
    for (; i + 8 <= n; i += 8) {
        const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), _mm_unpacklo_epi16(v, zero));
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 4), _mm_unpackhi_epi16(v, zero));
    }
    _mm_sfence();
    for (; i < n; ++i) dst[i] = src[i];
}
#endif

using WidenFn = void (*)(const uint16_t *, int32_t *, size_t);

static void widen_scalar(const uint16_t *src, int32_t *dst, size_t n) {
    for (size_t i = 0; i < n; ++i) dst[i] = src[i];
}

static WidenFn pick() {
#ifdef NRB_X86
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw")) return widen_avx512;
    if (__builtin_cpu_supports("avx2")) return widen_avx2;
    return widen_sse2;
#else
    return widen_scalar;
#endif
}

// uint16 -> int32, n elements
void widen_u16_to_i32(const uint16_t *src, int32_t *dst, size_t n) {
    static const WidenFn fn = pick();
    fn(src, dst, n);
}

// Copy with streaming (non-temporal) stores: staging of host data that a DMA engine reads next.  Lines written with
// ordinary stores sit dirty in the writing cores' caches and every PCIe read of them has to be snooped out (measured:
// a 4 MB column staged by memcpy uploads at 30 GB/s instead of the link's 50+); streamed lines are in DRAM.
#ifdef NRB_X86
__attribute__((target("avx2"))) static void copy_stream_avx2(const char *src, char *dst, size_t n) {
    size_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31)) {
        dst[i] = src[i];
        ++i;
    }
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 64));
        const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 96), d);
    }
    _mm_sfence();
    for (; i < n; ++i) dst[i] = src[i];
}
static void copy_stream_sse2(const char *src, char *dst, size_t n) {
    size_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 15)) {
        dst[i] = src[i];
        ++i;
    }
    for (; i + 32 <= n; i += 32) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 16));
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), a);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 16), b);
    }
    _mm_sfence();
    for (; i < n; ++i) dst[i] = src[i];
}
#endif

void copy_streaming(const void *src, void *dst, size_t bytes) {
#ifdef NRB_X86
    using Fn = void (*)(const char *, char *, size_t);
    static const Fn fn = [] {
        __builtin_cpu_init();
        return __builtin_cpu_supports("avx2") ? (Fn)copy_stream_avx2 : (Fn)copy_stream_sse2;
    }();
    fn(static_cast<const char *>(src), static_cast<char *>(dst), bytes);
#else
    __builtin_memcpy(dst, src, bytes);
#endif
}

const char *widen_variant() {
#ifdef NRB_X86
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw")) return "avx512";
    if (__builtin_cpu_supports("avx2")) return "avx2";
    return "sse2";
#else
    return "scalar";
#endif
}

}  // namespace nrb
