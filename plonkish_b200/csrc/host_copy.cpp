// Host-side copy into the pinned staging ring (api.cu, Stager).  A separate translation unit so that the host
// compiler sees the x86 intrinsics directly (nvcc's front end does not have to parse immintrin.h).
//
// The ring's slots are written once by a copier thread and read once by the GPU's DMA engine; they are never read
// by the CPU.  A plain memcpy below glibc's non-temporal threshold allocates every destination line in the cache
// (a read-for-ownership per line) and the DMA read then has to be served from there or from DRAM after the
// write-back; streaming stores write the lines straight to memory.  Which of the two is faster depends on how many
// processes share the host's memory controllers, so the choice is the caller's (PLONKISH_CUDA_STAGE_NT).
#include <cstddef>
#include <cstdint>
#include <cstring>

#if defined(__x86_64__)
#include <immintrin.h>

__attribute__((target("avx2"))) static void copy_stream_avx2(void *dst, const void *src, size_t len) {
    char *d = (char *)dst;
    const char *s = (const char *)src;
    // head up to a 32-byte boundary of the destination
    size_t head = (32 - ((uintptr_t)d & 31)) & 31;
    if (head > len) head = len;
    if (head) { memcpy(d, s, head); d += head; s += head; len -= head; }
    size_t blocks = len / 128;
    for (size_t i = 0; i < blocks; ++i) {
        const __m256i a = _mm256_loadu_si256((const __m256i *)(s));
        const __m256i b = _mm256_loadu_si256((const __m256i *)(s + 32));
        const __m256i c = _mm256_loadu_si256((const __m256i *)(s + 64));
        const __m256i e = _mm256_loadu_si256((const __m256i *)(s + 96));
        _mm256_stream_si256((__m256i *)(d), a);
        _mm256_stream_si256((__m256i *)(d + 32), b);
        _mm256_stream_si256((__m256i *)(d + 64), c);
        _mm256_stream_si256((__m256i *)(d + 96), e);
        s += 128;
        d += 128;
    }
    len -= blocks * 128;
    if (len) memcpy(d, s, len);
    _mm_sfence();  // the DMA is enqueued after this returns: the streamed lines must be globally visible
}
#endif

extern "C" int plonkish_cuda_host_copy_has_stream(void) {
#if defined(__x86_64__)
    return __builtin_cpu_supports("avx2") ? 1 : 0;
#else
    return 0;
#endif
}

// stream != 0: non-temporal stores when the CPU has AVX2, else memcpy.
extern "C" void plonkish_cuda_host_copy(void *dst, const void *src, size_t len, int stream) {
#if defined(__x86_64__)
    if (stream && __builtin_cpu_supports("avx2")) {
        copy_stream_avx2(dst, src, len);
        return;
    }
#endif
    memcpy(dst, src, len);
}
