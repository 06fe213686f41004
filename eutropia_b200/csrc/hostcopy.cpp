// copy_nt: memcpy whose bulk uses non-temporal (streaming) stores.  The staging copies of
// hoststage.cu move hundreds of megabytes that the writing core never reads again (a pinned slot the
// DMA engine reads, or the caller's result matrix): streaming stores skip the read-for-ownership of the
// destination lines and leave the caches to the source.  Compiled by the host compiler (AVX2 path
// selected at run time).
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <immintrin.h>

namespace eut {

__attribute__((target("avx2"))) static void copy_nt_avx2(char *d, const char *s, size_t n)
{
    // head: up to the first 32-byte boundary of the destination
    const size_t head = (32 - ((uintptr_t)d & 31)) & 31;
    if (head) { std::memcpy(d, s, head); d += head; s += head; n -= head; }
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256((const __m256i *)(s + i));
        const __m256i b = _mm256_loadu_si256((const __m256i *)(s + i + 32));
        const __m256i c = _mm256_loadu_si256((const __m256i *)(s + i + 64));
        const __m256i e = _mm256_loadu_si256((const __m256i *)(s + i + 96));
        _mm256_stream_si256((__m256i *)(d + i), a);
        _mm256_stream_si256((__m256i *)(d + i + 32), b);
        _mm256_stream_si256((__m256i *)(d + i + 64), c);
        _mm256_stream_si256((__m256i *)(d + i + 96), e);
    }
    _mm_sfence();
    if (i < n) std::memcpy(d + i, s + i, n - i);
}

void copy_nt(void *dst, const void *src, size_t bytes)
{
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (bytes < 4096 || !avx2) { std::memcpy(dst, src, bytes); return; }
    copy_nt_avx2((char *)dst, (const char *)src, bytes);
}

}  // namespace eut
