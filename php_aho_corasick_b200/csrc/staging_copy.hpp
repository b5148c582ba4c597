// staging_copy.hpp — how haystack bytes get into pinned staging on the host.  No CUDA: tests/test_helper_pool.py
// checks it against memcpy for every alignment and length class.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstring>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace acb200 {

// memcpy into pinned staging.  The CPU never reads these bytes again — the copy engine does — so on x86-64 the
// destination is written with non-temporal stores: no read-for-ownership of the staging lines (a third of a plain
// copy's memory traffic) and the haystacks' own lines are not pushed out of the cache by a buffer nobody reads.
// (131,072 strings of 8 KiB into one buffer, same box: 4.7-5.4 GB/s with memcpy, 5.7-6.4 GB/s this way, per core.)
inline void copy_to_staging(char *dst, const char *src, size_t n, bool stream_stores = true)
{
#if defined(__SSE2__)
    if (n >= 256 && stream_stores) {
        const size_t head = (16u - ((uintptr_t)dst & 15u)) & 15u;
        memcpy(dst, src, head);
        dst += head; src += head; n -= head;
        for (; n >= 64; n -= 64, src += 64, dst += 64) {
            const __m128i a = _mm_loadu_si128((const __m128i *)src), b = _mm_loadu_si128((const __m128i *)(src + 16));
            const __m128i c = _mm_loadu_si128((const __m128i *)(src + 32)), d = _mm_loadu_si128((const __m128i *)(src + 48));
            _mm_stream_si128((__m128i *)dst, a); _mm_stream_si128((__m128i *)(dst + 16), b);
            _mm_stream_si128((__m128i *)(dst + 32), c); _mm_stream_si128((__m128i *)(dst + 48), d);
        }
    }
#endif
    memcpy(dst, src, n);
}

// non-temporal stores are weakly ordered: fence before anyone is told that the bytes are there
inline void staging_fence()
{
#if defined(__SSE2__)
    _mm_sfence();
#endif
}

} // namespace acb200
