// filter_hash.hpp — hash functions of the gram prefilter, shared by the host
// (bitmap construction at finalize) and the device (filter kernel).
//
// A "gram" is an aligned W-byte word of the haystack stream (W = 8 or 4), read
// little-endian as (lo, hi) 32-bit halves (hi = 0 for W = 4).
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define ACB_HD __host__ __device__ __forceinline__
#else
#define ACB_HD inline
#endif

namespace acb200 {

constexpr uint32_t FILTER_L1_BYTES = 220u * 1024u;          // level-1 bitmap, lives in shared memory
constexpr uint32_t FILTER_L1_BITS = FILTER_L1_BYTES * 8u;

// level 1 is a blocked Bloom filter with two bits per gram inside ONE 32-bit word (one shared-memory
// load per haystack word): word = reduce(mix1) >> 5, first bit = reduce(mix1) & 31, the second one taken from
// low-order bits of the same hash, which the high-multiply range reduction does not look at.
// A gram is the aligned word plus the byte that follows it: an occurrence that "belongs" to the word ends
// at least one byte after it, so that byte is part of the occurrence too (one more byte of selectivity).
ACB_HD uint32_t filter_mix1(uint32_t lo, uint32_t hi, uint32_t next_byte)
{
    return lo * 0x9E3779B1u + hi * 0x85EBCA77u + next_byte * 0x7FEB352Du;
}

ACB_HD uint32_t filter_reduce(uint32_t t, uint32_t n_bits)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(t, n_bits);
#else
    return (uint32_t)(((uint64_t)t * n_bits) >> 32);
#endif
}

// "next byte unknown" (the last word of a 512-byte span, whose successor belongs to another warp): every gram
// is entered a second time with this value, so such a word is tested on its W bytes alone.  These entries
// live in their own eighth of the bitmap (only one word in 64 or 128 looks there), the others in the rest.
constexpr uint32_t FILTER_NEXT_UNKNOWN = 0x100u;
constexpr uint32_t FILTER_L1_WORDS = FILTER_L1_BITS / 32u;
constexpr uint32_t FILTER_L1_UNKNOWN_WORDS = FILTER_L1_WORDS / 8u;
constexpr uint32_t FILTER_L1_KNOWN_WORDS = FILTER_L1_WORDS - FILTER_L1_UNKNOWN_WORDS;

// 32-bit word of level 1 a gram hash selects (high multiply: looks at the high-order bits of the hash) ...
ACB_HD uint32_t filter_l1_word(uint32_t t, bool next_unknown)
{
    return next_unknown ? FILTER_L1_KNOWN_WORDS + filter_reduce(t, FILTER_L1_UNKNOWN_WORDS)
                        : filter_reduce(t, FILTER_L1_KNOWN_WORDS);
}
// ... and the two bits inside it (low-order bits of the hash, which the word index does not look at)
ACB_HD uint32_t filter_bit1(uint32_t t) { return t & 31u; }
ACB_HD uint32_t filter_bit2(uint32_t t) { return (t >> 5) & 31u; }

// level 2 (a 2^log2_bits-bit map in global memory) is indexed by a remix of the level-1 hash: three instructions
// where an independent hash of the word costs eight — the filter loop is issue-bound, and with W = 4 every word of the
// haystack pays for this index.  Two grams collide in BOTH levels only if their 32-bit level-1 hashes are equal.
ACB_HD uint32_t filter_l2_index(uint32_t t, uint32_t log2_bits)
{
    return ((t ^ (t >> 15)) * 0x2C1B3C6Du) >> (32u - log2_bits);
}

// third independent hash of a gram: the home slot of the exact gram table (gram_table.hpp)
ACB_HD uint32_t filter_mix3(uint32_t lo, uint32_t hi, uint32_t next_byte)
{
    uint32_t t = (lo ^ 0x5bd1e995u) * 0x165667B1u + (hi ^ 0x7feb352du) * 0xD3A2646Du + next_byte * 0x9E3779B1u;
    t ^= t >> 16;
    return t * 0x846CA68Bu;
}

} // namespace acb200
