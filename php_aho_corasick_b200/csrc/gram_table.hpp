// gram_table.hpp — exact table of the prefilter grams and the direct verification of a flagged word,
// shared by the host (table construction at finalize, acb200_direct_probe) and the device
// (ac_walk_kernel; ac_settle_kernel on the opt-in fused path).
//
// The prefilter (filter_kernels.cuh) flags an aligned W-byte haystack word k when (word, byte after it) MAY be
// the gram pattern[L-W-r, L-r] + pattern[L-r] of some accepted pattern, r = 1..W.  Walking the automaton over
// the (Lmax-1)-byte warm-up and the W bytes after the word settles what really ends at the W offsets the word
// owns — Lmax-1+W dependent table lookups per flagged word.  Most flagged words need none of them:
//
//   * the table below holds EVERY gram of every accepted pattern under its exact key (word, next byte).  A
//     flagged word whose key is absent was a Bloom false positive: nothing ends in its window.
//   * a key that belongs to exactly one (pattern P, r) admits exactly one candidate: P ending at
//     p = W(k+1) + r.  Any other pattern ending in the window would have entered the same key (all patterns are
//     at least 2W long, so each owns a gram at every r).  One comparison of the haystack against P's bytes
//     decides it.
//   * the event the reference reports at p carries the automaton state after byte p-1, i.e. the deepest trie
//     node that is a suffix of the text.  If no trie node has P's node as its failure target, no deeper node
//     ends with P, so that state is P's own node — known at finalize.
//
// Keys shared by several (pattern, r) pairs and patterns whose node is a failure target are marked
// GRAM_WALK: those words (and words whose window is clipped by the ends of the stream) are walked as before.
// Events are identical to the walk's by construction; tests/test_direct_tables.py checks the construction
// against the CPU oracle, tests/test_gpu_filter.py the kernel against the full walk.
//
// Replaces, for those words, the loop of src/multifast/ahocorasick.c:199-234 (same events).
#pragma once

#include <cstdint>

#include "filter_hash.hpp"

namespace acb200 {

struct GramSlot {                 // 32 bytes = one sector: patterns of up to 16 bytes are settled by this one load
    uint32_t key_lo, key_hi;      // the word (key_hi = 0 for W = 4)
    uint32_t meta;                // GRAM_USED | GRAM_WALK | GRAM_INLINE | length << 13 | r << 9 | next byte
    uint32_t ref;                 // GRAM_INLINE: the pattern's state id; else index into the pattern store of the word
                                  // after the pattern's last byte (which holds the state id)
    uint32_t tail[4];             // the pattern's last 16 bytes (zero-padded in front)
};

constexpr uint32_t GRAM_USED = 0x80000000u;
constexpr uint32_t GRAM_WALK = 0x40000000u;     // not decidable by one comparison: walk the automaton
constexpr uint32_t GRAM_INLINE = 0x20000000u;   // the whole pattern (<= 16 bytes) is in `tail`, `ref` is its state

ACB_HD uint32_t gram_meta(uint32_t len, uint32_t r, uint32_t next_byte) { return GRAM_USED | (len << 13) | (r << 9) | next_byte; }
ACB_HD uint32_t gram_meta_len(uint32_t m) { return (m >> 13) & 0x7ffu; }
ACB_HD uint32_t gram_meta_r(uint32_t m) { return (m >> 9) & 0xfu; }
ACB_HD uint32_t gram_meta_next(uint32_t m) { return m & 0x1ffu; }

// slot a key starts probing at (linear probing, table of 2^log2 slots)
ACB_HD uint32_t gram_home(uint32_t lo, uint32_t hi, uint32_t next_byte, uint32_t log2_slots)
{
    return filter_mix3(lo, hi, next_byte) >> (32u - log2_slots);
}

enum GramVerdict : int { GRAM_NOTHING = 0, GRAM_EVENT = 1, GRAM_NEEDS_WALK = 2 };

// W bytes of the haystack as one integer (little endian): the comparison unit
template <int W> struct GramChunk;
template <> struct GramChunk<8> { typedef uint64_t type; };
template <> struct GramChunk<4> { typedef uint32_t type; };

// Decides the W end offsets rs+1 .. rs+W owned by the aligned word at rs-W.
//   load_text(i)  -> the W aligned haystack bytes at stream offset i (i is a multiple of W) as a chunk
//   load_slot(i)  -> GramSlot i
//   load_pat(i)   -> the W pattern-store bytes at 32-bit word index i as a chunk;  load_state(i) -> word i
// The caller guarantees that [rs - warm, rs + W) lies inside the stream (warm = Lmax-1 rounded up to W) and that
// [rs, rs + W) lies inside the haystack that starts at stream offset `hay_begin` (<= rs): a candidate that would
// start before hay_begin does not fit into the haystack and is no occurrence.
template <int W, typename LT, typename LS, typename LP, typename LST>
ACB_HD GramVerdict gram_verify(uint32_t rs, uint32_t warm, uint32_t hay_begin, uint32_t log2_slots, LT load_text, LS load_slot,
                               LP load_pat, LST load_state, uint32_t *end, uint32_t *state)
{
    // Written for SIMT execution: one probe loop, then straight-line code with a single exit — lanes of a warp that
    // reach different verdicts stay converged (early returns fragment the warp and multiply the instruction count).
    typedef typename GramChunk<W>::type chunk_t;
    constexpr uint32_t WORDS = W / 4;
    constexpr uint32_t TAIL_CHUNKS = 16 / W;
    const chunk_t word = load_text(rs - W);
    const chunk_t next = load_text(rs);
    const chunk_t prev = load_text(rs - 2u * W);      // always inside the warm-up: every pattern is at least 2W long
    const uint32_t lo = (uint32_t)word, hi = (W == 8) ? (uint32_t)((uint64_t)word >> 32) : 0u;
    const uint32_t nb = (uint32_t)next & 0xffu;
    const uint32_t mask = (1u << log2_slots) - 1u;
    uint32_t i = gram_home(lo, hi, nb, log2_slots);
    GramSlot s = load_slot(i);
    while ((s.meta & GRAM_USED) && !(s.key_lo == lo && s.key_hi == hi && gram_meta_next(s.meta) == nb)) {
        i = (i + 1u) & mask;
        s = load_slot(i);
    }
    const bool found = (s.meta & GRAM_USED) != 0;                        // else: no pattern owns this gram
    const bool walk = found && (s.meta & GRAM_WALK) != 0;
    const uint32_t r = found ? gram_meta_r(s.meta) : (uint32_t)W, len = gram_meta_len(s.meta);
    // a candidate that starts before its haystack is no occurrence
    bool ok = found && !walk && rs + r >= hay_begin + len;
    // chunk j (from the end) of the candidate = haystack bytes [rs + r - W(j+1), rs + r - Wj): the top W-r bytes of
    // the group at rs - W(j+1) and the low r bytes of the group after it.  Chunks 0 and 1 always exist and are full.
    const uint32_t sh = 8u * (r & (uint32_t)(W - 1)), back_sh = 8u * ((uint32_t)W - (r & (uint32_t)(W - 1))) & (8u * W - 1u);
    const chunk_t have0 = (r == (uint32_t)W) ? next : (chunk_t)((word >> sh) | (next << back_sh));
    const chunk_t have1 = (r == (uint32_t)W) ? word : (chunk_t)((prev >> sh) | (word << back_sh));
    chunk_t want0, want1;
    if (W == 8) {
        want0 = (chunk_t)(((uint64_t)s.tail[3] << 32) | s.tail[2]);
        want1 = (chunk_t)(((uint64_t)s.tail[1] << 32) | s.tail[0]);
    } else {
        want0 = (chunk_t)s.tail[3];
        want1 = (chunk_t)s.tail[2];
    }
    ok = ok && have0 == want0 && have1 == want1;
    const uint32_t n_chunks = (len + W - 1) / W;
    if (ok && n_chunks > 2u) {                                           // patterns longer than 2W: the rest, chunk by chunk
        chunk_t hi_grp = prev, lo_grp = prev;
        for (uint32_t j = 2; ok && j < n_chunks; ++j) {
            hi_grp = lo_grp;
            const uint32_t back = W * (j + 1u);
            lo_grp = (back <= warm) ? load_text(rs - back) : (chunk_t)0; // bytes before the warm-up are never part of the candidate
            chunk_t have = (r == (uint32_t)W) ? hi_grp : (chunk_t)((lo_grp >> sh) | (hi_grp << back_sh));
            if (j == n_chunks - 1u) {
                const uint32_t valid = len - W * j;                      // bytes of the pattern in its first chunk
                if (valid < (uint32_t)W) have &= ~(chunk_t)0 << (8u * (W - valid));
            }
            chunk_t want;
            if (j < TAIL_CHUNKS) want = (chunk_t)((j == 2u) ? s.tail[1] : s.tail[0]);       // W = 4 only
            else want = load_pat(s.ref - WORDS * (j + 1u));
            ok = have == want;
        }
    }
    *end = rs + r;
    *state = 0;
    if (ok) *state = (s.meta & GRAM_INLINE) ? s.ref : load_state(s.ref);
    return walk ? GRAM_NEEDS_WALK : ok ? GRAM_EVENT : GRAM_NOTHING;
}

} // namespace acb200
