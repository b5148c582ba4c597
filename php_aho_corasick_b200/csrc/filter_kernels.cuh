// filter_kernels.cuh — the HBM-streaming front end of ahocorasick_match() for
// dictionaries whose patterns are all at least 8 bytes long.
//
// Walking the automaton costs one dependent table lookup per haystack byte,
// which caps the scan far below what HBM3e delivers.  When no pattern is
// shorter than 2W bytes (W = 8 or 4) almost all of that work can be skipped
// without changing a single result:
//
//   ac_filter_kernel  streams the haystack once with coalesced 16-byte loads,
//                     hashes every aligned W-byte word into a bitmap held in
//                     shared memory (one bit test per word, built at finalize
//                     from the words that can precede the end of a pattern —
//                     see FlatAutomaton) and writes one bit per word: "the W
//                     end offsets after this word need a look".  No false
//                     negatives by construction.
//   ac_collect_kernel turns the bit planes into work items, one warp per 16 KiB
//                     tile (ballot / popc / prefix sums): one item per flagged
//                     word, or — when so many words of a tile are flagged that
//                     walking all of it is cheaper — one item per 512-byte
//                     span.  A tile's items are appended to one global list at
//                     an offset taken from a counter; {offset, count} per tile.
//   ac_walk_kernel    one thread per item, no shared memory, as many warps per
//                     SM as registers allow (the walk is a chain of dependent
//                     lookups: only parallelism hides it).  For a flagged word
//                     the thread walks the automaton from the root over the
//                     (Lmax-1)-byte warm-up plus the W bytes after the word —
//                     exactly the halo argument of ac_scan_kernel, so states
//                     and events are those of an uninterrupted walk; the table
//                     rows of the shallow states this touches live in L1.  A
//                     span item is walked like an ac_scan_kernel slice.  Result
//                     per item: event count + first event.
//   ac_tile_count_kernel / ac_emit_kernel
//                     sum the counts per tile, turn them into offsets (every
//                     CTA adds up the lengths of all earlier tiles itself — no
//                     CTA waits for another one) and write the events straight
//                     to their final, ascending place, as the callback
//                     contract requires.  The worst case costs about what the
//                     plain scan costs.
//
// Replaces the same reference loop as ac_scan_kernel
// (src/multifast/ahocorasick.c:199-234); events are bit-identical.
#pragma once

#include "scan_kernels.cuh"
#include "filter_hash.hpp"
#include "gram_table.hpp"

namespace acb200 {

constexpr uint32_t SPAN_BYTES = 512;       // one warp-wide 16-byte load; one verify lane
constexpr int FILTER_UNROLL = 3;           // 16-byte loads per thread and round; the next round's loads are issued before a round is
                                           // looked at, so 2 x 3 are in flight (measured, 1 GiB of config 2 / 256 MiB of config 3:
                                           // 1 -> 0.246 ms, 2 -> 0.181 / 0.130, 3 -> 0.181 / 0.127, 4 -> 0.185, 6 -> 0.197 / 0.148;
                                           // without the overlap 4 -> 0.187, 6 -> 0.189, 8 -> 0.196)
constexpr uint32_t VER_DENSE_MAX = 64;     // flagged words per 16 KiB tile beyond which the whole tile is walked
constexpr uint32_t ITEM_SPAN = 0x80000000u;// work item: walk 512-byte span (item & ~ITEM_SPAN) completely
constexpr uint32_t ITEM_NONE = 0xffffffffu;
constexpr int COLLECT_THREADS = 1024;      // 32 tiles per CTA iteration share one atomic (same-address atomics serialise)
constexpr int COUNT_THREADS = 256;
constexpr int WALK_THREADS = 256;
#ifndef ACB_WALK_ILP
#define ACB_WALK_ILP 1
#endif
constexpr int WALK_ILP = ACB_WALK_ILP;      // items a walk thread verifies in lockstep
constexpr int EMIT_THREADS = 256;          // one thread per tile in the offset scan, one warp per 32 tiles when emitting

struct FilterArgs {
    const uint8_t *text;          // 16-byte aligned
    uint32_t total;               // bytes in the stream
    const uint32_t *l1;           // level-1 bitmap (FILTER_L1_BITS bits)
    uint32_t l1_bits;
    const uint32_t *l2;           // level-2 bitmap or nullptr
    uint32_t l2_shift;            // 32 - log2(bits of level 2)
    uint32_t *mask;               // n_spans x (16/W) words: plane j bit c <=> word (c*(16/W) + j) of the span
    uint32_t n_spans;             // ceil(total / 512)
    uint32_t span_begin, span_end;// this launch filters spans [span_begin, span_end)
    uint32_t *zero;               // scratch words of the kernels that follow (their counters and block sums), cleared by CTA 0
    uint32_t n_zero;              // ... instead of by a memset in front of every call; 0: nothing to clear
};

// ------------------------------------------------------------- filter -----

// Stages the level-1 bitmap in shared memory (all threads of the CTA) and returns its shared-window address.
__device__ __forceinline__ uint32_t stage_bitmap(const FilterArgs &a, uint32_t *s_bm, uint32_t tid)
{
    const uint4 *src = reinterpret_cast<const uint4 *>(a.l1);
    uint4 *dst = reinterpret_cast<uint4 *>(s_bm);
    const uint32_t n4 = a.l1_bits >> 7;
    for (uint32_t i = tid; i < n4; i += SCAN_THREADS) dst[i] = __ldg(src + i);
    __syncthreads();
    return (uint32_t)__cvta_generic_to_shared(s_bm);
}

// One flag per aligned word: both bits of the gram hash (the word + the byte after it) are set in level 1
// (and, with L2, the bit of the independent hash in the level-2 bitmap in HBM/L2).
//
// The streaming loop issues ~60 instructions per 512-byte span and is as much bound by that as by HBM (ncu: issue
// slots 55 % busy at 0.80 of the copy bandwidth; a variant with 40 % more instructions took 28 % longer), so the
// test is written instruction by instruction: `region` / `n_words` select the part of the bitmap (per-lane
// constants for the span's last word, whose successor only lanes 0..30 know), the two bit positions are taken
// with wrap-around funnel shifts (no masking of the shift amounts).
template <bool L2>
__device__ __forceinline__ bool test_word_at(const FilterArgs &a, uint32_t region, uint32_t n_words, uint32_t lo, uint32_t hi, uint32_t nb)
{
    const uint32_t t = filter_mix1(lo, hi, nb);
    uint32_t word;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(word) : "r"(region + __umulhi(t, n_words) * 4u));
    // bit (t & 31) and bit ((t >> 5) & 31) — filter_bit1 / filter_bit2
    bool p = (__funnelshift_r(word, word, t) & __funnelshift_r(word, word, t >> 5) & 1u) != 0;
    if (L2) {
        uint32_t word3 = 0;
        const uint32_t i3 = filter_l2_index(t, 32u - a.l2_shift);
        if (p) word3 = __ldg(a.l2 + (i3 >> 5));
        p = (word3 >> (i3 & 31u)) & 1u;
    }
    return p;
}

template <bool L2>
__device__ __forceinline__ bool test_word(const FilterArgs &a, uint32_t s_base, uint32_t lo, uint32_t hi, uint32_t nb, bool maybe_unknown)
{
    const bool unknown = maybe_unknown && nb == FILTER_NEXT_UNKNOWN;
    return test_word_at<L2>(a, s_base + (unknown ? FILTER_L1_KNOWN_WORDS * 4u : 0u), unknown ? FILTER_L1_UNKNOWN_WORDS : FILTER_L1_KNOWN_WORDS,
                            lo, hi, nb);
}

template <int W, bool L2>
__global__ void __launch_bounds__(SCAN_THREADS, 1) ac_filter_kernel(const FilterArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *s_bm = reinterpret_cast<uint32_t *>(smem_raw);
    constexpr int NB = 16 / W;
    constexpr int U = FILTER_UNROLL;

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    if (blockIdx.x == 0)
        for (uint32_t i = tid; i < a.n_zero; i += SCAN_THREADS) a.zero[i] = 0u;
    const uint32_t s_base = stage_bitmap(a, s_bm, tid);
    // the span's last word of lane 31 is tested on its W bytes alone, in the bitmap's "next byte unknown" part
    const uint32_t last_region = s_base + (lane == 31u ? FILTER_L1_KNOWN_WORDS * 4u : 0u);
    const uint32_t last_words = lane == 31u ? FILTER_L1_UNKNOWN_WORDS : FILTER_L1_KNOWN_WORDS;
    auto test = [&](uint32_t lo, uint32_t hi, uint32_t nb, bool maybe_unknown) -> bool {
        return test_word<L2>(a, s_base, lo, hi, nb, maybe_unknown);
    };
    // planes of one complete span whose 16-byte chunks are in v (one per lane)
    auto span_planes = [&](const uint4 &v) -> uint32_t {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        // byte after the lane's last word: the next lane's first byte; lane 31 does not know it
        uint32_t after = __shfl_down_sync(0xffffffffu, w[0], 1) & 0xffu;
        if (lane == 31u) after = FILTER_NEXT_UNKNOWN;
        uint32_t mine = 0;
        bool p[NB];
        uint32_t i3[NB], word3[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const uint32_t lo = (W == 8) ? w[2 * j] : w[j], hi = (W == 8) ? w[2 * j + 1] : 0u;
            const uint32_t nb = (j == NB - 1) ? after : (((W == 8) ? w[2] : w[j + 1]) & 0xffu);
            p[j] = test_word_at<false>(a, (j == NB - 1) ? last_region : s_base, (j == NB - 1) ? last_words : FILTER_L1_KNOWN_WORDS, lo, hi, nb);
            if (L2) {
                // level 2: the index is a three-instruction remix of the level-1 hash, and the probes of all the lane's
                // words are issued before any is looked at (one trip to L2 per span, not one per word)
                i3[j] = filter_l2_index(filter_mix1(lo, hi, nb), 32u - a.l2_shift);
                word3[j] = 0;
                if (p[j]) word3[j] = __ldg(a.l2 + (i3[j] >> 5));
            }
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            if (L2) p[j] = (word3[j] >> (i3[j] & 31u)) & 1u;
            const uint32_t plane = __ballot_sync(0xffffffffu, p[j]);
            if (lane == (uint32_t)j) mine = plane;
        }
        return mine;
    };

    const uint32_t n_full_all = a.total / SPAN_BYTES;         // spans that lie completely inside the stream
    const uint32_t n_full = min(n_full_all, a.span_end);
    const uint32_t n_warps = gridDim.x * (SCAN_THREADS / 32);
    const uint32_t warp = blockIdx.x * (SCAN_THREADS / 32) + (tid >> 5);

    // full rounds: U spans per warp, no bounds checks inside
    uint32_t g0 = a.span_begin + warp;
    const uint8_t *src = a.text + ((size_t)g0 * 32u + lane) * 16u;
    const size_t span_stride = (size_t)n_warps * SPAN_BYTES;
    // The loads of round r + 1 are issued before round r is looked at: ncu's source view had half of this kernel's
    // stall samples on the first use of a round's first load.
    auto round_is_full = [&](uint32_t g) { return (uint64_t)g + (uint64_t)(U - 1) * n_warps < n_full; };
    uint4 v[U];
    if (round_is_full(g0)) {
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ld_text16(src + span_stride * u);
    }
    for (; round_is_full(g0); g0 += n_warps * U, src += span_stride * U) {
        uint4 nv[U];
        if (round_is_full(g0 + n_warps * U)) {
#pragma unroll
            for (int u = 0; u < U; ++u) nv[u] = ld_text16(src + span_stride * (U + u));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t mine = span_planes(v[u]);
            if (lane < (uint32_t)NB) a.mask[(size_t)(g0 + u * n_warps) * NB + lane] = mine;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = nv[u];
    }
    // the last, partial round
    for (; g0 < n_full; g0 += n_warps) {
        const uint4 v = ld_text16(a.text + ((size_t)g0 * 32u + lane) * 16u);
        const uint32_t mine = span_planes(v);
        if (lane < (uint32_t)NB) a.mask[(size_t)g0 * NB + lane] = mine;
    }

    // The last, partial span: complete 16-byte chunks are tested, the partial chunk at the very end is not
    // read at all — its words are simply handed on to verification.
    if (n_full_all < a.n_spans && n_full_all >= a.span_begin && n_full_all < a.span_end && warp == (n_full_all % n_warps)) {
        const uint32_t n_full = n_full_all;
        const uint32_t n16 = a.total >> 4;
        const uint32_t c = n_full * 32u + lane;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (c < n16) v = ld_text16(a.text + (size_t)c * 16u);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        const bool tail = (a.total & 15u) && c == n16;
        uint32_t after = __shfl_down_sync(0xffffffffu, w[0], 1) & 0xffu;
        if (lane == 31u || c + 1u >= n16) after = FILTER_NEXT_UNKNOWN;      // the next chunk is not in this warp's registers
        uint32_t mine = 0;
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const uint32_t nb = (j == NB - 1) ? after : (((W == 8) ? w[2] : w[j + 1]) & 0xffu);
            bool p = (W == 8) ? test(w[2 * j], w[2 * j + 1], nb, j == NB - 1) : test(w[j], 0u, nb, j == NB - 1);
            p = (c < n16) ? p : tail;
            const uint32_t plane = __ballot_sync(0xffffffffu, p);
            if (lane == (uint32_t)j) mine = plane;
        }
        if (lane < (uint32_t)NB) a.mask[(size_t)n_full * NB + lane] = mine;
    }
}

// ------------------------------------------------------------ collect -----

struct VerifyArgs {
    ScanArgs s;                   // stream, automaton, event buffer (win_rows = 0: no shared-memory window)
    const uint32_t *mask;         // bit planes written by ac_filter_kernel
    uint32_t n_spans;             // 512-byte spans in the stream
    uint32_t n_tiles;             // 16 KiB tiles = ceil(n_spans / 32)
    uint32_t tile_begin, tile_end;// collect / walk launches work on the tiles [tile_begin, tile_end) of one part
    uint32_t item_base;           // that part's region of items / recs starts here
    uint32_t counter_slot;        // counters[counter_slot] = items of the part
    uint32_t dense_max;           // more flagged words than this in a tile: hand on the tile's spans instead
    uint32_t warm;                // warm-up bytes before a flagged word's end offsets (halo rounded up to W)
    uint32_t want_end_state;      // also compute the state at the end of the stream (counters[2])
    uint32_t clear_tile_len;      // ac_collect_kernel clears tile_len[tile] (first attempt of a call)
    const uint4 *gt_slots;        // exact gram table (gram_table.hpp) or nullptr
    const uint32_t *gt_pat;       // its pattern store
    uint32_t gt_log2;             // 2^gt_log2 slots; 0: every flagged word is walked
    uint2 *count_row;             // asynchronous calls: where ac_offsets_kernel leaves {event count, dense tiles}; else nullptr
    uint32_t *host_counters;      // synchronous calls: pinned host memory that receives counters[0..8) from ac_emit_kernel
                                  // (one copy-engine round trip less per call); else nullptr
    uint32_t *items;              // work items, tile runs in completion order (capacity n_tiles * VER_DENSE_MAX)
    uint2 *desc;                  // per tile {offset into items, count}
    uint2 *recs;                  // per item {state of the first event, count << 16 | first end - item origin}
    uint32_t *tile_len;           // per tile: events (zeroed before the launch; the walk kernel adds to it)
    uint32_t *tile_off;           // per tile: offset of its first event in the event buffer
    uint32_t *block_sum;          // per EMIT_THREADS tiles: events (zeroed before the launch)
};

template <int W>
__global__ void __launch_bounds__(COLLECT_THREADS) ac_collect_kernel(const __grid_constant__ VerifyArgs a)
{
    constexpr int NB = 16 / W;
    constexpr uint32_t WORDS_PER_SPAN = 32u * NB;
    constexpr int N_WARPS = COLLECT_THREADS / 32;
    __shared__ uint32_t s_n[N_WARPS];
    __shared__ uint32_t s_base;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t dense_tiles = 0, flagged = 0;
    // ac_walk_kernel may be put on the SMs from now on (it waits for this grid to finish before it reads a single item)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    auto load_planes = [&](uint32_t tile, uint32_t (&pl)[NB]) {
#pragma unroll
        for (int j = 0; j < NB; ++j) pl[j] = 0;
        const uint32_t span = tile * 32u + lane;
        if (tile < a.tile_end && span < a.n_spans) {
            if (NB == 2) {
                const uint2 m = __ldg(reinterpret_cast<const uint2 *>(a.mask) + span);
                pl[0] = m.x; pl[1] = m.y;
            } else {
                const uint4 m = __ldg(reinterpret_cast<const uint4 *>(a.mask) + span);
                pl[0] = m.x; pl[1] = m.y; pl[NB - 2] = m.z; pl[NB - 1] = m.w;
            }
        }
    };

    // a CTA iteration takes 32 consecutive tiles, one per warp; the loop bound is CTA-uniform
    uint32_t planes[NB], next_planes[NB];
    load_planes(a.tile_begin + blockIdx.x * N_WARPS + warp, planes);
    for (uint32_t tile0 = a.tile_begin + blockIdx.x * N_WARPS; tile0 < a.tile_end; tile0 += gridDim.x * N_WARPS) {
        const uint32_t tile = tile0 + warp;
        load_planes(tile + gridDim.x * N_WARPS, next_planes);      // in flight while this tile is compacted
        const uint32_t span = tile * 32u + lane;
        const bool active = tile < a.tile_end && span < a.n_spans;
        uint32_t cnt = 0;
#pragma unroll
        for (int j = 0; j < NB; ++j) cnt += __popc(planes[j]);
        flagged += cnt;
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        const uint32_t n_cand = __shfl_sync(0xffffffffu, incl, 31);
        const bool dense = n_cand > a.dense_max;       // cheaper to walk the whole tile
        const uint32_t n_act = __popc(__ballot_sync(0xffffffffu, active));
        const uint32_t n = dense ? n_act : n_cand;
        if (lane == 0) s_n[warp] = n;
        __syncthreads();
        // every warp scans the 32 counts; one atomic per CTA iteration reserves the items of all 32 tiles
        const uint32_t v = s_n[lane];
        uint32_t wincl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, wincl, d);
            if (lane >= d) wincl += u;
        }
        const uint32_t cta_total = __shfl_sync(0xffffffffu, wincl, 31);
        if (threadIdx.x == 0) s_base = a.item_base + (cta_total ? atomicAdd(&a.s.counters[a.counter_slot], cta_total) : 0u);
        __syncthreads();
        const uint32_t base = s_base + __shfl_sync(0xffffffffu, wincl - v, warp);
        if (lane == 0 && tile < a.tile_end) {
            a.desc[tile] = make_uint2(base, n);
            if (a.clear_tile_len) a.tile_len[tile] = 0u;     // (the walk kernel adds to it; no memset in front of the call)
        }
        if (n == 0) {
        } else if (dense) {
            ++dense_tiles;
            if (active) a.items[base + lane] = ITEM_SPAN | span;
        } else if (cnt) {
            // flagged words of this lane's span in ascending stream order
            uint32_t at = base + incl - cnt;
            uint32_t any = 0;
#pragma unroll
            for (int j = 0; j < NB; ++j) any |= planes[j];
            while (any) {
                const uint32_t ch = __ffs(any) - 1;
                any &= any - 1;
#pragma unroll
                for (int j = 0; j < NB; ++j)
                    if ((planes[j] >> ch) & 1u) a.items[at++] = span * WORDS_PER_SPAN + ch * NB + j;
            }
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) planes[j] = next_planes[j];
    }
    if (lane == 0 && dense_tiles) atomicAdd(&a.s.counters[4], dense_tiles);
    // flagged words of the stream (a statistic: the filter loop itself does not spend instructions on it)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) flagged += __shfl_xor_sync(0xffffffffu, flagged, d);
    if (lane == 0 && flagged) atomicAdd(&a.s.counters[3], flagged);
}

// --------------------------------------------------------------- walk -----

// W aligned bytes of the stream as 32-bit words (second word unused for W = 4)
template <int W>
__device__ __forceinline__ uint2 ld_group(const uint8_t *text, uint32_t i)
{
    if (W == 8) return __ldg(reinterpret_cast<const uint2 *>(text + i));
    return make_uint2(__ldg(reinterpret_cast<const uint32_t *>(text + i)), 0u);
}

template <int W>
__device__ __forceinline__ uint2 ld_group_guarded(const ScanArgs &a, uint32_t i)
{
    if (i + W <= a.readable) return ld_group<W>(a.text, i);
    uint32_t w[2] = {0, 0};
    for (uint32_t j = 0; j < (uint32_t)W; ++j)
        if (i + j < a.readable) w[j >> 2] |= (uint32_t)a.text[i + j] << ((j & 3u) * 8u);
    return make_uint2(w[0], w[1]);
}

// One automaton step straight from the dense table: the rows of the shallow states a verification walk
// visits stay in L1 (read-only path), deeper rows come from L2.
template <typename E, bool RANGE>
struct Stepper {
    const E *gtab;
    uint32_t s_cls;               // shared-window address of the 256-byte class map (unused when RANGE)
    uint32_t ncls, lo, n_used, final_bound, root;

    __device__ __forceinline__ uint32_t step(uint32_t s, uint32_t b) const
    {
        uint32_t c;
        if (RANGE) c = min(b - lo, n_used);
        else asm("ld.shared.u8 %0, [%1];" : "=r"(c) : "r"(s_cls + b));
        return (uint32_t)__ldg(gtab + (s * ncls + c));
    }
};

// per-item result
struct ItemEvents { uint32_t cnt, e0p, e0s; };

// The W end offsets owned by flagged word k are rs+1 .. rs+W with rs = W(k+1).  K such words are verified in
// lockstep (independent lookup chains hide each other's latency): each walk starts `warm` bytes before rs,
// is reset to the root where its haystack starts (w0[k], a multiple of W inside [rs-warm, rs)), and reports
// the final states reached inside [rs, rs+W).  The caller guarantees that every window [rs-warm, rs+W)
// lies inside the stream and that [w0, rs+W) lies inside one haystack.
template <int W, int K, typename ST>
__device__ __forceinline__ void walk_words_lockstep(const ST &st, const uint8_t *text, uint32_t warm,
                                                    const uint32_t (&rs)[K], const uint32_t (&w0)[K],
                                                    ItemEvents (&ev)[K])
{
    // four groups in flight per walk: a 24-byte window costs ONE memory latency, not three
    uint32_t s[K];
    uint2 cur[K], n1[K], n2[K], n3[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        s[k] = st.root;
        const uint32_t g0 = rs[k] - warm;                    // last group (the report group) starts at rs
        cur[k] = ld_group<W>(text, g0);
        n1[k] = ld_group<W>(text, min(g0 + W, rs[k]));
        n2[k] = ld_group<W>(text, min(g0 + 2u * W, rs[k]));
        n3[k] = ld_group<W>(text, min(g0 + 3u * W, rs[k]));
    }
    for (uint32_t off = warm; off > 0; off -= W) {          // this group starts at rs - off
        uint2 n4[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            n4[k] = n3[k];
            if (off >= 4u * W) n4[k] = ld_group<W>(text, rs[k] - off + 4u * W);
            if (rs[k] - off == w0[k]) s[k] = st.root;       // bytes before the haystack start do not count
        }
#pragma unroll
        for (int j = 0; j < W; ++j) {
#pragma unroll
            for (int k = 0; k < K; ++k)
                s[k] = st.step(s[k], __byte_perm((j < 4) ? cur[k].x : cur[k].y, 0, 0x4440 | (j & 3)));
        }
#pragma unroll
        for (int k = 0; k < K; ++k) { cur[k] = n1[k]; n1[k] = n2[k]; n2[k] = n3[k]; n3[k] = n4[k]; }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) ev[k] = ItemEvents{0, 0, 0};
#pragma unroll
    for (int j = 0; j < W; ++j) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            s[k] = st.step(s[k], __byte_perm((j < 4) ? cur[k].x : cur[k].y, 0, 0x4440 | (j & 3)));
            const bool f = s[k] < st.final_bound;
            if (f && ev[k].cnt == 0) { ev[k].e0p = rs[k] + j + 1u; ev[k].e0s = s[k]; }
            ev[k].cnt += f ? 1u : 0u;
        }
    }
}

// Everything else, out of line (rare): a flagged word whose window is clipped by the ends of the stream or
// starts at an unaligned haystack start, and span items (a 512-byte span of a densely flagged tile, walked
// like an ac_scan_kernel slice).  rs == 0xffffffff: report nothing, return the end state in e0s.
// EMIT: events go to a.out[obase..).
template <typename E, bool RANGE, int W, int EMIT>
__device__ __noinline__ ItemEvents walk_item_slow(const ScanArgs &a, uint32_t s_cls_addr, uint32_t item,
                                                  uint32_t ws, uint32_t rs, uint32_t re, uint32_t obase)
{
    Scanner<E, RANGE, false> sc;
    sc.gtab = static_cast<const E *>(a.table); sc.text = a.text;
    sc.ncls = a.ncls; sc.row_bytes = a.ncls * (uint32_t)sizeof(E);
    sc.win_lo = a.win_lo; sc.win_rows = 0;               // no shared-memory window: every step reads the table
    sc.s_tab = 0; sc.fin_rel = 1;
    sc.s_cls = s_cls_addr;
    sc.lo = a.range_lo; sc.n_used = a.n_used;
    sc.final_bound = a.final_bound; sc.readable = a.readable;
    sc.out = a.out; sc.cap = a.capacity;
    sc.found = false; sc.cnt = 0; sc.obase = obase; sc.have_pend = false; sc.report_from = 0;
    sc.e0p = sc.e0s = sc.e1p = sc.e1s = 0;

    if (item != ITEM_NONE && (item & ITEM_SPAN)) {
        // A span item stands for the WORDS of its span, and word k owns the end offsets W(k+1)+1 .. W(k+2): the
        // bytes to report are the span's shifted by W.  (Reporting the span's own bytes would duplicate the
        // ends owned by the last word of a sparse tile before it and lose those after a dense tile.)
        const uint32_t cs = min((item & ~ITEM_SPAN) * SPAN_BYTES + W, a.total);
        const uint32_t ce = min(cs + SPAN_BYTES, a.total);
        if (cs >= ce) return ItemEvents{0, 0, 0};
        const uint32_t h = find_haystack(a, cs);
        const uint32_t hb = hay_begin(a, h);
        uint32_t w0 = (cs - hb > a.halo) ? ((cs - a.halo) & ~15u) : hb;
        if (w0 < hb) w0 = hb;
        const uint32_t s_cs = sc.template walk<false, false>(a.root, w0, cs);
        scan_slice<EMIT>(a, sc, s_cs, h, cs, ce);
        return ItemEvents{sc.cnt, sc.e0p, sc.e0s};
    }

    uint32_t h = find_haystack(a, ws);
    uint32_t nb = hay_end(a, h);
    uint32_t s = a.root;
    for (uint32_t i = ws; i < re; i += W) {
        const uint2 cur = ld_group_guarded<W>(a, i);
#pragma unroll 1
        for (int j = 0; j < W; ++j) {
            const uint32_t ii = i + j;
            if (ii < re) {
                if (ii == nb) {                      // a haystack starts here
                    do { ++h; nb = hay_end(a, h); } while (nb == ii);
                    s = a.root;
                }
                s = sc.any_next(s, (((j < 4) ? cur.x : cur.y) >> ((j & 3) * 8)) & 0xffu);
                if (ii >= rs && s < a.final_bound) sc.template hit<EMIT>(ii + 1u, s);
            }
        }
    }
    if (rs == 0xffffffffu) return ItemEvents{0, 0, s};
    return ItemEvents{sc.cnt, sc.e0p, sc.e0s};
}

// Direct verification of the flagged word before rs (gram_table.hpp): one table probe and one comparison of the
// haystack with the only pattern that can end in the word's window.  false: the automaton has to be walked.
template <int W, typename LT>
__device__ __forceinline__ bool verify_word_direct(const VerifyArgs &a, uint32_t rs, uint32_t hay_begin, LT load_text, ItemEvents &ev)
{
    typedef typename GramChunk<W>::type chunk_t;
    const uint4 *slots = a.gt_slots;
    const uint32_t *pat = a.gt_pat;
    auto load_slot = [slots](uint32_t i) -> GramSlot {
        const uint4 v = __ldg(slots + 2u * i), t = __ldg(slots + 2u * i + 1u);      // one 32-byte sector
        return GramSlot{v.x, v.y, v.z, v.w, {t.x, t.y, t.z, t.w}};
    };
    auto load_pat = [pat](uint32_t i) -> chunk_t {
        if (W == 8) {
            const uint2 v = __ldg(reinterpret_cast<const uint2 *>(pat + i));
            return (chunk_t)(((uint64_t)v.y << 32) | v.x);
        }
        return (chunk_t)__ldg(pat + i);
    };
    auto load_state = [pat](uint32_t i) -> uint32_t { return __ldg(pat + i); };
    uint32_t end = 0, state = 0;
    const GramVerdict v = gram_verify<W>(rs, a.warm, hay_begin, a.gt_log2, load_text, load_slot, load_pat, load_state, &end, &state);
    if (v == GRAM_NEEDS_WALK) return false;
    ev = (v == GRAM_EVENT) ? ItemEvents{1u, end, state} : ItemEvents{0u, 0u, 0u};
    return true;
}

template <int W>
__device__ __forceinline__ typename GramChunk<W>::type group_chunk(uint2 g)
{
    typedef typename GramChunk<W>::type chunk_t;
    return (W == 8) ? (chunk_t)(((uint64_t)g.y << 32) | g.x) : (chunk_t)g.x;
}

// ... with every group read from the haystack in memory
template <int W>
__device__ __forceinline__ bool verify_word_direct(const VerifyArgs &a, uint32_t rs, uint32_t hay_begin, ItemEvents &ev)
{
    const uint8_t *text = a.s.text;
    return verify_word_direct<W>(a, rs, hay_begin, [text](uint32_t i) { return group_chunk<W>(ld_group<W>(text, i)); }, ev);
}

// origin of an item's end offsets: a record stores its first event's end relative to this
template <int W>
__device__ __forceinline__ uint32_t item_origin(uint32_t item)
{
    return (item & ITEM_SPAN) ? (item & ~ITEM_SPAN) * SPAN_BYTES : (item + 1u) * W;
}

// 16 KiB tile an item lies in
template <int W>
__device__ __forceinline__ uint32_t item_tile(uint32_t item)
{
    return (item & ITEM_SPAN) ? (item & ~ITEM_SPAN) >> 5 : item / (32u * (16u / W) * 32u);
}

template <typename E, bool RANGE, int W>
__global__ void __launch_bounds__(WALK_THREADS) ac_walk_kernel(const __grid_constant__ VerifyArgs a)
{
    __shared__ uint8_t s_cls[256];
    if (threadIdx.x < 256) s_cls[threadIdx.x] = a.s.cls_map[threadIdx.x];
    __syncthreads();
    const uint32_t s_cls_addr = (uint32_t)__cvta_generic_to_shared(s_cls);

    Stepper<E, RANGE> st;
    st.gtab = static_cast<const E *>(a.s.table);
    st.s_cls = s_cls_addr;
    st.ncls = a.s.ncls; st.lo = a.s.range_lo; st.n_used = a.s.n_used;
    st.final_bound = a.s.final_bound; st.root = a.s.root;

    // launched as a programmatic dependent of ac_collect_kernel: everything above overlapped with it, nothing below may
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // the dense list ac_collect_kernel made
    const uint32_t n_items = a.s.counters[a.counter_slot];
    const uint32_t n_threads = gridDim.x * WALK_THREADS;
    constexpr int K = WALK_ILP;
    // a thread takes K consecutive items and walks them in lockstep
    for (uint32_t i0 = (blockIdx.x * WALK_THREADS + threadIdx.x) * K; i0 < n_items; i0 += n_threads * K) {
        uint32_t item[K], rs[K], w0[K], slot[K];
        bool lock[K], valid[K];
        bool all_lock = true;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            slot[k] = a.item_base + i0 + k;
            valid[k] = i0 + k < n_items;
            item[k] = valid[k] ? a.items[slot[k]] : ITEM_NONE;
            rs[k] = 0; w0[k] = 0; lock[k] = false;
            if (item[k] != ITEM_NONE && !(item[k] & ITEM_SPAN)) {
                rs[k] = (item[k] + 1u) * W;            // the W end offsets owned by word k are rs+1 .. rs+W
                if (rs[k] >= a.warm && rs[k] + W <= a.s.total) {
                    // a walk that would start before the haystack of byte rs starts at that haystack's
                    // first byte instead (the state there is the root by definition)
                    const uint32_t h = find_haystack(a.s, rs[k]);
                    w0[k] = max(rs[k] - a.warm, hay_begin(a.s, h));
                    lock[k] = hay_end(a.s, h) >= rs[k] + W && ((rs[k] - w0[k]) & (uint32_t)(W - 1)) == 0 && w0[k] < rs[k];
                }
            }
            all_lock = all_lock && lock[k];
        }
        ItemEvents ev[K];
        // most flagged words are settled by one comparison against the only pattern their gram belongs to
        bool all_direct = a.gt_log2 != 0;
        if (all_direct) {
#pragma unroll
            for (int k = 0; k < K; ++k)
                all_direct = all_direct && lock[k] && verify_word_direct<W>(a, rs[k], w0[k], ev[k]);
        }
        if (all_direct) {
        } else if (all_lock) {
            walk_words_lockstep<W, K>(st, a.s.text, a.warm, rs, w0, ev);
        } else {
#pragma unroll 1
            for (int k = 0; k < K; ++k) {
                uint32_t it = ITEM_NONE, r1 = 0, w1 = 0; bool lk = false;
#pragma unroll
                for (int kk = 0; kk < K; ++kk) if (kk == k) { it = item[kk]; r1 = rs[kk]; w1 = w0[kk]; lk = lock[kk]; }
                ItemEvents e1{0, 0, 0};
                if (lk) {
                    const uint32_t r_[1] = {r1}, w_[1] = {w1};
                    ItemEvents e_[1];
                    walk_words_lockstep<W, 1>(st, a.s.text, a.warm, r_, w_, e_);
                    e1 = e_[0];
                } else if (it != ITEM_NONE) {
                    if (it & ITEM_SPAN) e1 = walk_item_slow<E, RANGE, W, 0>(a.s, s_cls_addr, it, 0u, 0u, 0u, 0u);
                    else if (r1 < a.s.total)           // else nothing ends after this word
                        e1 = walk_item_slow<E, RANGE, W, 0>(a.s, s_cls_addr, it, (r1 > a.warm) ? r1 - a.warm : 0u, r1,
                                                                min(r1 + W, a.s.total), 0u);
                }
#pragma unroll
                for (int kk = 0; kk < K; ++kk) if (kk == k) ev[kk] = e1;
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (valid[k]) {
                const uint32_t rel = ev[k].cnt ? ev[k].e0p - item_origin<W>(item[k]) : 0u;
                a.recs[slot[k]] = make_uint2(ev[k].e0s, (min(ev[k].cnt, 0xffffu) << 16) | (rel & 0xffffu));
                if (ev[k].cnt) atomicAdd(&a.tile_len[item_tile<W>(item[k])], ev[k].cnt);   // events per tile, for the offsets
            }
        }
        // events per block of EMIT_THREADS tiles: the items of a warp almost always lie in one block, so one
        // atomic per warp (same-address atomics serialise)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const bool has = valid[k] && ev[k].cnt;
            const uint32_t blk = has ? item_tile<W>(item[k]) / EMIT_THREADS : 0xffffffffu;
            const uint32_t act = __activemask();
            const uint32_t voters = __ballot_sync(act, has);
            if (voters) {
                const uint32_t lead_blk = __shfl_sync(act, blk, __ffs(voters) - 1);
                if (__all_sync(act, !has || blk == lead_blk)) {
                    uint32_t sum = has ? ev[k].cnt : 0u;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(act, sum, d);
                    if ((threadIdx.x & 31u) == (uint32_t)(__ffs(act) - 1)) atomicAdd(&a.block_sum[lead_blk], sum);
                } else if (has) {
                    atomicAdd(&a.block_sum[blk], ev[k].cnt);
                }
            }
        }
    }

    // state at the end of the stream (keep=1 continuation): the last Lmax bytes decide it
    if (a.want_end_state && blockIdx.x == gridDim.x - 1 && threadIdx.x == WALK_THREADS - 1) {
        const uint32_t back = a.s.halo + 1u;
        const uint32_t ws = (a.s.total > back) ? ((a.s.total - back) & ~(uint32_t)(W - 1)) : 0u;
        a.s.counters[2] = walk_item_slow<E, RANGE, W, 0>(a.s, s_cls_addr, ITEM_NONE, ws, 0xffffffffu, a.s.total, 0u).e0s;
    }
}

// --------------------------------------------------------------- emit -----

// Exclusive prefix sum of the events per tile.  One CTA per 256 tiles: it adds up the earlier blocks' sums
// itself (no CTA waits for another one) and scans its own 256 tile lengths.  tile_len -> tile_off.
__global__ void __launch_bounds__(EMIT_THREADS) ac_offsets_kernel(const __grid_constant__ VerifyArgs a)
{
    __shared__ uint32_t s_warp[EMIT_THREADS / 32];
    __shared__ uint32_t s_prev[EMIT_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // ac_emit_kernel may be put on the SMs (it waits for this grid)

    uint32_t part = 0;
    for (uint32_t j = tid; j < blockIdx.x; j += EMIT_THREADS) part += a.block_sum[j];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if (lane == 0) s_prev[warp] = part;

    const uint32_t tile = blockIdx.x * EMIT_THREADS + tid;
    const uint32_t len = (tile < a.n_tiles) ? a.tile_len[tile] : 0u;
    uint32_t incl = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t base = 0;
#pragma unroll
    for (int w = 0; w < EMIT_THREADS / 32; ++w) {
        base += s_prev[w];
        if ((uint32_t)w < warp) base += s_warp[w];
    }
    if (tile < a.n_tiles) a.tile_off[tile] = base + incl - len;
    if (tile == a.n_tiles - 1) {
        a.s.counters[1] = base + incl;                                // all events of the call
        // asynchronous calls: row 0 of the caller's rows = {event count, densely flagged tiles} (ac_collect_kernel, which
        // counts the latter, has finished)
        if (a.count_row) *a.count_row = make_uint2(base + incl, a.s.counters[4]);
    }
}

// One warp per tile with events: write them to their final place, in item order.
template <typename E, bool RANGE, int W>
__global__ void __launch_bounds__(COUNT_THREADS) ac_emit_kernel(const __grid_constant__ VerifyArgs a)
{
    __shared__ uint8_t s_cls[256];
    const uint32_t lane = threadIdx.x & 31u;
    if (threadIdx.x < 256) s_cls[threadIdx.x] = a.s.cls_map[threadIdx.x];
    __syncthreads();
    const uint32_t s_cls_addr = (uint32_t)__cvta_generic_to_shared(s_cls);
    const uint32_t n_warps = gridDim.x * (COUNT_THREADS / 32);
    asm volatile("griddepcontrol.wait;" ::: "memory");      // launched as a programmatic dependent of ac_offsets_kernel
    // the call's counters are final (every kernel that adds to them has finished): hand them to the host
    if (a.host_counters && blockIdx.x == 0 && threadIdx.x < 8) {
        a.host_counters[threadIdx.x] = a.s.counters[threadIdx.x];
        __threadfence_system();
    }

    // The kernel is a chain of dependent loads per tile (events of the tile -> descriptor + offset -> records -> items).
    // A warp looks two tiles ahead for the event count and one tile ahead for descriptor + offset (only where the
    // count says something ends there), and fetches the records and items of a round together.
    auto load_len = [&](uint32_t t) -> uint32_t { return (t < a.n_tiles) ? a.tile_len[t] : 0u; };
    auto load_header = [&](uint32_t t, uint32_t len, uint2 &d, uint32_t &off) {
        d = make_uint2(0u, 0u); off = 0;
        if (len) { d = a.desc[t]; off = a.tile_off[t]; }
    };
    uint32_t tile = blockIdx.x * (COUNT_THREADS / 32) + (threadIdx.x >> 5);
    uint32_t len = load_len(tile), next_len = load_len(tile + n_warps), off;
    uint2 d;
    load_header(tile, len, d, off);
    for (; tile < a.n_tiles; tile += n_warps) {
        const uint32_t next2_len = load_len(tile + 2u * n_warps);
        uint32_t next_off;
        uint2 next_d;
        load_header(tile + n_warps, next_len, next_d, next_off);
        const uint32_t n_items = len ? d.y : 0u;                      // warp-uniform; 0: nothing ends in this tile
        for (uint32_t i0 = 0; i0 < n_items; i0 += 32u) {             // warp-uniform
            const uint32_t i = i0 + lane;
            uint32_t item = ITEM_NONE, cnt = 0;
            uint2 rec = make_uint2(0u, 0u);
            if (i < n_items) {
                rec = a.recs[d.x + i];
                item = a.items[d.x + i];
                cnt = rec.y >> 16;
            }
            if (!__any_sync(0xffffffffu, cnt != 0)) continue;
            uint32_t pincl = cnt;
#pragma unroll
            for (int k = 1; k < 32; k <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, pincl, k);
                if (lane >= k) pincl += v;
            }
            const uint32_t o = off + pincl - cnt;
            if (cnt == 1) {
                if (o < a.s.capacity) a.s.out[o] = make_uint2(item_origin<W>(item) + (rec.y & 0xffffu), rec.x);
            } else if (cnt > 1 && o < a.s.capacity) {
                uint32_t ws = 0, rs = 0, re = 0;
                if (!(item & ITEM_SPAN)) {
                    rs = (item + 1u) * W;
                    re = min(rs + W, a.s.total);
                    ws = (rs > a.warm) ? rs - a.warm : 0u;
                }
                walk_item_slow<E, RANGE, W, 1>(a.s, s_cls_addr, item, ws, rs, re, o);
            }
            off += __shfl_sync(0xffffffffu, pincl, 31);
        }
        len = next_len; d = next_d; off = next_off; next_len = next2_len;
    }
}

} // namespace acb200
