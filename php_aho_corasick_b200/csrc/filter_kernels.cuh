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

namespace acb200 {

constexpr uint32_t SPAN_BYTES = 512;       // one warp-wide 16-byte load; one verify lane
constexpr int FILTER_UNROLL = 4;           // 16-byte loads in flight per thread
constexpr uint32_t VER_DENSE_MAX = 64;     // flagged words per 16 KiB tile beyond which the whole tile is walked
constexpr uint32_t ITEM_SPAN = 0x80000000u;// work item: walk 512-byte span (item & ~ITEM_SPAN) completely
constexpr uint32_t ITEM_NONE = 0xffffffffu;
constexpr int COLLECT_THREADS = 256;
constexpr int WALK_THREADS = 256;
constexpr int EMIT_THREADS = 1024;         // one thread per tile in the offset scan, one warp per 32 tiles when emitting

struct FilterArgs {
    const uint8_t *text;          // 16-byte aligned
    uint32_t total;               // bytes in the stream
    const uint32_t *l1;           // level-1 bitmap (FILTER_L1_BITS bits)
    uint32_t l1_bits;
    const uint32_t *l2;           // level-2 bitmap or nullptr
    uint32_t l2_shift;            // 32 - log2(bits of level 2)
    uint32_t *mask;               // n_spans x (16/W) words: plane j bit c <=> word (c*(16/W) + j) of the span
    uint32_t n_spans;             // ceil(total / 512)
    uint32_t *counters;           // [3] += flagged words
};

// ------------------------------------------------------------- filter -----

template <int W, bool L2>
__global__ void __launch_bounds__(SCAN_THREADS, 1) ac_filter_kernel(const FilterArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *s_bm = reinterpret_cast<uint32_t *>(smem_raw);
    constexpr int NB = 16 / W;

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.l1);
        uint4 *dst = reinterpret_cast<uint4 *>(s_bm);
        const uint32_t n4 = a.l1_bits >> 7;
        for (uint32_t i = tid; i < n4; i += SCAN_THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(s_bm);

    // one flag per aligned word: both bits of the word's gram hash are set in the level-1 bitmap
    auto test_word = [&](uint32_t lo, uint32_t hi) -> bool {
        const uint32_t t = filter_mix1(lo, hi);
        const uint32_t idx = filter_reduce(t, a.l1_bits);
        uint32_t word;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(word) : "r"(s_base + ((idx >> 5) << 2)));
        bool p = ((word >> (idx & 31u)) & (word >> filter_bit2(t)) & 1u) != 0;
        if (L2) {
            uint32_t word3 = 0;
            const uint32_t i3 = filter_mix3(lo, hi) >> a.l2_shift;
            if (p) word3 = __ldg(a.l2 + (i3 >> 5));
            p = (word3 >> (i3 & 31u)) & 1u;
        }
        return p;
    };

    const uint32_t n_full = a.total / SPAN_BYTES;             // spans that lie completely inside the stream
    const uint32_t n_warps = gridDim.x * (SCAN_THREADS / 32);
    const uint32_t warp = blockIdx.x * (SCAN_THREADS / 32) + (tid >> 5);
    uint32_t flagged = 0;

    for (uint32_t g0 = warp; g0 < n_full; g0 += n_warps * FILTER_UNROLL) {
        uint4 v[FILTER_UNROLL];
#pragma unroll
        for (int u = 0; u < FILTER_UNROLL; ++u) {
            const uint32_t g = g0 + u * n_warps;
            v[u] = make_uint4(0, 0, 0, 0);
            if (g < n_full) v[u] = ld_text16(a.text + ((size_t)g * 32u + lane) * 16u);
        }
#pragma unroll
        for (int u = 0; u < FILTER_UNROLL; ++u) {
            const uint32_t g = g0 + u * n_warps;
            if (g >= n_full) break;                           // warp-uniform
            const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
            uint32_t mine = 0;
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                const bool p = (W == 8) ? test_word(w[2 * j], w[2 * j + 1]) : test_word(w[j], 0u);
                const uint32_t plane = __ballot_sync(0xffffffffu, p);
                if (lane == (uint32_t)j) mine = plane;
            }
            if (lane < (uint32_t)NB) {
                a.mask[(size_t)g * NB + lane] = mine;
                flagged += __popc(mine);
            }
        }
    }

    // The last, partial span: complete 16-byte chunks are tested, the partial chunk at the very end is not
    // read at all — its words are simply handed on to verification.
    if (n_full < a.n_spans && warp == (n_full % n_warps)) {
        const uint32_t n16 = a.total >> 4;
        const uint32_t c = n_full * 32u + lane;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (c < n16) v = ld_text16(a.text + (size_t)c * 16u);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        const bool tail = (a.total & 15u) && c == n16;
        uint32_t mine = 0;
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            bool p = (W == 8) ? test_word(w[2 * j], w[2 * j + 1]) : test_word(w[j], 0u);
            p = (c < n16) ? p : tail;
            const uint32_t plane = __ballot_sync(0xffffffffu, p);
            if (lane == (uint32_t)j) mine = plane;
        }
        if (lane < (uint32_t)NB) {
            a.mask[(size_t)n_full * NB + lane] = mine;
            flagged += __popc(mine);
        }
    }
    if (lane < (uint32_t)NB && flagged) atomicAdd(&a.counters[3], flagged);
}

// ------------------------------------------------------------ collect -----

struct VerifyArgs {
    ScanArgs s;                   // stream, automaton, event buffer (win_rows = 0: no shared-memory window)
    const uint32_t *mask;         // bit planes written by ac_filter_kernel
    uint32_t n_spans;             // 512-byte spans in the stream
    uint32_t n_tiles;             // 16 KiB tiles = ceil(n_spans / 32)
    uint32_t dense_max;           // more flagged words than this in a tile: hand on the tile's spans instead
    uint32_t warm;                // warm-up bytes before a flagged word's end offsets (halo rounded up to W)
    uint32_t want_end_state;      // also compute the state at the end of the stream (counters[2])
    uint32_t *items;              // work items, tile runs in completion order (capacity n_tiles * VER_DENSE_MAX)
    uint2 *desc;                  // per tile {offset into items, count}
    uint2 *recs;                  // per item {state of the first event, count << 16 | first end - item origin}
    uint32_t *tile_len;           // per tile: events
};

template <int W>
__global__ void __launch_bounds__(COLLECT_THREADS) ac_collect_kernel(const __grid_constant__ VerifyArgs a)
{
    constexpr int NB = 16 / W;
    constexpr uint32_t WORDS_PER_SPAN = 32u * NB;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_warps = gridDim.x * (COLLECT_THREADS / 32);
    uint32_t dense_tiles = 0;

    for (uint32_t tile = blockIdx.x * (COLLECT_THREADS / 32) + (threadIdx.x >> 5); tile < a.n_tiles; tile += n_warps) {
        const uint32_t span = tile * 32u + lane;
        const bool active = span < a.n_spans;
        uint32_t planes[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) planes[j] = 0;
        if (active) {
            if (NB == 2) {
                const uint2 m = __ldg(reinterpret_cast<const uint2 *>(a.mask) + span);
                planes[0] = m.x; planes[1] = m.y;
            } else {
                const uint4 m = __ldg(reinterpret_cast<const uint4 *>(a.mask) + span);
                planes[0] = m.x; planes[1] = m.y; planes[NB - 2] = m.z; planes[NB - 1] = m.w;
            }
        }
        uint32_t cnt = 0;
#pragma unroll
        for (int j = 0; j < NB; ++j) cnt += __popc(planes[j]);
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
        uint32_t base = 0;
        if (lane == 0) {
            if (n) base = atomicAdd(&a.s.counters[5], n);
            a.desc[tile] = make_uint2(base, n);
        }
        if (n == 0) continue;                           // warp-uniform
        base = __shfl_sync(0xffffffffu, base, 0);
        if (dense) {
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
    }
    if (lane == 0 && dense_tiles) atomicAdd(&a.s.counters[4], dense_tiles);
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

// The W end offsets owned by flagged word k are rs+1 .. rs+W with rs = W(k+1): walk from the root over the
// warm-up [w0, rs) — w0 a multiple of W, the whole window [w0, rs+W) inside one haystack and inside the
// stream (the caller checked) — and report the final states reached inside [rs, rs+W).
template <int W, typename ST>
__device__ __forceinline__ ItemEvents walk_word_fast(const ST &st, const uint8_t *text, uint32_t w0, uint32_t rs)
{
    uint32_t s = st.root;
    uint2 cur = ld_group<W>(text, w0);
    for (uint32_t i = w0; i < rs; i += W) {
        const uint2 nxt = ld_group<W>(text, i + W);        // the last one is the report group
#pragma unroll
        for (int j = 0; j < W; ++j)
            s = st.step(s, __byte_perm((j < 4) ? cur.x : cur.y, 0, 0x4440 | (j & 3)));
        cur = nxt;
    }
    ItemEvents ev{0, 0, 0};
#pragma unroll
    for (int j = 0; j < W; ++j) {
        s = st.step(s, __byte_perm((j < 4) ? cur.x : cur.y, 0, 0x4440 | (j & 3)));
        const bool f = s < st.final_bound;
        if (f && ev.cnt == 0) { ev.e0p = rs + j + 1u; ev.e0s = s; }
        ev.cnt += f ? 1u : 0u;
    }
    return ev;
}

// Everything else, out of line (rare): a flagged word whose window is clipped by the ends of the stream or
// starts at an unaligned haystack start, and span items (a 512-byte span of a densely flagged tile, walked
// like an ac_scan_kernel slice).  rs == 0xffffffff: report nothing, return the end state in e0s.
// EMIT: events go to a.out[obase..).
template <typename E, bool RANGE, int W, bool EMIT>
__device__ __noinline__ ItemEvents walk_item_slow(const ScanArgs &a, uint32_t s_cls_addr, uint32_t item,
                                                  uint32_t ws, uint32_t rs, uint32_t re, uint32_t obase)
{
    Scanner<E, RANGE, false> sc;
    sc.gtab = static_cast<const E *>(a.table); sc.text = a.text;
    sc.ncls = a.ncls; sc.row_bytes = a.ncls * (uint32_t)sizeof(E);
    sc.win_lo = a.win_lo; sc.win_rows = 0;               // no shared-memory window: every step reads the table
    sc.s_tab = 0;
    sc.s_cls = s_cls_addr;
    sc.lo = a.range_lo; sc.n_used = a.n_used;
    sc.final_bound = a.final_bound; sc.readable = a.readable;
    sc.out = a.out; sc.cap = a.capacity;
    sc.found = false; sc.cnt = 0; sc.obase = obase;
    sc.e0p = sc.e0s = sc.e1p = sc.e1s = 0;

    if (item != ITEM_NONE && (item & ITEM_SPAN)) {
        const uint32_t cs = (item & ~ITEM_SPAN) * SPAN_BYTES;
        const uint32_t ce = min(cs + SPAN_BYTES, a.total);
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

// origin of an item's end offsets: a record stores its first event's end relative to this
template <int W>
__device__ __forceinline__ uint32_t item_origin(uint32_t item)
{
    return (item & ITEM_SPAN) ? (item & ~ITEM_SPAN) * SPAN_BYTES : (item + 1u) * W;
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

    const uint32_t n_items = a.s.counters[5];
    const uint32_t n_threads = gridDim.x * WALK_THREADS;
    for (uint32_t i = blockIdx.x * WALK_THREADS + threadIdx.x; i < n_items; i += n_threads) {
        const uint32_t item = a.items[i];
        ItemEvents ev{0, 0, 0};
        if (item & ITEM_SPAN) {
            ev = walk_item_slow<E, RANGE, W, false>(a.s, s_cls_addr, item, 0u, 0u, 0u, 0u);
        } else {
            const uint32_t rs = (item + 1u) * W;       // the W end offsets owned by word k are rs+1 .. rs+W
            if (rs < a.s.total) {                      // else nothing ends after this word
                const uint32_t re = min(rs + W, a.s.total);
                const uint32_t ws = (rs > a.warm) ? rs - a.warm : 0u;
                // a walk that would start before the haystack of byte rs starts at that haystack's
                // first byte instead (the state there is the root by definition)
                const uint32_t h = find_haystack(a.s, rs);
                const uint32_t w0 = max(ws, hay_begin(a.s, h));
                const bool plain = (re == rs + W) && hay_end(a.s, h) >= re &&
                                   ((rs - w0) & (uint32_t)(W - 1)) == 0 && w0 < rs;
                if (plain) ev = walk_word_fast<W>(st, a.s.text, w0, rs);
                else ev = walk_item_slow<E, RANGE, W, false>(a.s, s_cls_addr, item, ws, rs, re, 0u);
            }
        }
        const uint32_t rel = ev.cnt ? ev.e0p - item_origin<W>(item) : 0u;
        a.recs[i] = make_uint2(ev.e0s, (min(ev.cnt, 0xffffu) << 16) | (rel & 0xffffu));
    }

    // state at the end of the stream (keep=1 continuation): the last Lmax bytes decide it
    if (a.want_end_state && blockIdx.x == gridDim.x - 1 && threadIdx.x == WALK_THREADS - 1) {
        const uint32_t back = a.s.halo + 1u;
        const uint32_t ws = (a.s.total > back) ? ((a.s.total - back) & ~(uint32_t)(W - 1)) : 0u;
        a.s.counters[2] = walk_item_slow<E, RANGE, W, false>(a.s, s_cls_addr, ITEM_NONE, ws, 0xffffffffu, a.s.total, 0u).e0s;
    }
}

// --------------------------------------------------------------- emit -----

// events per tile: one warp per tile sums the counts of the tile's items
__global__ void __launch_bounds__(COLLECT_THREADS) ac_tile_count_kernel(const __grid_constant__ VerifyArgs a)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_warps = gridDim.x * (COLLECT_THREADS / 32);
    for (uint32_t tile = blockIdx.x * (COLLECT_THREADS / 32) + (threadIdx.x >> 5); tile < a.n_tiles; tile += n_warps) {
        const uint2 d = a.desc[tile];
        uint32_t sum = 0;
        for (uint32_t i = lane; i < d.y; i += 32u) sum += a.recs[d.x + i].y >> 16;
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, k);
        if (lane == 0) a.tile_len[tile] = sum;
    }
}

// One CTA per 1024 tiles.  A tile's first event goes to (events of all earlier tiles): the CTA adds up the
// earlier blocks' tile lengths itself (a coalesced read of at most n_tiles words from L2), scans its own
// 1024 lengths, then each warp writes the events of its 32 tiles in item order.
template <typename E, bool RANGE, int W>
__global__ void __launch_bounds__(EMIT_THREADS) ac_emit_kernel(const __grid_constant__ VerifyArgs a)
{
    __shared__ uint32_t s_warp[EMIT_THREADS / 32];
    __shared__ uint32_t s_prev[EMIT_THREADS / 32];
    __shared__ uint32_t s_off[EMIT_THREADS];
    __shared__ uint8_t s_cls[256];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t first = blockIdx.x * EMIT_THREADS;
    if (tid < 256) s_cls[tid] = a.s.cls_map[tid];
    const uint32_t s_cls_addr = (uint32_t)__cvta_generic_to_shared(s_cls);

    uint32_t part = 0;
    for (uint32_t j = tid; j < first; j += EMIT_THREADS) part += a.tile_len[j];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if (lane == 0) s_prev[warp] = part;

    const uint32_t my_tile = first + tid;
    const uint32_t len = (my_tile < a.n_tiles) ? a.tile_len[my_tile] : 0u;
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
    s_off[tid] = base + incl - len;
    if (my_tile == a.n_tiles - 1) a.s.counters[1] = base + incl;     // all events of the call
    __syncwarp();

    const uint32_t busy = __ballot_sync(0xffffffffu, len != 0);
    for (uint32_t m = busy; m; m &= m - 1) {
        const uint32_t t = warp * 32u + (__ffs(m) - 1);
        const uint2 d = a.desc[first + t];
        uint32_t off = s_off[t];
        for (uint32_t i0 = 0; i0 < d.y; i0 += 32u) {                 // warp-uniform
            const uint32_t i = i0 + lane;
            uint32_t item = ITEM_NONE, cnt = 0;
            uint2 rec = make_uint2(0u, 0u);
            if (i < d.y) {
                rec = a.recs[d.x + i];
                cnt = rec.y >> 16;
                if (cnt) item = a.items[d.x + i];
            }
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
                walk_item_slow<E, RANGE, W, true>(a.s, s_cls_addr, item, ws, rs, re, o);
            }
            off += __shfl_sync(0xffffffffu, pincl, 31);
        }
    }
}

} // namespace acb200
