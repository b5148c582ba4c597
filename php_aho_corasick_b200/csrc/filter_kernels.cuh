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
//   ac_verify_kernel  a warp takes a 64 KiB tile of the stream at a time and needs
//                     no other warp: its lanes turn the tile's bit planes into
//                     an ordered list of work items (ballot / popc / prefix
//                     sums) — one item per flagged word, or, when so many words
//                     are flagged that walking all of the tile is cheaper, one
//                     item per 512-byte span.  Then one lane per item, two
//                     items per lane in lockstep (independent lookup chains
//                     hide each other's latency): for a
//                     flagged word the lane walks the automaton from the root
//                     over the (Lmax-1)-byte warm-up plus the W bytes after the
//                     word — exactly the halo argument of ac_scan_kernel, so
//                     states and events are those of an uninterrupted walk; a
//                     span item is walked like an ac_scan_kernel slice.  The
//                     tile's events are written as one ordered run at an offset
//                     taken from a global counter (no warp ever waits for
//                     another one).
//   ac_runs_scan_kernel / ac_reorder_kernel
//                     prefix-sum the run lengths in tile order and copy the
//                     runs there: the final event list is ascending, as the
//                     callback contract requires.  The worst case costs what
//                     the plain scan costs.
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
constexpr int VERIFY_THREADS = 512;        // 16 warps: room for 128 registers per thread
#ifndef ACB_VT_SUB
#define ACB_VT_SUB 4
#endif
#ifndef ACB_VT_LOCK
#define ACB_VT_LOCK 2
#endif
constexpr int VT_SUB = ACB_VT_SUB;         // 16 KiB sub-tiles per warp tile: a lane owns that many 512-byte spans
constexpr int VT_LOCK = ACB_VT_LOCK;       // batches walked in lockstep
constexpr int VT_BATCHES = VT_SUB * (int)VER_DENSE_MAX / 32;   // 8: batches of 32 items per warp tile at most
constexpr uint32_t VT_LIST_CAP = VT_SUB * VER_DENSE_MAX;       // items per warp tile
constexpr uint32_t VER_FIXED_SMEM = (VERIFY_THREADS / 32) * VT_LIST_CAP * 4u;   // per-warp item lists
static_assert(VT_BATCHES % VT_LOCK == 0 && VT_SUB % VT_LOCK == 0, "lockstep groups must tile the batches and the sub-tiles");
static_assert(VER_FIXED_SMEM % 16 == 0, "table window must stay 16-byte aligned");
constexpr int REORDER_THREADS = 256;
constexpr int RUNSCAN_THREADS = 1024;

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

// ------------------------------------------------------------- verify -----

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

// true table entry from HBM/L2, only where the shared-memory window answered 0
template <typename E> __device__ __forceinline__ void ldg_if_zero(uint32_t &e, const E *gtab, uint32_t s, uint32_t ncls, uint32_t c);
template <> __device__ __forceinline__ void ldg_if_zero<uint16_t>(uint32_t &e, const uint16_t *gtab, uint32_t s, uint32_t ncls, uint32_t c)
{
    asm("{\n\t.reg .pred p;\n\t.reg .u32 t;\n\t.reg .u64 a;\n\t"
                 "setp.eq.u32 p, %0, 0;\n\t"
                 "@p mad.lo.u32 t, %2, %3, %4;\n\t"
                 "@p mad.wide.u32 a, t, 2, %1;\n\t"
                 "@p ld.global.nc.u16 %0, [a];\n\t}"
                 : "+r"(e) : "l"(gtab), "r"(s), "r"(ncls), "r"(c));
}
template <> __device__ __forceinline__ void ldg_if_zero<uint32_t>(uint32_t &e, const uint32_t *gtab, uint32_t s, uint32_t ncls, uint32_t c)
{
    asm("{\n\t.reg .pred p;\n\t.reg .u32 t;\n\t.reg .u64 a;\n\t"
                 "setp.eq.u32 p, %0, 0;\n\t"
                 "@p mad.lo.u32 t, %2, %3, %4;\n\t"
                 "@p mad.wide.u32 a, t, 4, %1;\n\t"
                 "@p ld.global.nc.u32 %0, [a];\n\t}"
                 : "+r"(e) : "l"(gtab), "r"(s), "r"(ncls), "r"(c));
}

// Branch-free automaton step for the verify kernel.  Row `win_rows` of the shared-memory window is all
// zero, states outside the window are clamped onto it, and a zero entry means "ask the full table".
template <typename E, bool RANGE>
struct Stepper {
    const E *gtab;
    uint32_t s_tab;          // shared-window byte address of row win_lo
    uint32_t s_cls;
    uint32_t row_bytes, ncls, win_lo, win_rows, lo, n_used, final_bound, root;

    __device__ __forceinline__ uint32_t step(uint32_t s, uint32_t b) const
    {
        uint32_t c;
        if (RANGE) c = min(b - lo, n_used);
        else asm("ld.shared.u8 %0, [%1];" : "=r"(c) : "r"(s_cls + b));
        const uint32_t row = min(s - win_lo, win_rows);
        uint32_t e;
        if (sizeof(E) == 2) asm("ld.shared.u16 %0, [%1];" : "=r"(e) : "r"(s_tab + row * row_bytes + c * 2u));
        else asm("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(s_tab + row * row_bytes + c * 4u));
        ldg_if_zero<E>(e, gtab, s, ncls, c);
        return e;
    }
};

// per-lane event record of one item
struct ItemEvents { uint32_t cnt, e0p, e0s; };

// The W end offsets owned by flagged word k are rs+1 .. rs+W with rs = W(k+1).  K such words are verified in
// lockstep: each walk starts `warm` bytes before rs, is reset to the root where its haystack starts (w0[k], a
// multiple of W inside [rs-warm, rs)), and reports the final states reached inside [rs, rs+W).  The caller
// guarantees that every window [rs-warm, rs+W) lies inside the stream and that [w0, rs+W) lies inside one
// haystack.  Unused slots simply repeat a valid walk and are ignored.
template <int W, int K, typename ST>
__device__ __forceinline__ void walk_words_lockstep(const ST &st, const uint8_t *text, uint32_t warm,
                                                    const uint32_t (&rs)[K], const uint32_t (&w0)[K],
                                                    ItemEvents (&ev)[K])
{
    uint32_t s[K];
    uint2 cur[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        s[k] = st.root;
        cur[k] = ld_group<W>(text, rs[k] - warm);
    }
    for (uint32_t off = warm; off > 0; off -= W) {          // this group starts at rs - off
        uint2 nxt[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            nxt[k] = ld_group<W>(text, rs[k] - off + W);    // the next group (the last one is the report group)
            if (rs[k] - off == w0[k]) s[k] = st.root;       // bytes before the haystack start do not count
        }
#pragma unroll
        for (int j = 0; j < W; ++j) {
#pragma unroll
            for (int k = 0; k < K; ++k)
                s[k] = st.step(s[k], __byte_perm((j < 4) ? cur[k].x : cur[k].y, 0, 0x4440 | (j & 3)));
        }
#pragma unroll
        for (int k = 0; k < K; ++k) cur[k] = nxt[k];
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

// Everything else, out of line (rare): a flagged word whose window contains a haystack start or is clipped
// by the end of the stream, and span items (a 512-byte span of a densely flagged tile, walked like an
// ac_scan_kernel slice).  rs == 0xffffffff: report nothing, return the end state in e0s (end-state walk).
template <typename E, bool RANGE, int W, bool EMIT>
__device__ __noinline__ ItemEvents walk_item_slow(const ScanArgs &a, uint32_t s_tab_addr, uint32_t s_cls_addr,
                                                  uint32_t item, uint32_t ws, uint32_t rs, uint32_t re, uint32_t obase)
{
    Scanner<E, RANGE, false> sc;
    sc.gtab = static_cast<const E *>(a.table); sc.text = a.text;
    sc.ncls = a.ncls; sc.row_bytes = a.ncls * (uint32_t)sizeof(E);
    sc.win_lo = a.win_lo; sc.win_rows = a.win_rows;
    sc.s_tab = s_tab_addr - a.win_lo * sc.row_bytes;
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
                const uint32_t b = (((j < 4) ? cur.x : cur.y) >> ((j & 3) * 8)) & 0xffu;
                if (s - sc.win_lo < sc.win_rows) {
                    const uint32_t e = sc.hot_next(s, b);
                    s = e ? e : sc.any_next(s, b);
                } else {
                    s = sc.any_next(s, b);
                }
                if (ii >= rs && s < a.final_bound) sc.template hit<EMIT>(ii + 1u, s);
            }
        }
    }
    if (rs == 0xffffffffu) return ItemEvents{0, 0, s};
    return ItemEvents{sc.cnt, sc.e0p, sc.e0s};
}

template <typename E, bool RANGE, int W>
__global__ void __launch_bounds__(VERIFY_THREADS, 1) ac_verify_kernel(const __grid_constant__ ScanArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint8_t s_cls[256];
    constexpr int NB = 16 / W;
    constexpr uint32_t WORDS_PER_SPAN = 32u * NB;
    constexpr uint32_t TILE_SPANS = 32u * VT_SUB;

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;

    // dynamic shared memory: per-warp item lists, then the table window
    uint32_t *my_list = reinterpret_cast<uint32_t *>(smem_raw) + (tid >> 5) * VT_LIST_CAP;
    E *s_tab = reinterpret_cast<E *>(smem_raw + VER_FIXED_SMEM);
    const E *gtab = static_cast<const E *>(a.table);

    // window rows (targets outside the window replaced by 0) plus one all-zero row behind them
    {
        constexpr uint32_t PER = 16 / sizeof(E);           // entries per 16-byte load
        const uint32_t win_entries = a.win_rows * a.ncls;
        const uint32_t win_first = a.win_lo * a.ncls;
        const uint32_t lead = min((PER - (win_first % PER)) % PER, win_entries);   // entries before the first aligned group
        for (uint32_t idx = tid; idx < lead; idx += VERIFY_THREADS) {
            uint32_t e = gtab[win_first + idx];
            if (e - a.win_lo >= a.win_rows) e = 0;
            s_tab[idx] = (E)e;
        }
        const uint32_t n_vec = (win_entries - lead) / PER;
        const uint4 *src = reinterpret_cast<const uint4 *>(gtab + win_first + lead);
        for (uint32_t v = tid; v < n_vec; v += VERIFY_THREADS) {
            const uint4 q = __ldg(src + v);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (sizeof(E) == 2) {
                    uint32_t lo = w[k] & 0xffffu, hi = w[k] >> 16;
                    if (lo - a.win_lo >= a.win_rows) lo = 0;
                    if (hi - a.win_lo >= a.win_rows) hi = 0;
                    s_tab[lead + v * PER + 2 * k] = (E)lo;
                    s_tab[lead + v * PER + 2 * k + 1] = (E)hi;
                } else {
                    uint32_t e = w[k];
                    if (e - a.win_lo >= a.win_rows) e = 0;
                    s_tab[lead + v * PER + k] = (E)e;
                }
            }
        }
        for (uint32_t idx = lead + n_vec * PER + tid; idx < win_entries + a.ncls; idx += VERIFY_THREADS) {
            uint32_t e = 0;
            if (idx < win_entries) {
                e = gtab[win_first + idx];
                if (e - a.win_lo >= a.win_rows) e = 0;
            }
            s_tab[idx] = (E)e;
        }
    }
    if (tid < 256) s_cls[tid] = a.cls_map[tid];
    __syncthreads();          // the only CTA-wide barrier: from here on warps run independently

    Stepper<E, RANGE> st;
    st.gtab = gtab;
    st.row_bytes = a.ncls * (uint32_t)sizeof(E); st.ncls = a.ncls;
    st.win_lo = a.win_lo; st.win_rows = a.win_rows;
    st.lo = a.range_lo; st.n_used = a.n_used;
    st.final_bound = a.final_bound; st.root = a.root;
    const uint32_t s_tab_addr = (uint32_t)__cvta_generic_to_shared(s_tab);
    const uint32_t s_cls_addr = (uint32_t)__cvta_generic_to_shared(s_cls);
    asm volatile("mov.u32 %0, %1;" : "=r"(st.s_tab) : "r"(s_tab_addr));
    asm volatile("mov.u32 %0, %1;" : "=r"(st.s_cls) : "r"(s_cls_addr));

    const uint32_t n_tiles = (a.n_spans + TILE_SPANS - 1) / TILE_SPANS;
    const uint32_t n_warps = gridDim.x * (VERIFY_THREADS / 32);
    const bool lockstep_ok = a.total >= a.warm + 2u * W;    // the stand-in walk of unused slots must be in bounds
    uint32_t dense_tiles = 0;

    auto load_planes = [&](uint32_t tile, uint32_t (&pl)[VT_SUB][NB]) {
#pragma unroll
        for (int q = 0; q < VT_SUB; ++q) {
#pragma unroll
            for (int j = 0; j < NB; ++j) pl[q][j] = 0;
            const uint32_t span = tile * TILE_SPANS + q * 32u + lane;
            if (tile < n_tiles && span < a.n_spans) {
                if (NB == 2) {
                    const uint2 m = __ldg(reinterpret_cast<const uint2 *>(a.mask) + span);
                    pl[q][0] = m.x; pl[q][1] = m.y;
                } else {
                    const uint4 m = __ldg(reinterpret_cast<const uint4 *>(a.mask) + span);
                    pl[q][0] = m.x; pl[q][1] = m.y; pl[q][NB - 2] = m.z; pl[q][NB - 1] = m.w;
                }
            }
        }
    };

    uint32_t tile = blockIdx.x * (VERIFY_THREADS / 32) + (tid >> 5);
    uint32_t planes[VT_SUB][NB];
    load_planes(tile, planes);
    for (; tile < n_tiles; tile += n_warps) {
        uint32_t next_planes[VT_SUB][NB];
        load_planes(tile + n_warps, next_planes);      // in flight while this tile is verified

        // ---- flagged words of the four sub-tiles, in stream order
        uint32_t cnt[VT_SUB], incl[VT_SUB];
#pragma unroll
        for (int q = 0; q < VT_SUB; ++q) {
            cnt[q] = 0;
#pragma unroll
            for (int j = 0; j < NB; ++j) cnt[q] += __popc(planes[q][j]);
            incl[q] = cnt[q];
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
            for (int q = 0; q < VT_SUB; ++q) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl[q], d);
                if (lane >= d) incl[q] += v;
            }
        }
        uint32_t base[VT_SUB], n_cand = 0;
        bool dense = false;
#pragma unroll
        for (int q = 0; q < VT_SUB; ++q) {
            const uint32_t n_q = __shfl_sync(0xffffffffu, incl[q], 31);
            base[q] = n_cand;
            n_cand += n_q;
            dense = dense || n_q > a.dense_max;        // cheaper to walk everything
        }

        if (n_cand) {                                   // warp-uniform
            if (dense) ++dense_tiles;
            else {
#pragma unroll
                for (int q = 0; q < VT_SUB; ++q) {
                    if (cnt[q]) {
                        const uint32_t span = tile * TILE_SPANS + q * 32u + lane;
                        uint32_t at = base[q] + incl[q] - cnt[q];
                        uint32_t any = 0;
#pragma unroll
                        for (int j = 0; j < NB; ++j) any |= planes[q][j];
                        while (any) {
                            const uint32_t ch = __ffs(any) - 1;
                            any &= any - 1;
#pragma unroll
                            for (int j = 0; j < NB; ++j)
                                if ((planes[q][j] >> ch) & 1u) my_list[at++] = span * WORDS_PER_SPAN + ch * NB + j;
                        }
                    }
                }
                __syncwarp();
            }

            // ---- one lane per item, VT_LOCK batches at a time; every group of batches is one run of events
            const uint32_t n_items = dense ? 32u * VT_SUB : n_cand;    // dense: batch q = sub-tile q, lane = span
#pragma unroll 1
            for (uint32_t g = 0; g * (32u * VT_LOCK) < n_items; ++g) {
                uint32_t item[VT_LOCK];
                ItemEvents res[VT_LOCK];
#pragma unroll
                for (int k = 0; k < VT_LOCK; ++k) res[k] = ItemEvents{0, 0, 0};

                if (dense) {
#pragma unroll 1
                    for (int q = 0; q < VT_LOCK; ++q) {
                        const uint32_t span = tile * TILE_SPANS + (g * VT_LOCK + q) * 32u + lane;
                        const uint32_t it = (span < a.n_spans) ? (ITEM_SPAN | span) : ITEM_NONE;
                        ItemEvents ev{0, 0, 0};
                        if (it != ITEM_NONE)
                            ev = walk_item_slow<E, RANGE, W, false>(a, s_tab_addr, s_cls_addr, it, 0u, 0u, 0u, 0u);
#pragma unroll
                        for (int k = 0; k < VT_LOCK; ++k) if (k == q) { res[k] = ev; item[k] = it; }
                    }
                } else {
                    uint32_t rs[VT_LOCK], w0[VT_LOCK];
                    bool plain[VT_LOCK], slow[VT_LOCK];
                    uint32_t good_rs = 0, good_w0 = 0;
                    bool have_good = false;
#pragma unroll
                    for (int k = 0; k < VT_LOCK; ++k) {
                        const uint32_t idx = (g * VT_LOCK + k) * 32u + lane;
                        item[k] = (idx < n_cand) ? my_list[idx] : ITEM_NONE;
                        plain[k] = false; slow[k] = false; rs[k] = 0; w0[k] = 0;
                        if (item[k] != ITEM_NONE) {
                            rs[k] = (item[k] + 1u) * W;
                            if (rs[k] < a.total) {         // else nothing ends after this word
                                const uint32_t h = find_haystack(a, rs[k]);
                                const uint32_t hb = hay_begin(a, h);
                                w0[k] = (rs[k] >= a.warm && rs[k] - a.warm > hb) ? rs[k] - a.warm : hb;
                                plain[k] = lockstep_ok && rs[k] >= a.warm && rs[k] + W <= a.total &&
                                           hay_end(a, h) >= rs[k] + W && w0[k] < rs[k] &&
                                           ((rs[k] - w0[k]) & (uint32_t)(W - 1)) == 0;
                                slow[k] = !plain[k];
                                if (plain[k]) { good_rs = rs[k]; good_w0 = w0[k]; have_good = true; }
                            }
                        }
                    }
                    if (__any_sync(0xffffffffu, have_good)) {
                        if (!have_good) { good_rs = a.warm; good_w0 = 0; }     // a harmless walk at the stream start
#pragma unroll
                        for (int k = 0; k < VT_LOCK; ++k)
                            if (!plain[k]) { rs[k] = good_rs; w0[k] = good_w0; }
                        ItemEvents ev[VT_LOCK];
                        walk_words_lockstep<W, VT_LOCK>(st, a.text, a.warm, rs, w0, ev);
#pragma unroll
                        for (int k = 0; k < VT_LOCK; ++k) if (plain[k]) res[k] = ev[k];
                    }
                    // windows clipped by the stream ends or starting at an unaligned haystack start: out of line
#pragma unroll 1
                    for (int k = 0; k < VT_LOCK; ++k) {
                        bool sl = false; uint32_t it = 0;
#pragma unroll
                        for (int kk = 0; kk < VT_LOCK; ++kk) if (kk == k) { sl = slow[kk]; it = item[kk]; }
                        if (sl) {
                            const uint32_t rs1 = (it + 1u) * W;
                            const uint32_t ws1 = (rs1 > a.warm) ? rs1 - a.warm : 0u;
                            const ItemEvents ev1 = walk_item_slow<E, RANGE, W, false>(a, s_tab_addr, s_cls_addr, it, ws1, rs1,
                                                                                      min(rs1 + W, a.total), 0u);
#pragma unroll
                            for (int kk = 0; kk < VT_LOCK; ++kk) if (kk == k) res[kk] = ev1;
                        }
                    }
                }
                __syncwarp();

                // ---- the group's run: offsets inside it, its place in the event buffer
                uint32_t ri[VT_LOCK];
#pragma unroll
                for (int k = 0; k < VT_LOCK; ++k) ri[k] = res[k].cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
                    for (int k = 0; k < VT_LOCK; ++k) {
                        const uint32_t v = __shfl_up_sync(0xffffffffu, ri[k], d);
                        if (lane >= d) ri[k] += v;
                    }
                }
                uint32_t total = 0, boff[VT_LOCK];
#pragma unroll
                for (int k = 0; k < VT_LOCK; ++k) {
                    boff[k] = total;
                    total += __shfl_sync(0xffffffffu, ri[k], 31);
                }
                if (total) {                            // warp-uniform
                    uint32_t run_base = 0;
                    if (lane == 0) {
                        run_base = atomicAdd(&a.counters[1], total);
                        a.runs[tile * (VT_BATCHES / VT_LOCK) + g] = make_uint2(run_base, total);
                    }
                    run_base = __shfl_sync(0xffffffffu, run_base, 0);
#pragma unroll
                    for (int k = 0; k < VT_LOCK; ++k) {
                        if (res[k].cnt) {
                            const uint32_t off = run_base + boff[k] + ri[k] - res[k].cnt;
                            if (res[k].cnt == 1) {
                                if (off < a.capacity) a.out[off] = make_uint2(res[k].e0p, res[k].e0s);
                            } else if (off < a.capacity) {
                                uint32_t ws1 = 0, rs1 = 0, re1 = 0;
                                if (!(item[k] & ITEM_SPAN)) {
                                    rs1 = (item[k] + 1u) * W;
                                    re1 = min(rs1 + W, a.total);
                                    ws1 = (rs1 > a.warm) ? rs1 - a.warm : 0u;
                                }
                                walk_item_slow<E, RANGE, W, true>(a, s_tab_addr, s_cls_addr, item[k], ws1, rs1, re1, off);
                            }
                        }
                    }
                }
            }
            __syncwarp();                               // the list is rewritten for the next tile
        }
#pragma unroll
        for (int q = 0; q < VT_SUB; ++q)
#pragma unroll
            for (int j = 0; j < NB; ++j) planes[q][j] = next_planes[q][j];
    }
    if (lane == 0 && dense_tiles) atomicAdd(&a.counters[4], dense_tiles);

    // state at the end of the stream (keep=1 continuation): the last Lmax bytes decide it
    if (a.want_end_state && blockIdx.x == gridDim.x - 1 && tid == VERIFY_THREADS - 1) {
        const uint32_t back = a.halo + 1u;
        const uint32_t ws = (a.total > back) ? ((a.total - back) & ~(uint32_t)(W - 1)) : 0u;
        a.counters[2] = walk_item_slow<E, RANGE, W, false>(a, s_tab_addr, s_cls_addr, ITEM_NONE, ws, 0xffffffffu, a.total, 0u).e0s;
    }
}

// ------------------------------------------------------------ reorder -----

// One CTA per 1024 tiles, one thread per tile.  A run's final offset is the number of events of all
// earlier tiles: the CTA sums the earlier blocks' run lengths (a coalesced read of at most n_tiles words
// from L2), scans its own 1024 lengths, then its warps copy their 32 runs cooperatively.
__global__ void __launch_bounds__(RUNSCAN_THREADS) ac_reorder_kernel(const uint2 *__restrict__ runs, uint32_t n_tiles,
                                                                      const uint2 *__restrict__ tmp,
                                                                      uint2 *__restrict__ out, uint32_t capacity)
{
    __shared__ uint32_t s_warp[RUNSCAN_THREADS / 32];
    __shared__ uint32_t s_prev[RUNSCAN_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t first = blockIdx.x * RUNSCAN_THREADS;

    uint32_t part = 0;
    for (uint32_t j = tid; j < first; j += RUNSCAN_THREADS) part += runs[j].y;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if (lane == 0) s_prev[warp] = part;

    const uint32_t tile = first + tid;
    const uint2 run = (tile < n_tiles) ? runs[tile] : make_uint2(0u, 0u);
    uint32_t incl = run.y;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t base = 0;
#pragma unroll
    for (int w = 0; w < RUNSCAN_THREADS / 32; ++w) {
        base += s_prev[w];
        if ((uint32_t)w < warp) base += s_warp[w];
    }
    const uint32_t dst0 = base + incl - run.y;

    const uint32_t busy = __ballot_sync(0xffffffffu, run.y != 0);
    for (uint32_t m = busy; m; m &= m - 1) {
        const int src_lane = __ffs(m) - 1;
        const uint32_t src = __shfl_sync(0xffffffffu, run.x, src_lane);
        const uint32_t cnt = __shfl_sync(0xffffffffu, run.y, src_lane);
        const uint32_t dst = __shfl_sync(0xffffffffu, dst0, src_lane);
        for (uint32_t i = lane; i < cnt; i += 32u) {
            const unsigned long long s_i = (unsigned long long)src + i, d_i = (unsigned long long)dst + i;
            if (s_i < capacity && d_i < capacity) out[d_i] = tmp[s_i];
        }
    }
}

} // namespace acb200
