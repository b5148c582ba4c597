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
//   ac_verify_kernel  a CTA takes a 512 KiB chunk of the stream at a time.  Its
//                     warps turn the chunk's bit planes into a dense, ordered
//                     list of work items in shared memory (ballot / popc /
//                     prefix sums): one item per flagged word, or — for a
//                     16 KiB tile with so many flagged words that walking all
//                     of it is cheaper — one item per 512-byte span of the
//                     tile.  Then one lane per item: for a flagged word the
//                     lane walks the automaton from the root over the
//                     (Lmax-1)-byte warm-up plus the W bytes after the word —
//                     exactly the halo argument of ac_scan_kernel, so states
//                     and events are those of an uninterrupted walk; a span
//                     item is walked like an ac_scan_kernel slice.  The
//                     chunk's events are written as one ordered run at an
//                     offset taken from a global counter (no CTA ever waits
//                     for another one).
//   ac_reorder_kernel copies the runs into chunk order: the final event list
//                     is ascending, as the callback contract requires.  The
//                     worst case costs what the plain scan costs.
//
// Replaces the same reference loop as ac_scan_kernel
// (src/multifast/ahocorasick.c:199-234); events are bit-identical.
#pragma once

#include "scan_kernels.cuh"
#include "filter_hash.hpp"

namespace acb200 {

constexpr uint32_t SPAN_BYTES = 512;       // one warp-wide 16-byte load; one verify lane
constexpr int FILTER_UNROLL = 4;           // 16-byte loads in flight per thread
constexpr uint32_t VER_DENSE_MAX = 128;    // flagged words per 16 KiB tile beyond which the whole tile is walked
constexpr uint32_t ITEM_SPAN = 0x80000000u;// work item: walk 512-byte span (item & ~ITEM_SPAN) completely
constexpr uint32_t ITEM_NONE = 0xffffffffu;
constexpr uint32_t CHUNK_SPANS = 32u * (SCAN_THREADS / 32);   // spans per CTA chunk: one 32-span tile per warp (512 KiB)
constexpr int VER_ROUNDS = VER_DENSE_MAX / 32;                // batches of 32 items a warp may get per chunk
constexpr uint32_t VER_LIST_CAP = (SCAN_THREADS / 32) * VER_DENSE_MAX;   // items per chunk
constexpr uint32_t VER_FIXED_SMEM = VER_LIST_CAP * 4u + 2048u;           // item list + scan scratch
constexpr int REORDER_THREADS = 256;

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

    const uint32_t n16 = a.total >> 4;                        // complete 16-byte chunks
    const uint32_t tail_chunk = (a.total & 15u) ? n16 : 0xffffffffu;
    const uint32_t n_warps = gridDim.x * (SCAN_THREADS / 32);
    const uint32_t warp = blockIdx.x * (SCAN_THREADS / 32) + (tid >> 5);
    uint32_t flagged = 0;

    for (uint32_t g0 = warp; g0 < a.n_spans; g0 += n_warps * FILTER_UNROLL) {
        uint4 v[FILTER_UNROLL];
#pragma unroll
        for (int u = 0; u < FILTER_UNROLL; ++u) {
            const uint32_t g = g0 + u * n_warps;
            const uint32_t c = g * 32u + lane;
            v[u] = make_uint4(0, 0, 0, 0);
            if (g < a.n_spans && c < n16) v[u] = ld_text16(a.text + (size_t)c * 16u);
        }
#pragma unroll
        for (int u = 0; u < FILTER_UNROLL; ++u) {
            const uint32_t g = g0 + u * n_warps;
            if (g >= a.n_spans) break;                        // warp-uniform
            const uint32_t c = g * 32u + lane;
            const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
            uint32_t planes[NB];
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                const uint32_t lo = (W == 8) ? w[2 * j] : w[j];
                const uint32_t hi = (W == 8) ? w[2 * j + 1] : 0u;
                const uint32_t idx = filter_reduce(filter_mix1(lo, hi), a.l1_bits);
                uint32_t word;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(word) : "r"(s_base + ((idx >> 5) << 2)));
                bool p = (word >> (idx & 31u)) & 1u;
                {   // second probe of the same bitmap, only where the first one hit
                    const uint32_t idx2 = filter_reduce(filter_mix2(lo, hi), a.l1_bits);
                    uint32_t word2 = 0;
                    if (p) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(word2) : "r"(s_base + ((idx2 >> 5) << 2)));
                    p = (word2 >> (idx2 & 31u)) & 1u;
                }
                if (L2) {
                    uint32_t word3 = 0;
                    const uint32_t i3 = filter_mix3(lo, hi) >> a.l2_shift;
                    if (p) word3 = __ldg(a.l2 + (i3 >> 5));
                    p = (word3 >> (i3 & 31u)) & 1u;
                }
                // the partial chunk at the very end is not read: its words are simply handed on
                p = (c < n16) ? p : (c == tail_chunk);
                planes[j] = __ballot_sync(0xffffffffu, p);
            }
            uint32_t mine = planes[0];
#pragma unroll
            for (int j = 1; j < NB; ++j) if (lane == (uint32_t)j) mine = planes[j];
            if (lane < (uint32_t)NB) {
                a.mask[(size_t)g * NB + lane] = mine;
                flagged += __popc(mine);
            }
        }
    }
    if (lane < (uint32_t)NB && flagged) atomicAdd(&a.counters[3], flagged);
}

// ------------------------------------------------------------- verify -----

template <typename SC>
__device__ __forceinline__ uint32_t dfa_step(const SC &sc, uint32_t s, uint32_t b)
{
    if (s - sc.win_lo < sc.win_rows) {
        const uint32_t e = sc.hot_next(s, b);
        if (e) return e;
    }
    return sc.any_next(s, b);
}

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

__device__ __forceinline__ uint32_t pair_byte(const uint2 &v, int j)
{
    return (((j < 4) ? v.x : v.y) >> ((j & 3) * 8)) & 0xffu;
}

// The W end offsets owned by flagged word k (bytes rs .. rs+W-1 with rs = W(k+1)): walk from the root over
// the warm-up [ws, rs) and report the final states reached inside [rs, re).  Fast version: the window lies
// inside one haystack and inside the stream.
template <int W, bool EMIT, typename SC>
__device__ __forceinline__ void walk_word_fast(const ScanArgs &a, SC &sc, uint32_t ws, uint32_t rs)
{
    uint32_t s = a.root;
    uint2 cur = ld_group<W>(a.text, ws);
    uint2 nxt = (ws + W <= rs) ? ld_group<W>(a.text, ws + W) : cur;
    for (uint32_t i = ws; i < rs; i += W) {
        uint2 nn = nxt;
        if (i + 2u * W <= rs) nn = ld_group<W>(a.text, i + 2u * W);
#pragma unroll
        for (int j = 0; j < W; ++j) s = dfa_step(sc, s, pair_byte(cur, j));
        cur = nxt; nxt = nn;
    }
#pragma unroll
    for (int j = 0; j < W; ++j) {
        s = dfa_step(sc, s, pair_byte(cur, j));
        if (s < a.final_bound) sc.template hit<EMIT>(rs + j + 1u, s);
    }
}

// General version: haystack starts inside the window reset the state, the window may be clipped at the
// end of the stream, nothing is read past `readable`.  rs = 0xffffffff: report nothing (end-state walk).
template <int W, bool EMIT, typename SC>
__device__ __forceinline__ uint32_t walk_word_careful(const ScanArgs &a, SC &sc, uint32_t ws, uint32_t rs, uint32_t re)
{
    uint32_t h = find_haystack(a, ws);
    uint32_t nb = hay_end(a, h);
    uint32_t s = a.root;
    for (uint32_t i = ws; i < re; i += W) {
        const uint2 cur = ld_group_guarded<W>(a, i);
#pragma unroll
        for (int j = 0; j < W; ++j) {
            const uint32_t ii = i + j;
            if (ii < re) {
                if (ii == nb) {                      // a haystack starts here
                    do { ++h; nb = hay_end(a, h); } while (nb == ii);
                    s = a.root;
                }
                s = dfa_step(sc, s, pair_byte(cur, j));
                if (ii >= rs && s < a.final_bound) sc.template hit<EMIT>(ii + 1u, s);
            }
        }
    }
    return s;
}

template <int W, bool EMIT, typename SC>
__device__ __forceinline__ void walk_word(const ScanArgs &a, SC &sc, uint32_t k)
{
    const uint32_t rs = (k + 1u) * W;
    if (rs >= a.total) return;                       // nothing ends after this word
    const uint32_t re = min(rs + W, a.total);
    const uint32_t ws = (rs > a.warm) ? rs - a.warm : 0u;
    bool plain = (re == rs + W);
    if (plain) {
        const uint32_t h = find_haystack(a, ws);
        plain = hay_end(a, h) >= re;
    }
    if (plain) walk_word_fast<W, EMIT>(a, sc, ws, rs);
    else walk_word_careful<W, EMIT>(a, sc, ws, rs, re);
}

template <typename E, bool RANGE, int W>
__global__ void __launch_bounds__(SCAN_THREADS, 1) ac_verify_kernel(const ScanArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint8_t s_cls[256];
    constexpr int NB = 16 / W;
    constexpr uint32_t WORDS_PER_SPAN = 32u * NB;
    constexpr int N_WARPS = SCAN_THREADS / 32;
    constexpr int MAX_BATCHES = N_WARPS * VER_ROUNDS;

    // dynamic shared memory: item list, scan scratch, then the table window
    uint32_t *s_list = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *s_wcnt = s_list + VER_LIST_CAP;             // items per warp tile            [N_WARPS]
    uint32_t *s_btot = s_wcnt + N_WARPS;                  // events per batch -> offsets    [MAX_BATCHES]
    uint32_t *s_misc = s_btot + MAX_BATCHES;              // [0] offset of the chunk's run
    E *s_tab = reinterpret_cast<E *>(smem_raw + VER_FIXED_SMEM);
    static_assert((VER_LIST_CAP + N_WARPS + MAX_BATCHES + 4) * 4 <= VER_FIXED_SMEM, "scan scratch does not fit");

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;
    const E *gtab = static_cast<const E *>(a.table);

    const uint32_t win_entries = a.win_rows * a.ncls;
    const uint32_t win_first = a.win_lo * a.ncls;
    for (uint32_t idx = tid; idx < win_entries; idx += SCAN_THREADS) {
        uint32_t e = gtab[win_first + idx];
        if (e - a.win_lo >= a.win_rows) e = 0;
        s_tab[idx] = (E)e;
    }
    if (tid < 256) s_cls[tid] = a.cls_map[tid];
    __syncthreads();

    Scanner<E, RANGE, false> sc;
    sc.gtab = gtab; sc.text = a.text;
    sc.ncls = a.ncls; sc.row_bytes = a.ncls * (uint32_t)sizeof(E);
    sc.win_lo = a.win_lo; sc.win_rows = a.win_rows;
    {
        const uint32_t t0 = (uint32_t)__cvta_generic_to_shared(s_tab) - a.win_lo * sc.row_bytes;
        const uint32_t c0 = (uint32_t)__cvta_generic_to_shared(s_cls);
        asm volatile("mov.u32 %0, %1;" : "=r"(sc.s_tab) : "r"(t0));
        asm volatile("mov.u32 %0, %1;" : "=r"(sc.s_cls) : "r"(c0));
    }
    sc.lo = a.range_lo; sc.n_used = a.n_used;
    sc.final_bound = a.final_bound; sc.readable = a.readable;
    sc.out = a.out; sc.cap = a.capacity;
    sc.found = false;

    const uint32_t n_chunks = (a.n_spans + CHUNK_SPANS - 1) / CHUNK_SPANS;
    uint32_t dense_tiles = 0;

    for (uint32_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        // ---- 1. flagged words of this warp's 32-span tile
        const uint32_t span = chunk * CHUNK_SPANS + warp * 32u + lane;
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
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const uint32_t n_cand = __shfl_sync(0xffffffffu, incl, 31);
        const bool dense = n_cand > a.dense_max;       // cheaper to walk the whole tile
        const uint32_t n_act = __popc(__ballot_sync(0xffffffffu, active));
        if (lane == 0) {
            s_wcnt[warp] = dense ? n_act : n_cand;
            if (dense) ++dense_tiles;
        }
        __syncthreads();

        // ---- 2. ordered item list of the chunk
        uint32_t n_items;
        {
            const uint32_t v = s_wcnt[lane];
            uint32_t wincl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, wincl, d);
                if (lane >= d) wincl += t;
            }
            n_items = __shfl_sync(0xffffffffu, wincl, 31);
            const uint32_t base = __shfl_sync(0xffffffffu, wincl - v, warp);
            if (dense) {
                if (active) s_list[base + lane] = ITEM_SPAN | span;
            } else if (cnt) {
                uint32_t at = base + incl - cnt;
                uint32_t any = 0;
#pragma unroll
                for (int j = 0; j < NB; ++j) any |= planes[j];
                while (any) {
                    const uint32_t ch = __ffs(any) - 1;
                    any &= any - 1;
#pragma unroll
                    for (int j = 0; j < NB; ++j)
                        if ((planes[j] >> ch) & 1u) s_list[at++] = span * WORDS_PER_SPAN + ch * NB + j;
                }
            }
        }
        __syncthreads();

        // ---- 3. one lane per item: count events (first one kept in registers)
        const uint32_t n_batches = (n_items + 31u) >> 5;
        uint32_t r_item[VER_ROUNDS], r_cnt[VER_ROUNDS], r_excl[VER_ROUNDS], r_e0p[VER_ROUNDS], r_e0s[VER_ROUNDS];
#pragma unroll
        for (int r = 0; r < VER_ROUNDS; ++r) {
            r_item[r] = ITEM_NONE; r_cnt[r] = 0; r_excl[r] = 0; r_e0p[r] = 0; r_e0s[r] = 0;
            const uint32_t b = warp + r * N_WARPS;
            if (b < n_batches) {                       // warp-uniform
                const uint32_t idx = b * 32u + lane;
                const uint32_t item = (idx < n_items) ? s_list[idx] : ITEM_NONE;
                sc.cnt = 0;
                if (item == ITEM_NONE) {
                } else if (item & ITEM_SPAN) {
                    const uint32_t cs = (item & ~ITEM_SPAN) * SPAN_BYTES;
                    const uint32_t ce = min(cs + SPAN_BYTES, a.total);
                    const uint32_t h = find_haystack(a, cs);
                    const uint32_t hb = hay_begin(a, h);
                    uint32_t ws = (cs - hb > a.halo) ? ((cs - a.halo) & ~15u) : hb;
                    if (ws < hb) ws = hb;
                    const uint32_t s_cs = sc.template walk<false, false>(a.root, ws, cs);
                    scan_slice<false>(a, sc, s_cs, h, cs, ce);
                } else {
                    walk_word<W, false>(a, sc, item);
                }
                __syncwarp();
                uint32_t bincl = sc.cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, bincl, d);
                    if (lane >= d) bincl += t;
                }
                if (lane == 31) s_btot[b] = bincl;
                r_item[r] = item; r_cnt[r] = sc.cnt; r_excl[r] = bincl - sc.cnt; r_e0p[r] = sc.e0p; r_e0s[r] = sc.e0s;
            }
        }
        __syncthreads();

        // ---- 4. batch offsets inside the chunk's run; the run's place in the event buffer
        if (warp == 0) {
            constexpr int PER = MAX_BATCHES / 32;
            uint32_t v[PER];
            uint32_t sum = 0;
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const uint32_t b = lane * PER + k;
                v[k] = (b < n_batches) ? s_btot[b] : 0u;
                sum += v[k];
            }
            uint32_t sincl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, sincl, d);
                if (lane >= d) sincl += t;
            }
            uint32_t run = sincl - sum;
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const uint32_t b = lane * PER + k;
                if (b < n_batches) s_btot[b] = run;
                run += v[k];
            }
            if (lane == 31) {
                const uint32_t base = sincl ? atomicAdd(&a.counters[1], sincl) : 0u;
                a.runs[chunk] = make_uint2(base, sincl);
                s_misc[0] = base;
            }
        }
        __syncthreads();

        // ---- 5. emit
        const uint32_t run_base = s_misc[0];
#pragma unroll
        for (int r = 0; r < VER_ROUNDS; ++r) {
            if (r_cnt[r]) {
                const uint32_t b = warp + r * N_WARPS;
                const uint32_t off = run_base + s_btot[b] + r_excl[r];
                if (r_cnt[r] == 1) {
                    if (off < a.capacity) a.out[off] = make_uint2(r_e0p[r], r_e0s[r]);
                } else if (off < a.capacity) {
                    sc.obase = off;
                    sc.cnt = 0;
                    const uint32_t item = r_item[r];
                    if (item & ITEM_SPAN) {
                        const uint32_t cs = (item & ~ITEM_SPAN) * SPAN_BYTES;
                        const uint32_t ce = min(cs + SPAN_BYTES, a.total);
                        const uint32_t h = find_haystack(a, cs);
                        const uint32_t hb = hay_begin(a, h);
                        uint32_t ws = (cs - hb > a.halo) ? ((cs - a.halo) & ~15u) : hb;
                        if (ws < hb) ws = hb;
                        const uint32_t s_cs = sc.template walk<false, false>(a.root, ws, cs);
                        scan_slice<true>(a, sc, s_cs, h, cs, ce);
                    } else {
                        walk_word<W, true>(a, sc, item);
                    }
                }
            }
        }
        // the next chunk's first barrier orders these reads before the scratch is overwritten
    }
    if (lane == 0 && dense_tiles) atomicAdd(&a.counters[4], dense_tiles);

    // state at the end of the stream (keep=1 continuation): the last Lmax bytes decide it
    if (a.want_end_state && blockIdx.x == gridDim.x - 1 && tid == SCAN_THREADS - 1) {
        const uint32_t back = a.halo + 1u;
        const uint32_t ws = (a.total > back) ? ((a.total - back) & ~(uint32_t)(W - 1)) : 0u;
        a.counters[2] = walk_word_careful<W, false>(a, sc, ws, 0xffffffffu, a.total);
    }
}

// ------------------------------------------------------------ reorder -----

// One CTA per chunk: the run's final offset is the number of events of all earlier chunks.
__global__ void __launch_bounds__(REORDER_THREADS) ac_reorder_kernel(const uint2 *__restrict__ runs,
                                                                      const uint2 *__restrict__ tmp,
                                                                      uint2 *__restrict__ out, uint32_t capacity)
{
    __shared__ uint32_t s_part[REORDER_THREADS / 32];
    const uint32_t chunk = blockIdx.x;
    const uint2 run = runs[chunk];
    if (run.y == 0) return;                            // CTA-uniform
    uint32_t part = 0;
    for (uint32_t j = threadIdx.x; j < chunk; j += REORDER_THREADS) part += runs[j].y;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if ((threadIdx.x & 31u) == 0) s_part[threadIdx.x >> 5] = part;
    __syncthreads();
    uint32_t excl = 0;
#pragma unroll
    for (int w = 0; w < REORDER_THREADS / 32; ++w) excl += s_part[w];
    for (uint32_t i = threadIdx.x; i < run.y; i += REORDER_THREADS) {
        const unsigned long long src = (unsigned long long)run.x + i, dst = (unsigned long long)excl + i;
        if (src < capacity && dst < capacity) out[dst] = tmp[src];
    }
}

} // namespace acb200
