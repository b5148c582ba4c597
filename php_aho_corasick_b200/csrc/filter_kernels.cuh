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
//   ac_verify_kernel  reads the bit planes; for every flagged word it walks the
//                     automaton from the root over the (Lmax-1)-byte warm-up
//                     plus the W bytes after the word — exactly the halo
//                     argument of ac_scan_kernel, so the states and events are
//                     those of an uninterrupted walk — and emits the events in
//                     ascending order through the same decoupled look-back.
//                     A 16 KiB tile with too many flagged words is walked
//                     completely, lane per 512-byte span, like ac_scan_kernel:
//                     the worst case costs what the plain scan costs.
//
// Replaces the same reference loop as ac_scan_kernel
// (src/multifast/ahocorasick.c:199-234); events are bit-identical.
#pragma once

#include "scan_kernels.cuh"
#include "filter_hash.hpp"

namespace acb200 {

constexpr uint32_t SPAN_BYTES = 512;       // one warp-wide 16-byte load; one verify lane
constexpr int FILTER_UNROLL = 4;           // 16-byte loads in flight per thread
constexpr int VER_LIST_CAP = 256;          // flagged words per tile the sparse path takes
constexpr int VER_STAGE_CAP = 64;          // events per tile staged in shared memory

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
                if (L2) {
                    uint32_t word2 = 0;
                    const uint32_t i2 = filter_mix2(lo, hi) >> a.l2_shift;
                    if (p) word2 = __ldg(a.l2 + (i2 >> 5));
                    p = (word2 >> (i2 & 31u)) & 1u;
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

__device__ __forceinline__ uint4 ld_text16_guarded(const ScanArgs &a, uint32_t i)
{
    if (i + 16u <= a.readable) return ld_text16(a.text + i);
    uint32_t w[4] = {0, 0, 0, 0};
    for (uint32_t j = 0; j < 16u; ++j)
        if (i + j < a.readable) w[j >> 2] |= (uint32_t)a.text[i + j] << ((j & 3u) * 8u);
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// Walks bytes [ws, re) of the stream from the root (ws is a multiple of 16; haystack starts inside
// the window reset the state) and records the reporting states reached by bytes at index >= rs.
// EMIT: events go to dst[0..); otherwise they are counted and the first one is kept in `first`.
// Returns the number of events; *end_state receives the state after byte re-1.
template <bool EMIT, typename SC>
__device__ __forceinline__ uint32_t walk_window(const ScanArgs &a, const SC &sc, uint32_t ws, uint32_t rs,
                                                uint32_t re, uint2 *dst, uint2 &first, uint32_t *end_state)
{
    uint32_t h = find_haystack(a, ws);
    uint32_t nb = hay_end(a, h);
    uint32_t s = a.root;
    uint32_t n = 0;
    uint4 cur = ld_text16_guarded(a, ws);
    for (uint32_t i = ws; i < re; i += 16u) {
        uint4 nxt = cur;
        if (i + 16u < re) nxt = ld_text16_guarded(a, i + 16u);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint32_t ii = i + j;
            if (ii < re) {
                if (ii == nb) {                      // a haystack starts here
                    do { ++h; nb = hay_end(a, h); } while (nb == ii);
                    s = a.root;
                }
                s = dfa_step(sc, s, SC::group_byte(cur, j));
                if (ii >= rs && s < a.final_bound) {
                    if (EMIT) dst[n] = make_uint2(ii + 1u, s);
                    else if (n == 0) first = make_uint2(ii + 1u, s);
                    ++n;
                }
            }
        }
        cur = nxt;
    }
    if (end_state) *end_state = s;
    return n;
}

template <typename E, bool RANGE, int W>
__global__ void __launch_bounds__(SCAN_THREADS, 1) ac_verify_kernel(const ScanArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint8_t s_cls[256];
    constexpr int NB = 16 / W;
    constexpr uint32_t WORDS_PER_TILE = 32u * 32u * NB;

    // dynamic shared memory: per-warp candidate lists, per-warp event staging, then the table window
    uint16_t *s_list = reinterpret_cast<uint16_t *>(smem_raw);
    uint2 *s_stage = reinterpret_cast<uint2 *>(smem_raw + (SCAN_THREADS / 32) * VER_LIST_CAP * sizeof(uint16_t));
    E *s_tab = reinterpret_cast<E *>(smem_raw + (SCAN_THREADS / 32) * (VER_LIST_CAP * sizeof(uint16_t) +
                                                                       VER_STAGE_CAP * sizeof(uint2)));

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
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

    uint16_t *my_list = s_list + (tid >> 5) * VER_LIST_CAP;
    uint2 *my_stage = s_stage + (tid >> 5) * VER_STAGE_CAP;
    const uint32_t prior = a.counters[1];
    uint32_t dense_tiles = 0;

    while (true) {
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(&a.counters[0], 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= a.n_tiles) break;

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
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const uint32_t n_cand = __shfl_sync(0xffffffffu, incl, 31);

        bool dense = n_cand > (uint32_t)VER_LIST_CAP;
        uint32_t total = 0;                        // events of this tile

        if (!dense && n_cand) {
            // flagged words of the tile in ascending stream order
            uint32_t at = incl - cnt;
            uint32_t any = 0;
#pragma unroll
            for (int j = 0; j < NB; ++j) any |= planes[j];
            while (any) {
                const uint32_t c = __ffs(any) - 1;
                any &= any - 1;
#pragma unroll
                for (int j = 0; j < NB; ++j)
                    if ((planes[j] >> c) & 1u) my_list[at++] = (uint16_t)((lane * 32u + c) * NB + j);
            }
            __syncwarp();

            for (uint32_t b = 0; b < n_cand; b += 32u) {
                const uint32_t idx = b + lane;
                uint32_t n_ev = 0, ws = 0, rs = 0, re = 0;
                uint2 first = make_uint2(0, 0);
                if (idx < n_cand) {
                    const uint32_t k = tile * WORDS_PER_TILE + my_list[idx];
                    rs = (k + 1u) * W;             // the W end offsets owned by word k are rs+1 .. rs+W
                    if (rs < a.total) {
                        re = min(rs + W, a.total);
                        ws = (rs > a.halo) ? ((rs - a.halo) & ~15u) : 0u;
                        n_ev = walk_window<false>(a, sc, ws, rs, re, nullptr, first, nullptr);
                    }
                }
                uint32_t bincl = n_ev;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, bincl, d);
                    if (lane >= d) bincl += t;
                }
                const uint32_t btotal = __shfl_sync(0xffffffffu, bincl, 31);
                if (total + btotal > (uint32_t)VER_STAGE_CAP) { dense = true; break; }   // warp-uniform
                if (n_ev == 1) my_stage[total + bincl - 1u] = first;
                else if (n_ev > 1) walk_window<true>(a, sc, ws, rs, re, my_stage + (total + bincl - n_ev), first, nullptr);
                total += btotal;
            }
            __syncwarp();
        }

        uint32_t cs = 0, ce = 0, h = 0, s_cs = 0;
        sc.cnt = 0;
        if (dense) {
            // too many flagged words (or events): walk the whole tile, one 512-byte span per lane
            ++dense_tiles;
            if (active) {
                cs = span * SPAN_BYTES;
                ce = min(cs + SPAN_BYTES, a.total);
                h = find_haystack(a, cs);
                const uint32_t hb = hay_begin(a, h);
                uint32_t ws = (cs - hb > a.halo) ? ((cs - a.halo) & ~15u) : hb;
                if (ws < hb) ws = hb;
                uint32_t s = sc.template walk<false, false>(a.root, ws, cs);
                s_cs = s;
                s = scan_slice<false>(a, sc, s, h, cs, ce);
                if (ce == a.total) a.counters[2] = s;
            }
            __syncwarp();
            incl = sc.cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            total = __shfl_sync(0xffffffffu, incl, 31);
        } else if (tile == a.n_tiles - 1 && lane == 0) {
            // state at the end of the stream (keep=1 continuation): the last Lmax bytes decide it
            const uint32_t back = a.halo + 1u;
            const uint32_t ws = (a.total > back) ? ((a.total - back) & ~15u) : 0u;
            uint2 dummy;
            uint32_t s_end = a.root;
            walk_window<false>(a, sc, ws, 0xffffffffu, a.total, nullptr, dummy, &s_end);
            a.counters[2] = s_end;
        }

        const unsigned long long excl = tile_lookback(a, tile, total, prior, lane);

        if (!dense) {
            for (uint32_t i = lane; i < total; i += 32u) {
                const unsigned long long o = excl + i;
                if (o < a.capacity) a.out[o] = my_stage[i];
            }
        } else if (sc.cnt) {
            const uint32_t off = (uint32_t)excl + (incl - sc.cnt);
            if (sc.cnt <= 2) {
                if (off < a.capacity) a.out[off] = make_uint2(sc.e0p, sc.e0s);
                if (sc.cnt == 2 && off + 1 < a.capacity) a.out[off + 1] = make_uint2(sc.e1p, sc.e1s);
            } else if (off < a.capacity) {
                sc.obase = off;
                sc.cnt = 0;
                scan_slice<true>(a, sc, s_cs, h, cs, ce);
            }
        }
        __syncwarp();
    }
    if (lane == 0 && dense_tiles) atomicAdd(&a.counters[4], dense_tiles);
}

} // namespace acb200
