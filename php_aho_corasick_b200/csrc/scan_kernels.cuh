// scan_kernels.cuh — sm_100a kernels of the matcher.
//
//  * expand_*      : ahocorasick_finalize() on the device — turns the uploaded
//                    trie edges + failure links into the dense transition table
//                    delta[state][class], level by level (every failure
//                    transition pre-resolved, so the scan does exactly one
//                    lookup per haystack byte).
//  * ac_scan_kernel: ahocorasick_match() — the replacement of the reference's
//                    hot loop (src/multifast/ahocorasick.c:199-234 with
//                    src/multifast/node.c:119-140 inlined into the table).
//
// The haystack batch is one flat byte stream in HBM.  It is cut into fixed-size
// slices ("chunks"), one per thread; a thread warms its state up over the
// (Lmax-1) bytes before its slice (clamped to the haystack start) and reports
// the events that END inside its slice.  After Lmax-1 bytes the state reached
// from the root equals the state of an uninterrupted scan (every trie node is
// at most Lmax deep), so event lists are identical to a sequential walk.
//
// Events leave the kernel already in ascending buffer order: threads count
// their events, the CTA prefix-sums the counts, and CTAs chain their totals
// through a decoupled look-back over `tile_status` (tiles are handed out by an
// atomic ticket, so a tile only ever waits for tiles that already started).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace acb200 {

constexpr int SCAN_THREADS = 1024;       // one persistent CTA per SM
constexpr int EXPAND_THREADS = 256;

struct ScanArgs {
    const uint8_t *text;          // flat haystack bytes, 16-byte aligned, >=32 readable bytes past `total`
    const uint32_t *hay_off;      // n_hay+1 ascending offsets into text (ragged batches), or nullptr
    uint32_t n_hay;
    uint32_t uniform_len;         // >0: every haystack is exactly this long (hay_off unused)
    uint32_t total;               // bytes in the stream
    uint32_t chunk;               // bytes per thread slice, multiple of 16
    uint32_t halo;                // Lmax-1
    uint32_t chunk_begin;         // this launch covers slices [chunk_begin, chunk_end)
    uint32_t chunk_end;
    uint32_t n_tiles;             // ceil((chunk_end-chunk_begin)/SCAN_THREADS)
    const void *table;            // dense delta, n_states x ncls entries
    const uint8_t *cls_map;       // 256-byte byte->class map
    uint32_t ncls;
    uint32_t first_final;
    uint32_t smem_entries;        // leading table entries cached in shared memory
    uint32_t range_lo;            // RANGE kernels: class = min(byte - range_lo, n_used)
    uint32_t n_used;
    uint32_t init_state;          // state at offset 0 of haystack 0 (keep=1 continuation)
    uint2 *out;                   // events {end offset in stream, state}
    uint32_t capacity;            // events that fit in `out`
    unsigned long long *tile_status;  // n_tiles words, zeroed before launch
    uint32_t *counters;           // [0] ticket (zeroed per launch) [1] running event total [2] end state
    uint32_t *first_end;          // FIRST kernels: per haystack earliest event end seen (init 0xffffffff)
};

// ------------------------------------------------------------ finalize ----

// delta[s][*] = delta[fail(s)][*] for every state s of one breadth-first level
// (root: all zero).  fail(s) is shallower, so its row is already complete.
template <typename E>
__global__ void expand_inherit_kernel(E *__restrict__ table, const uint32_t *__restrict__ order,
                                      const uint32_t *__restrict__ fail, uint32_t lvl_begin,
                                      uint32_t lvl_end, uint32_t ncls)
{
    const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long n = (unsigned long long)(lvl_end - lvl_begin) * ncls;
    if (idx >= n) return;
    const uint32_t si = (uint32_t)(idx / ncls);
    const uint32_t c = (uint32_t)(idx - (unsigned long long)si * ncls);
    const uint32_t s = order[lvl_begin + si];
    E v = 0;
    if (s != 0) v = table[(size_t)fail[s] * ncls + c];
    table[(size_t)s * ncls + c] = v;
}

// then the level's own trie edges overwrite the inherited entries
template <typename E>
__global__ void expand_edges_kernel(E *__restrict__ table, const uint32_t *__restrict__ src,
                                    const uint32_t *__restrict__ dst, const uint16_t *__restrict__ cls,
                                    uint32_t e_begin, uint32_t e_end, uint32_t ncls)
{
    const uint32_t e = e_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= e_end) return;
    table[(size_t)src[e] * ncls + cls[e]] = (E)dst[e];
}

// ---------------------------------------------------------------- scan ----

__device__ __forceinline__ uint4 ld_text16(const uint8_t *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

constexpr unsigned long long ST_AGG = 1ull << 62;     // tile total published
constexpr unsigned long long ST_PREFIX = 2ull << 62;  // inclusive prefix published
constexpr unsigned long long ST_MASK = (1ull << 62) - 1;

template <typename E, bool RANGE, bool FIRST>
struct Scanner {
    const E *__restrict__ gtab;
    const E *s_tab;
    const uint8_t *s_cls;
    const uint8_t *__restrict__ text;
    uint32_t ncls, smem_entries, lo, n_used, first_final;

    // per-thread event record
    uint32_t cnt;
    uint32_t e0p, e0s, e1p, e1s;
    // EMIT mode
    uint2 *out;
    uint32_t obase, cap;
    bool found;   // FIRST: an event was taken in the current haystack segment

    __device__ __forceinline__ uint32_t next(uint32_t s, uint32_t b) const
    {
        uint32_t c;
        if (RANGE) c = min(b - lo, n_used);
        else c = s_cls[b];
        const uint32_t idx = s * ncls + c;
        E e;
        if (idx < smem_entries) e = s_tab[idx];
        else e = __ldg(gtab + idx);
        return (uint32_t)e;
    }

    template <bool EMIT>
    __device__ __forceinline__ void hit(uint32_t pos, uint32_t s)
    {
        if (FIRST) {
            if (found) return;
            found = true;
        }
        if (EMIT) {
            const uint32_t o = obase + cnt;
            if (o < cap) out[o] = make_uint2(pos, s);
        } else {
            if (cnt == 0) { e0p = pos; e0s = s; }
            else if (cnt == 1) { e1p = pos; e1s = s; }
        }
        ++cnt;
    }

    // walk bytes [i, end) without reporting (warm-up over the halo)
    __device__ __forceinline__ uint32_t walk_quiet(uint32_t s, uint32_t i, uint32_t end) const
    {
        while (i < end && (i & 15u)) { s = next(s, text[i]); ++i; }
        while (i + 16 <= end) {
            const uint4 v = ld_text16(text + i);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int k = 0; k < 4; ++k) s = next(s, (w[q] >> (8 * k)) & 0xffu);
            i += 16;
        }
        while (i < end) { s = next(s, text[i]); ++i; }
        return s;
    }

    // walk bytes [i, end) of one haystack, reporting every state >= first_final
    template <bool EMIT>
    __device__ __forceinline__ uint32_t walk_report(uint32_t s, uint32_t i, uint32_t end)
    {
        while (i < end && (i & 15u)) {
            s = next(s, text[i]); ++i;
            if (s >= first_final) hit<EMIT>(i, s);
        }
        if (i + 16 <= end) {
            uint4 v = ld_text16(text + i);
            while (true) {
                // prefetch the following 16 bytes (the buffer is padded past `total`)
                const uint4 nv = ld_text16(text + i + 16);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        s = next(s, (w[q] >> (8 * k)) & 0xffu);
                        if (s >= first_final) hit<EMIT>(i + q * 4 + k + 1, s);
                    }
                i += 16;
                if (FIRST && found) return s;
                if (i + 16 > end) break;
                v = nv;
            }
        }
        while (i < end) {
            s = next(s, text[i]); ++i;
            if (s >= first_final) hit<EMIT>(i, s);
        }
        return s;
    }
};

// index of the haystack that contains stream offset `pos` (pos < total)
__device__ __forceinline__ uint32_t find_haystack(const ScanArgs &a, uint32_t pos)
{
    if (a.uniform_len) return pos / a.uniform_len;
    uint32_t lo = 0, hi = a.n_hay;          // invariant: off[lo] <= pos < off[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(a.hay_off + mid) <= pos) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ uint32_t hay_begin(const ScanArgs &a, uint32_t h)
{
    return a.uniform_len ? h * a.uniform_len : __ldg(a.hay_off + h);
}

__device__ __forceinline__ uint32_t hay_end(const ScanArgs &a, uint32_t h)
{
    return a.uniform_len ? (h + 1) * a.uniform_len : __ldg(a.hay_off + h + 1);
}

// Scans slice [cs, ce).  `s` must be the state at cs, `h` the haystack at cs.
template <bool EMIT, typename SC>
__device__ __forceinline__ uint32_t scan_slice(const ScanArgs &a, SC &sc, uint32_t s, uint32_t h,
                                               uint32_t cs, uint32_t ce)
{
    uint32_t i = cs;
    uint32_t nb = hay_end(a, h);
    sc.found = false;
    while (i < ce) {
        if (i == nb) {                       // haystack boundary: next haystack starts at the root
            do { ++h; nb = hay_end(a, h); } while (nb == i);   // skips empty haystacks
            s = 0;
            sc.found = false;
        }
        const uint32_t seg_end = min(ce, nb);
        s = sc.template walk_report<EMIT>(s, i, seg_end);
        i = seg_end;                         // FIRST kernels may have stopped early; the rest is irrelevant
    }
    return s;
}

template <typename E, bool RANGE, bool FIRST>
__global__ void __launch_bounds__(SCAN_THREADS, 1) ac_scan_kernel(const ScanArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    E *s_tab = reinterpret_cast<E *>(smem_raw);
    __shared__ uint8_t s_cls[256];
    __shared__ uint32_t s_warp_tot[SCAN_THREADS / 32];
    __shared__ uint32_t s_tile, s_base, s_total;

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u, warp = tid >> 5;
    const E *gtab = static_cast<const E *>(a.table);

    for (uint32_t idx = tid; idx < a.smem_entries; idx += SCAN_THREADS) s_tab[idx] = gtab[idx];
    if (tid < 256) s_cls[tid] = a.cls_map[tid];
    __syncthreads();

    Scanner<E, RANGE, FIRST> sc;
    sc.gtab = gtab; sc.s_tab = s_tab; sc.s_cls = s_cls; sc.text = a.text;
    sc.ncls = a.ncls; sc.smem_entries = a.smem_entries; sc.lo = a.range_lo; sc.n_used = a.n_used;
    sc.first_final = a.first_final;
    sc.out = a.out; sc.cap = a.capacity;

    const uint32_t prior = a.counters[1];    // events of earlier launches in this call (stream-ordered)

    while (true) {
        if (tid == 0) s_tile = atomicAdd(&a.counters[0], 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= a.n_tiles) break;

        const uint32_t chunk_id = a.chunk_begin + tile * SCAN_THREADS + tid;
        const bool active = chunk_id < a.chunk_end;
        uint32_t cs = 0, ce = 0, h = 0, s_cs = 0;
        sc.cnt = 0; sc.found = false;
        if (active) {
            cs = chunk_id * a.chunk;
            ce = min(cs + a.chunk, a.total);
            h = find_haystack(a, cs);
            const uint32_t hb = hay_begin(a, h);
            bool skip = false;
            if (FIRST) {
                // whole slice inside one haystack that already has an earlier event: nothing to add
                if (hay_end(a, h) >= ce && a.first_end[h] <= cs) skip = true;
            }
            if (!skip) {
                // warm-up start: (Lmax-1) bytes back, rounded down to 16, clamped to the haystack start
                uint32_t ws = (cs - hb > a.halo) ? ((cs - a.halo) & ~15u) : hb;
                if (ws < hb) ws = hb;
                uint32_t s = (ws == hb && h == 0) ? a.init_state : 0u;
                s = sc.walk_quiet(s, ws, cs);
                s_cs = s;
                s = scan_slice<false>(a, sc, s, h, cs, ce);
                if (ce == a.total) a.counters[2] = s;
                if (FIRST && sc.cnt) {
                    // publish the earliest event of the slice's first reporting haystack
                    const uint32_t hh = find_haystack(a, sc.e0p - 1);
                    atomicMin(&a.first_end[hh], sc.e0p);
                }
            }
        }

        // CTA exclusive prefix of the per-thread event counts
        uint32_t incl = sc.cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t wt = s_warp_tot[lane];
            uint32_t winc = wt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, winc, d);
                if (lane >= d) winc += t;
            }
            s_warp_tot[lane] = winc - wt;                 // exclusive warp offsets
            const uint32_t total = __shfl_sync(0xffffffffu, winc, 31);

            // decoupled look-back over earlier tiles
            unsigned long long excl = prior;
            if (tile == 0) {
                if (lane == 0) st_status(a.tile_status, ST_PREFIX | (excl + total));
            } else {
                if (lane == 0) st_status(a.tile_status + tile, ST_AGG | total);
                long long j = (long long)tile - 1 - lane;
                unsigned long long sum = 0;
                while (true) {
                    unsigned long long v = ST_PREFIX;     // before tile 0: empty prefix (prior added below)
                    bool virt = j < 0;
                    if (!virt) {
                        do { v = ld_status(a.tile_status + j); } while ((v >> 62) == 0);
                    }
                    const bool is_prefix = (v >> 62) == 2;
                    const unsigned pm = __ballot_sync(0xffffffffu, is_prefix);
                    const int first_p = pm ? (__ffs(pm) - 1) : 32;
                    unsigned long long contrib = ((int)lane <= first_p) ? (v & ST_MASK) : 0ull;
                    if (virt && (int)lane == first_p) contrib = prior;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
                    sum += contrib;
                    if (pm) break;
                    j -= 32;
                }
                excl = sum;
                if (lane == 0) st_status(a.tile_status + tile, ST_PREFIX | (excl + total));
            }
            if (lane == 0) {
                s_base = (uint32_t)excl;
                s_total = total;
                if (tile == a.n_tiles - 1) a.counters[1] = (uint32_t)(excl + total);
            }
        }
        __syncthreads();

        if (sc.cnt) {
            const uint32_t off = s_base + s_warp_tot[warp] + (incl - sc.cnt);
            if (sc.cnt <= 2) {
                if (off < a.capacity) a.out[off] = make_uint2(sc.e0p, sc.e0s);
                if (sc.cnt == 2 && off + 1 < a.capacity) a.out[off + 1] = make_uint2(sc.e1p, sc.e1s);
            } else if (off < a.capacity) {
                // dense slice: walk it again from the saved entry state and write in place
                sc.obase = off;
                sc.cnt = 0;
                scan_slice<true>(a, sc, s_cs, h, cs, ce);
            }
        }
        __syncthreads();
    }
}

} // namespace acb200
