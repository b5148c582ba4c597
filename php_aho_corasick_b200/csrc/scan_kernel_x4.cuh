// scan_kernel_x4.cuh — ac_scan_kernel_x4: the scan with FOUR slices per lane walked in lockstep.
//
// Why: one dependent shared-memory lookup chain per thread cannot fill the shared-memory pipe
// (a bank-conflicted random LDS takes ~38 cycles x conflict degree; scripts/micro/lds_chain.cu
// measures 3.76 lookups/cycle/SM with one chain per thread at 1,024 threads and 9.05 with four
// chains per thread at 512 threads — the pipe's conflict-wavefront limit).  So each lane owns four
// independent slices and issues their lookups back to back; the four chains overlap their latency.
//
// The lockstep path covers the regular case — a slice that lies inside one haystack and has the
// full halo — which is every slice of a uniform batch.  Irregular slices (haystack boundary inside,
// clamped halo, stream tail) and steps that leave the shared-memory window fall back to the
// single-slice walker of scan_kernels.cuh, so results are identical by construction.
//
// Tile = 128 consecutive slices per warp: lane l, chain q walks slice  tile*128 + q*32 + l.
#pragma once

#include "scan_kernels.cuh"

namespace acb200 {

constexpr int X4_THREADS = 512;
constexpr int X4 = 4;

template <typename E, bool RANGE>
struct Lockstep {
    using SC = Scanner<E, RANGE, false>;
    SC &sc;
    uint32_t cnt[X4], e0p[X4], e0s[X4], e1p[X4], e1s[X4];

    __device__ __forceinline__ explicit Lockstep(SC &s) : sc(s) {}

    template <int Q>
    __device__ __forceinline__ void hit(uint32_t pos, uint32_t st)
    {
        if (cnt[Q] == 0) { e0p[Q] = pos; e0s[Q] = st; }
        else if (cnt[Q] == 1) { e1p[Q] = pos; e1s[Q] = st; }
        ++cnt[Q];
    }

    // bytes [j0, 16) of one group for one chain on the generic path (window when inside, full table otherwise)
    template <bool REPORT, int Q>
    __device__ __forceinline__ uint32_t finish_group(uint32_t s, const uint4 &v, uint32_t i, int j0)
    {
        for (int j = j0; j < 16; ++j) {
            const uint32_t b = SC::group_byte(v, j);
            uint32_t e = 0;
            if (s - sc.win_lo < sc.win_rows) e = sc.hot_next(s, b);
            if (e == 0) e = sc.any_next(s, b);
            s = e;
            if (REPORT && s < sc.final_bound) hit<Q>(i + j + 1, s);
        }
        return s;
    }

    // One 16-byte group of all four chains.  s[] are true state ids, i[] the stream offsets of the groups.
    template <bool REPORT>
    __device__ __forceinline__ void group(uint32_t (&s)[X4], const uint4 (&v)[X4], const uint32_t (&i)[X4])
    {
        const bool inside = (s[0] - sc.win_lo < sc.win_rows) & (s[1] - sc.win_lo < sc.win_rows) &
                            (s[2] - sc.win_lo < sc.win_rows) & (s[3] - sc.win_lo < sc.win_rows);
        int j0 = 0;
        if (inside) {
            j0 = 16;
#define ACB_X4_STEP(J, W)                                                                           \
            {                                                                                       \
                const uint32_t e0 = sc.hot_next(s[0], __byte_perm(v[0].W, 0, 0x4440 | ((J) & 3)));   \
                const uint32_t e1 = sc.hot_next(s[1], __byte_perm(v[1].W, 0, 0x4440 | ((J) & 3)));   \
                const uint32_t e2 = sc.hot_next(s[2], __byte_perm(v[2].W, 0, 0x4440 | ((J) & 3)));   \
                const uint32_t e3 = sc.hot_next(s[3], __byte_perm(v[3].W, 0, 0x4440 | ((J) & 3)));   \
                if (min(min(e0, e1), min(e2, e3)) < sc.final_bound) {                               \
                    if (e0 == 0 || e1 == 0 || e2 == 0 || e3 == 0) { j0 = J; goto left_window; }     \
                    if (REPORT) {                                                                   \
                        if (e0 < sc.final_bound) hit<0>(i[0] + (J) + 1, e0);                        \
                        if (e1 < sc.final_bound) hit<1>(i[1] + (J) + 1, e1);                        \
                        if (e2 < sc.final_bound) hit<2>(i[2] + (J) + 1, e2);                        \
                        if (e3 < sc.final_bound) hit<3>(i[3] + (J) + 1, e3);                        \
                    }                                                                               \
                }                                                                                   \
                s[0] = e0; s[1] = e1; s[2] = e2; s[3] = e3;                                         \
            }
            ACB_X4_STEP(0, x) ACB_X4_STEP(1, x) ACB_X4_STEP(2, x) ACB_X4_STEP(3, x)
            ACB_X4_STEP(4, y) ACB_X4_STEP(5, y) ACB_X4_STEP(6, y) ACB_X4_STEP(7, y)
            ACB_X4_STEP(8, z) ACB_X4_STEP(9, z) ACB_X4_STEP(10, z) ACB_X4_STEP(11, z)
            ACB_X4_STEP(12, w) ACB_X4_STEP(13, w) ACB_X4_STEP(14, w) ACB_X4_STEP(15, w)
#undef ACB_X4_STEP
            return;
        }
    left_window:
        // some chain is (or just stepped) outside the window: finish the group chain by chain
        s[0] = finish_group<REPORT, 0>(s[0], v[0], i[0], j0);
        s[1] = finish_group<REPORT, 1>(s[1], v[1], i[1], j0);
        s[2] = finish_group<REPORT, 2>(s[2], v[2], i[2], j0);
        s[3] = finish_group<REPORT, 3>(s[3], v[3], i[3], j0);
    }

    // n16 groups of all four chains starting at the 16-aligned offsets p[q]
    template <bool REPORT>
    __device__ __forceinline__ void walk(uint32_t (&s)[X4], const uint32_t (&p)[X4], uint32_t n16)
    {
        if (n16 == 0) return;
        uint4 cur[X4], nxt[X4];
        uint32_t i[X4];
#pragma unroll
        for (int q = 0; q < X4; ++q) { i[q] = p[q]; cur[q] = ld_text16(sc.text + i[q]); }
        for (uint32_t g = 0; g < n16; ++g) {
#pragma unroll
            for (int q = 0; q < X4; ++q) {
                nxt[q] = cur[q];
                if (g + 1 < n16) nxt[q] = ld_text16(sc.text + i[q] + 16);
            }
            group<REPORT>(s, cur, i);
#pragma unroll
            for (int q = 0; q < X4; ++q) { i[q] += 16; cur[q] = nxt[q]; }
        }
    }
};

template <typename E, bool RANGE>
__global__ void __launch_bounds__(X4_THREADS, 1) ac_scan_kernel_x4(const ScanArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    E *s_tab = reinterpret_cast<E *>(smem_raw);
    __shared__ uint8_t s_cls[256];

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const E *gtab = static_cast<const E *>(a.table);

    const uint32_t win_entries = a.win_rows * a.ncls;
    const uint32_t win_first = a.win_lo * a.ncls;
    for (uint32_t idx = tid; idx < win_entries; idx += X4_THREADS) {
        uint32_t e = gtab[win_first + idx];
        if (e - a.win_lo >= a.win_rows) e = 0;
        s_tab[idx] = (E)e;
    }
    if (tid < 256) s_cls[tid] = a.cls_map[tid];
    __syncthreads();

    using SC = Scanner<E, RANGE, false>;
    SC sc;
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

    const uint32_t prior = a.counters[1];

    while (true) {
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(&a.counters[0], 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= a.n_tiles) break;

        Lockstep<E, RANGE> ls(sc);
        uint32_t cs[X4], ce[X4], hh[X4], s_cs[X4], st[X4], ws[X4];
        bool active[X4];
        bool regular = true;
#pragma unroll
        for (int q = 0; q < X4; ++q) {
            const uint32_t chunk_id = a.chunk_begin + tile * (32u * X4) + q * 32u + lane;
            active[q] = chunk_id < a.chunk_end;
            ls.cnt[q] = 0; cs[q] = ce[q] = hh[q] = s_cs[q] = st[q] = ws[q] = 0;
            if (active[q]) {
                cs[q] = chunk_id * a.chunk;
                ce[q] = min(cs[q] + a.chunk, a.total);
                hh[q] = find_haystack(a, cs[q]);
                const uint32_t hb = hay_begin(a, hh[q]);
                uint32_t w = (cs[q] - hb > a.halo) ? ((cs[q] - a.halo) & ~15u) : hb;
                if (w < hb) w = hb;
                ws[q] = w;
                st[q] = (w == hb && hh[q] == 0) ? a.init_state : a.root;
                // regular: full-length slice inside one haystack, 16-aligned warm-up of the common length
                if (ce[q] - cs[q] != a.chunk || hay_end(a, hh[q]) < ce[q] || (w & 15u) ||
                    cs[q] - w != cs[0] - ws[0])
                    regular = false;
            } else {
                regular = false;
            }
        }

        if (regular) {
            ls.template walk<false>(st, ws, (cs[0] - ws[0]) >> 4);
#pragma unroll
            for (int q = 0; q < X4; ++q) s_cs[q] = st[q];
            ls.template walk<true>(st, cs, a.chunk >> 4);
#pragma unroll
            for (int q = 0; q < X4; ++q)
                if (ce[q] == a.total) a.counters[2] = st[q];
        } else {
            // irregular tile edge: one chain at a time with the single-slice walker
#pragma unroll
            for (int q = 0; q < X4; ++q) {
                if (!active[q]) continue;
                sc.cnt = 0; sc.found = false;
                uint32_t s = sc.template walk<false, false>(st[q], ws[q], cs[q]);
                s_cs[q] = s;
                s = scan_slice<false>(a, sc, s, hh[q], cs[q], ce[q]);
                if (ce[q] == a.total) a.counters[2] = s;
                ls.cnt[q] = sc.cnt; ls.e0p[q] = sc.e0p; ls.e0s[q] = sc.e0s; ls.e1p[q] = sc.e1p; ls.e1s[q] = sc.e1s;
            }
        }
        __syncwarp();

        // Event order = slice order = (q, lane): four warp prefix sums, chained.
        uint32_t off[X4];
        uint32_t total = 0;
#pragma unroll
        for (int q = 0; q < X4; ++q) {
            uint32_t incl = ls.cnt[q];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            off[q] = total + incl - ls.cnt[q];
            total += __shfl_sync(0xffffffffu, incl, 31);
        }

        const unsigned long long excl = tile_lookback(a, tile, total, prior, lane);

#pragma unroll
        for (int q = 0; q < X4; ++q) {
            if (!ls.cnt[q]) continue;
            const uint32_t o = (uint32_t)excl + off[q];
            if (ls.cnt[q] <= 2) {
                if (o < a.capacity) a.out[o] = make_uint2(ls.e0p[q], ls.e0s[q]);
                if (ls.cnt[q] == 2 && o + 1 < a.capacity) a.out[o + 1] = make_uint2(ls.e1p[q], ls.e1s[q]);
            } else if (o < a.capacity) {
                sc.obase = o;
                sc.cnt = 0; sc.found = false;
                scan_slice<true>(a, sc, s_cs[q], hh[q], cs[q], ce[q]);
            }
        }
        __syncwarp();
    }
}

} // namespace acb200
