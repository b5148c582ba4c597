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
// Events leave the kernel already in ascending buffer order: lanes count their
// events, the warp prefix-sums the counts, and warps chain their totals through
// a decoupled look-back over `tile_status` (tiles of 32 slices are handed out by
// an atomic ticket, so a tile only ever waits for tiles that already started).
// After the table is staged there is no CTA-wide barrier.
#pragma once

#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace acb200 {

constexpr int SCAN_THREADS = 1024;       // one persistent CTA per SM
constexpr int EXPAND_THREADS = 256;
constexpr uint32_t DENSE_PAIRED_MIN = 64; // events in one slice beyond which its second walk pairs them into 16-byte stores

struct ScanArgs {
    const uint8_t *text;          // flat haystack bytes, 16-byte aligned, >=32 readable bytes past `total`
    const uint32_t *hay_off;      // n_hay+1 ascending offsets into text (ragged batches), or nullptr
    uint32_t n_hay;
    uint32_t uniform_len;         // >0: every haystack is exactly this long (hay_off unused)
    uint32_t total;               // bytes in the stream
    uint32_t readable;            // bytes that may be read from `text` (>= total; staging buffers are padded)
    uint32_t chunk;               // bytes per thread slice, multiple of 16
    uint32_t halo;                // Lmax-1
    uint32_t chunk_begin;         // this launch covers slices [chunk_begin, chunk_end)
    uint32_t chunk_end;
    uint32_t n_tiles;             // ceil((chunk_end-chunk_begin)/32): a tile is one warp's 32 slices
    const void *table;            // dense delta, n_rows x ncls entries (row 0 unused)
    const uint8_t *cls_map;       // 256-byte byte->class map
    uint32_t ncls;
    uint32_t final_bound;         // states in [1, final_bound) report patterns
    uint32_t root;                // == final_bound
    uint32_t win_lo;              // states [win_lo, win_lo + win_rows) have their rows in shared memory
    uint32_t win_rows;            // 0: no hot window, every byte takes the careful path
    uint32_t range_lo;            // RANGE kernels: class = min(byte - range_lo, n_used)
    uint32_t n_used;
    uint32_t init_state;          // state at offset 0 of haystack 0 (root, or the keep=1 continuation)
    uint2 *out;                   // events {end offset in stream, state}
    uint32_t capacity;            // events that fit in `out`
    unsigned long long *tile_status;  // n_tiles words, zeroed before launch
    uint32_t *counters;           // [0] ticket (zeroed per launch) [1] running event total [2] end state
                                  // [3] words flagged by the filter [4] tiles handed on for a complete walk [5] work items
    uint32_t *first_end;          // FIRST kernels: per haystack earliest event end seen (init 0xffffffff)
    uint32_t *host_counters;      // pinned host memory that also receives counters[1] and counters[2] (the host reads them
                                  // after its wait: one copy-engine round trip less per call); or nullptr
};

// ------------------------------------------------------------ finalize ----

// delta[s][*] = delta[fail(s)][*] for every state s of one breadth-first level
// (root: every byte leads back to the root).  fail(s) is shallower, so its row
// is already complete.
template <typename E>
__global__ void expand_inherit_kernel(E *__restrict__ table, const uint32_t *__restrict__ order,
                                      const uint32_t *__restrict__ fail, uint32_t lvl_begin,
                                      uint32_t lvl_end, uint32_t ncls, uint32_t root)
{
    const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long n = (unsigned long long)(lvl_end - lvl_begin) * ncls;
    if (idx >= n) return;
    const uint32_t si = (uint32_t)(idx / ncls);
    const uint32_t c = (uint32_t)(idx - (unsigned long long)si * ncls);
    const uint32_t s = order[lvl_begin + si];
    E v = (E)root;
    if (s != root) v = table[(size_t)fail[s] * ncls + c];
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

// Decoupled look-back over earlier tiles, 32 predecessors per step: publishes this
// tile's total in status[tile], returns the sum over all earlier tiles plus `prior`
// and publishes the inclusive prefix.  Tiles must be handed out in ascending order
// (atomic ticket) so that a tile only ever waits for tiles that already started.
// Must be called by all 32 lanes of the warp that owns `tile`.
__device__ __forceinline__ unsigned long long lookback(unsigned long long *status, uint32_t tile, uint32_t total,
                                                       uint32_t prior, uint32_t lane)
{
    unsigned long long excl = prior;
    if (tile == 0) {
        if (lane == 0) st_status(status, ST_PREFIX | (excl + total));
    } else {
        if (lane == 0) st_status(status + tile, ST_AGG | total);
        long long j = (long long)tile - 1 - lane;
        unsigned long long sum = 0;
        while (true) {
            unsigned long long v = ST_PREFIX;         // before tile 0: the events of earlier launches
            const bool virt = j < 0;
            if (!virt) {
                // (a warp that finished early polls until the slowest of its 32 predecessors has published — ncu counted 480
                // polls per tile and lane, 16 % of the kernel's instructions, issued next to the warps that still walk: sleep)
                while (((v = ld_status(status + j)) >> 62) == 0) __nanosleep(256);
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
        if (lane == 0) st_status(status + tile, ST_PREFIX | (excl + total));
    }
    return excl;
}

// The scan kernels' use of it: event offsets; the last tile stores the grand total in counters[1].
__device__ __forceinline__ unsigned long long tile_lookback(const ScanArgs &a, uint32_t tile, uint32_t total,
                                                            uint32_t prior, uint32_t lane)
{
    const unsigned long long excl = lookback(a.tile_status, tile, total, prior, lane);
    if (lane == 0 && tile == a.n_tiles - 1) {
        a.counters[1] = (uint32_t)(excl + total);
        if (a.host_counters) a.host_counters[1] = (uint32_t)(excl + total);
    }
    return excl;
}

// Shared-memory loads by 32-bit shared-window address (keeps the address math to one IMAD).
template <typename E> __device__ __forceinline__ uint32_t lds_entry(uint32_t addr);
template <> __device__ __forceinline__ uint32_t lds_entry<uint16_t>(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
template <> __device__ __forceinline__ uint32_t lds_entry<uint32_t>(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// Stages the window rows of the dense table into shared memory with 16-byte loads, in WINDOW-RELATIVE ids: row 0 of
// the copy is a sink (all zeros), row r + 1 is the row of state win_lo + r, and an entry is its target's row number —
// 0 when the target lies outside the window.  A walk that steps out of the window therefore just stays in the sink,
// reading valid shared memory, and the hot loop needs no exit in the middle of a 16-byte group.
// The copy starts up to 15 bytes into s_base, so that the table's aligned 16-byte groups are aligned in shared memory
// as well: a group is ONE 16-byte load, four of them in flight per thread, and ONE conflict-free 16-byte store.
// (ncu on a 0.25 MiB call, where the kernel is 35 us long: a third of the active cycles went into this copy when every
// load was waited for and its eight entries left as eight 2-byte stores, four-way bank-conflicted each.)
// s_base must hold (win_rows + 1) rows + 16 bytes.  Called by all threads of the CTA (n_threads of them); returns
// the address of the sink row.
template <typename E>
__device__ __forceinline__ E *stage_window(E *s_base, const E *gtab, uint32_t win_lo, uint32_t win_rows, uint32_t ncls,
                                           uint32_t tid, uint32_t n_threads)
{
    constexpr uint32_t PER = 16 / sizeof(E);               // entries per 16-byte load
    auto rel = [&](uint32_t e) -> uint32_t { e -= win_lo; return (e < win_rows) ? e + 1u : 0u; };
    auto rel_group = [&](const uint4 &q) -> uint4 {
        if (sizeof(E) == 2) {
            auto two = [&](uint32_t w) -> uint32_t { return rel(w & 0xffffu) | (rel(w >> 16) << 16); };
            return make_uint4(two(q.x), two(q.y), two(q.z), two(q.w));
        }
        return make_uint4(rel(q.x), rel(q.y), rel(q.z), rel(q.w));
    };
    const uint32_t win_entries = win_rows * ncls;
    const uint32_t win_first = win_lo * ncls;
    const uint32_t lead = min((PER - (win_first % PER)) % PER, win_entries);   // entries before the first aligned group
    E *s_tab = s_base + (PER - (ncls + lead) % PER) % PER;
    for (uint32_t idx = tid; idx < ncls; idx += n_threads) s_tab[idx] = (E)0;          // the sink row
    E *rows = s_tab + ncls;
    for (uint32_t idx = tid; idx < lead; idx += n_threads) rows[idx] = (E)rel(gtab[win_first + idx]);
    const uint32_t n_vec = (win_entries - lead) / PER;
    const uint4 *src = reinterpret_cast<const uint4 *>(gtab + win_first + lead);
    uint4 *dst = reinterpret_cast<uint4 *>(rows + lead);
    uint32_t v = tid;
    for (; v + 3u * n_threads < n_vec; v += 4u * n_threads) {
        uint4 q[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) q[k] = __ldg(src + v + k * n_threads);
#pragma unroll
        for (int k = 0; k < 4; ++k) dst[v + k * n_threads] = rel_group(q[k]);
    }
    for (; v < n_vec; v += n_threads) dst[v] = rel_group(__ldg(src + v));
    for (uint32_t idx = lead + n_vec * PER + tid; idx < win_entries; idx += n_threads) rows[idx] = (E)rel(gtab[win_first + idx]);
    return s_tab;
}

// Per-thread walker.
//
// Shared memory holds the rows of a contiguous window of state ids around
// final_bound: the shallowest final states just below it and the shallowest
// non-final states (root first) just above it — the rows a scan visits almost
// all the time — in window-relative ids (stage_window): 0 is the sink, the
// reporting states of the window are the ids below fin_rel.  ONE compare per
// byte (entry < fin_rel) catches both rare cases: a reporting state (record
// the event, keep walking) and a step out of the window (remember where; the
// rest of the 16-byte group runs on in the sink and is redone afterwards on
// the careful path, which reads true entries from the table in HBM/L2).  The
// common byte costs {PRMT, class, address, IMAD, LDS, ISETP, BRA}.
template <typename E, bool RANGE, bool FIRST>
struct Scanner {
    const E *__restrict__ gtab;
    const uint8_t *__restrict__ text;
    uint32_t s_tab;          // shared-window byte address of the staged window's sink row (row 0)
    uint32_t fin_rel;        // window-relative ids below this are the sink (0) and the window's reporting states
    uint32_t s_cls;          // shared-window byte address of the 256-byte class map
    uint32_t ncls, row_bytes, win_lo, win_rows, lo, n_used, final_bound, readable;

    // per-thread event record
    uint32_t cnt;
    uint32_t e0p, e0s, e1p, e1s;
    // EMIT mode
    uint2 *out;
    uint32_t obase, cap;
    uint32_t report_from;    // count pass: events that end at or before this stream offset belong to the warm-up
    uint2 pend;              // event at an even output index, waiting to be stored together with its successor
    bool have_pend;

    __device__ __forceinline__ void flush_pending()
    {
        if (have_pend) { out[obase + cnt - 1u] = pend; have_pend = false; }
    }

    bool found;   // FIRST: an event was taken in the current haystack segment

    __device__ __forceinline__ uint32_t cls(uint32_t b) const
    {
        if (RANGE) return min(b - lo, n_used);
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(s_cls + b));
        return v;
    }

    // fast step in window-relative ids: s = row number of the state (0 = sink); returns the target's, 0 when it is outside
    __device__ __forceinline__ uint32_t hot_next(uint32_t s, uint32_t b) const
    {
        // t is off the dependent chain; the chain is LDS -> IMAD -> LDS
        const uint32_t t = s_tab + cls(b) * (uint32_t)sizeof(E);
        uint32_t addr;
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(addr) : "r"(s), "r"(row_bytes), "r"(t));
        return lds_entry<E>(addr);
    }

    // careful step: any state, true entry from the full table
    __device__ __forceinline__ uint32_t any_next(uint32_t s, uint32_t b) const
    {
        return (uint32_t)__ldg(gtab + (s * ncls + cls(b)));
    }

    // EMIT: 0 = count (and keep the first two events), 1 = write in place, 2 = write in place, two events per store
    template <int EMIT>
    __device__ __forceinline__ void hit(uint32_t pos, uint32_t s)
    {
        // (the warm-up before a slice runs through the SAME unrolled code as the slice — one copy less of the 16-step
        // switch in the instruction cache, which this kernel stalls on — and what it finds does not count)
        if (EMIT == 0 && pos <= report_from) return;
        if (FIRST) {
            if (found) return;
            found = true;
        }
        if (EMIT == 2) {
            // Very dense slices: every lane writes into its own region, so a warp store touches 32 different sectors
            // and what it costs is store wavefronts, not bytes.  Two events go out as ONE 16-byte store (the event at a
            // 16-byte aligned slot waits in registers for its successor; flush_pending() writes a last odd one).
            const uint32_t o = obase + cnt;
            if (o < cap) {
                if ((reinterpret_cast<uintptr_t>(out + o) & 15u) == 0u) { pend = make_uint2(pos, s); have_pend = true; }
                else if (have_pend) {
                    *reinterpret_cast<uint4 *>(out + (o - 1u)) = make_uint4(pend.x, pend.y, pos, s);
                    have_pend = false;
                } else out[o] = make_uint2(pos, s);
            } else if (have_pend) {                    // the buffer ends between the two: the waiting one goes out alone
                out[o - 1u] = pend;
                have_pend = false;
            }
        } else if (EMIT == 1) {
            const uint32_t o = obase + cnt;
            if (o < cap) out[o] = make_uint2(pos, s);
        } else {
            if (cnt == 0) { e0p = pos; e0s = s; }
            else if (cnt == 1) { e1p = pos; e1s = s; }
        }
        ++cnt;
    }

    template <bool REPORT, int EMIT>
    __device__ __forceinline__ uint32_t byte_step(uint32_t s, uint32_t i)
    {
        s = any_next(s, text[i]);
        if (REPORT && s < final_bound) hit<EMIT>(i + 1, s);
        return s;
    }

    __device__ __forceinline__ static uint32_t group_byte(const uint4 &v, int j)
    {
        const uint32_t word = (j < 8) ? ((j < 4) ? v.x : v.y) : ((j < 12) ? v.z : v.w);
        return (word >> ((j & 3) * 8)) & 0xffu;
    }
    __device__ __forceinline__ static uint32_t group_byte_dyn(const uint4 &v, int j)
    {
        const uint32_t lo = (j & 4) ? v.y : v.x, hi = (j & 4) ? v.w : v.z;
        return (((j & 8) ? hi : lo) >> ((j & 3) * 8)) & 0xffu;
    }

    // One 16-byte group.  Inside the window the 16 steps are straight-line code in window-relative ids with NO branch:
    // every step's entry stays in a register and one predicate collects "some entry was below fin_rel" (a reporting
    // state, or the sink = the walk left the window; the remaining steps then stay in the sink).  Only a group with
    // such an entry looks at its 16 entries again: events are recorded from the registers, and from the first step
    // that left the window the compact careful loop below finishes the group (ONE copy per instantiation).
    template <bool REPORT, int EMIT>
    __device__ __forceinline__ uint32_t walk_group(uint32_t s, const uint4 &v, uint32_t i)
    {
        int j = 0;                                       // first byte of the group the careful loop has to take
        if (s - win_lo < win_rows) {
            const uint32_t q_in = s - win_lo + 1u;       // the state's row in the staged window
            uint32_t e[16];
            bool rare = false;
#define ACB_STEP(J, W, P)                                                                       \
            e[J] = hot_next(P, __byte_perm(W, 0, 0x4440 | ((J) & 3)));                          \
            rare = rare || (e[J] < fin_rel);
            ACB_STEP(0, v.x, q_in) ACB_STEP(1, v.x, e[0]) ACB_STEP(2, v.x, e[1]) ACB_STEP(3, v.x, e[2])
            ACB_STEP(4, v.y, e[3]) ACB_STEP(5, v.y, e[4]) ACB_STEP(6, v.y, e[5]) ACB_STEP(7, v.y, e[6])
            ACB_STEP(8, v.z, e[7]) ACB_STEP(9, v.z, e[8]) ACB_STEP(10, v.z, e[9]) ACB_STEP(11, v.z, e[10])
            ACB_STEP(12, v.w, e[11]) ACB_STEP(13, v.w, e[12]) ACB_STEP(14, v.w, e[13]) ACB_STEP(15, v.w, e[14])
#undef ACB_STEP
            if (__builtin_expect(!rare, 1)) return e[15] + win_lo - 1u;
            j = 16;
            uint32_t prev = q_in;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                if (j == 16 && e[k] < fin_rel) {
                    if (e[k] == 0) { j = k; s = prev + win_lo - 1u; }
                    else if (REPORT) hit<EMIT>(i + k + 1, e[k] + win_lo - 1u);
                }
                prev = e[k];
            }
            if (j == 16) return e[15] + win_lo - 1u;
        }
        // careful, from byte j on: the window's copy where the state is inside it, the true entry from the full table
        // where it is not or the step leaves it (a planted needle walks out of the window and drops back within a few bytes)
        // (two tight loops instead of one that decides per byte: ncu counted 28 % of the kernel's instructions in here on
        // a text with a needle per KiB and lane)
#pragma unroll 1
        while (j < 16) {
            if (s - win_lo < win_rows) {
                uint32_t q = s - win_lo + 1u;            // inside: the window's copy until a step leaves it
#pragma unroll 1
                while (j < 16) {
                    const uint32_t e = hot_next(q, group_byte_dyn(v, j));
                    if (e == 0) break;
                    if (REPORT && e < fin_rel) hit<EMIT>(i + j + 1, e + win_lo - 1u);
                    q = e;
                    ++j;
                }
                s = q + win_lo - 1u;
                if (j == 16) break;
            }
            // outside, or the step that leaves: the true entry from the full table
            s = any_next(s, group_byte_dyn(v, j));
            if (REPORT && s < final_bound) hit<EMIT>(i + j + 1, s);
            ++j;
        }
        return s;
    }

    // Walks bytes [i, end) of one haystack from state s.  REPORT: record every
    // reporting state (EMIT: straight into the output, else count + keep 2).
    // Text arrives through a three-deep register pipeline of 16-byte loads.
    // The second walk of a VERY dense slice (EMIT == 2), which writes its events in place two per store.  Deliberately
    // compact (one byte per iteration, nothing unrolled): this pass is bound by its stores, and a third unrolled
    // copy of the 16-step switch would push the count pass's hot loop out of the instruction cache.  Slices with a
    // handful of events take the unrolled walk (EMIT == 1): one warp in three re-walks a slice at one event per
    // KiB, and the compact loop costs 4x the instructions per byte (measured: +36 % instructions, +18 % time).
    __device__ __forceinline__ uint32_t walk_emit(uint32_t s, uint32_t i, uint32_t end)
    {
#pragma unroll 1
        while (i < end) {
            uint32_t w, n = 1;
            if ((i & 3u) == 0u && i + 4u <= end) { w = __ldg(reinterpret_cast<const uint32_t *>(text + i)); n = 4; }
            else w = __ldg(text + i);
#pragma unroll 1
            for (uint32_t j = 0; j < n; ++j, ++i) {
                const uint32_t b = (w >> (8u * j)) & 0xffu;
                uint32_t e = 0;
                if (s - win_lo < win_rows) e = hot_next(s - win_lo + 1u, b);
                if (e == 0) e = any_next(s, b);              // outside the window, or leaving it: the true entry
                else e += win_lo - 1u;
                s = e;
                if (s < final_bound) hit<2>(i + 1u, s);
                if (FIRST && found) return s;
            }
        }
        return s;
    }

    template <bool REPORT, int EMIT>
    __device__ __forceinline__ uint32_t walk(uint32_t s, uint32_t i, uint32_t end)
    {
        if constexpr (EMIT == 2) return walk_emit(s, i, end);
        else {
            while (i < end && (i & 15u)) { s = byte_step<REPORT, EMIT>(s, i); ++i; }
            if (i + 16 <= end) {
                uint4 v0 = ld_text16(text + i);
                uint4 v1 = v0, v2 = v0;
                if (i + 32 <= readable) v1 = ld_text16(text + i + 16);
                if (i + 48 <= readable) v2 = ld_text16(text + i + 32);
                while (true) {
                    uint4 v3 = v2;
                    if (i + 64 <= readable) v3 = ld_text16(text + i + 48);
                    s = walk_group<REPORT, EMIT>(s, v0, i);
                    i += 16;
                    if (REPORT && FIRST && found) return s;
                    if (i + 16 > end) break;
                    v0 = v1; v1 = v2; v2 = v3;
                }
            }
            while (i < end) { s = byte_step<REPORT, EMIT>(s, i); ++i; }
            return s;
        }
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
template <int EMIT, typename SC>
__device__ __forceinline__ uint32_t scan_slice(const ScanArgs &a, SC &sc, uint32_t s, uint32_t h,
                                               uint32_t cs, uint32_t ce)
{
    uint32_t i = cs;
    uint32_t nb = hay_end(a, h);
    sc.found = false;
    while (i < ce) {
        if (i == nb) {                       // haystack boundary: next haystack starts at the root
            do { ++h; nb = hay_end(a, h); } while (nb == i);   // skips empty haystacks
            s = a.root;
            sc.found = false;
        }
        const uint32_t seg_end = min(ce, nb);
        s = sc.template walk<true, EMIT>(s, i, seg_end);
        i = seg_end;                         // FIRST kernels may have stopped early; the rest is irrelevant
    }
    return s;
}

template <typename E, bool RANGE, bool FIRST>
__global__ void __launch_bounds__(SCAN_THREADS, 1) ac_scan_kernel(const ScanArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint8_t s_cls[256];

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const E *gtab = static_cast<const E *>(a.table);

    // window rows, with targets outside the window replaced by 0
    E *s_tab = stage_window<E>(reinterpret_cast<E *>(smem_raw), gtab, a.win_lo, a.win_rows, a.ncls, tid, SCAN_THREADS);
    if (tid < 256) s_cls[tid] = a.cls_map[tid];
    __syncthreads();      // the only CTA-wide barrier: from here on warps run independently

    Scanner<E, RANGE, FIRST> sc;
    sc.gtab = gtab; sc.text = a.text;
    sc.ncls = a.ncls; sc.row_bytes = a.ncls * (uint32_t)sizeof(E);
    sc.win_lo = a.win_lo; sc.win_rows = a.win_rows;
    sc.fin_rel = (a.final_bound > a.win_lo) ? a.final_bound - a.win_lo + 1u : 1u;
    {   // opaque moves keep the shared-window addresses in registers instead of being rematerialised
        const uint32_t t0 = (uint32_t)__cvta_generic_to_shared(s_tab);
        const uint32_t c0 = (uint32_t)__cvta_generic_to_shared(s_cls);
        asm volatile("mov.u32 %0, %1;" : "=r"(sc.s_tab) : "r"(t0));
        asm volatile("mov.u32 %0, %1;" : "=r"(sc.s_cls) : "r"(c0));
    }
    sc.lo = a.range_lo; sc.n_used = a.n_used;
    sc.final_bound = a.final_bound; sc.readable = a.readable;
    sc.out = a.out; sc.cap = a.capacity;

    const uint32_t prior = a.counters[1];    // events of earlier launches in this call (stream-ordered)

    // A tile is 32 consecutive slices, one per lane; warps take tiles by ticket.
    while (true) {
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(&a.counters[0], 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= a.n_tiles) break;

        const uint32_t chunk_id = a.chunk_begin + tile * 32u + lane;
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
                uint32_t s = (ws == hb && h == 0) ? a.init_state : a.root;
                sc.report_from = cs;
                s = sc.template walk<true, 0>(s, ws, cs);
                s_cs = s;
                s = scan_slice<0>(a, sc, s, h, cs, ce);
                if (ce == a.total) { a.counters[2] = s; if (a.host_counters) a.host_counters[2] = s; }
                if (FIRST && sc.cnt) {
                    // publish the earliest event of the slice's first reporting haystack
                    const uint32_t hh = find_haystack(a, sc.e0p - 1);
                    atomicMin(&a.first_end[hh], sc.e0p);
                }
            }
        }
        __syncwarp();

        // warp prefix of the per-lane event counts
        uint32_t incl = sc.cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);

        const unsigned long long excl = tile_lookback(a, tile, total, prior, lane);

        if (sc.cnt) {
            const uint32_t off = (uint32_t)excl + (incl - sc.cnt);
            if (sc.cnt <= 2) {
                if (off < a.capacity) a.out[off] = make_uint2(sc.e0p, sc.e0s);
                if (sc.cnt == 2 && off + 1 < a.capacity) a.out[off + 1] = make_uint2(sc.e1p, sc.e1s);
            } else if (off < a.capacity) {
                // dense slice: walk it again from the saved entry state and write in place
                const bool very_dense = sc.cnt > DENSE_PAIRED_MIN;
                sc.obase = off;
                sc.cnt = 0;
                sc.have_pend = false;
                if (very_dense) { scan_slice<2>(a, sc, s_cs, h, cs, ce); sc.flush_pending(); }
                else scan_slice<1>(a, sc, s_cs, h, cs, ce);
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------- full walk, text by TMA --
//
// ac_scan_kernel is bound by the L1TEX data stage, which its two kinds of loads share: the table lookups (one LDS of
// 32 random entries = 4.1 wavefronts per haystack byte-step of a warp) and the text (one LDG.128 per lane per 16 bytes,
// 32 lanes in 32 different lines = 42 wavefronts per load, 2.6 per byte-step — 39 % of the stage, ncu).  One more
// divergent request per 128 bytes (an L2 prefetch) made the kernel 12 % slower: that pipe is the limit, not latency.
// This variant takes the text out of it.  The stream is described to the TMA unit as a 2-D array of rows of `chunk`
// bytes — row r is slice r — and ONE cp.async.bulk.tensor per warp fetches a box of 32 bytes x 32 rows: the next 32
// bytes of all 32 slices of the warp's tile, laid down in shared memory lane after lane.  Lanes read their 32 bytes
// with two conflict-free LDS.128 (8 wavefronts per 32 byte-steps instead of 84), a two-stage ring per warp with one
// mbarrier per stage keeps one box in flight (a warp needs ~4 us for 32 byte-steps, an HBM trip takes 1).  The ring
// costs 64 KB of the table window.  Tiles the box cannot serve — a haystack boundary inside a slice, a ragged last
// tile, patterns longer than 32 bytes (the warm-up must fit one box) — take the loads of ac_scan_kernel.
constexpr uint32_t TMA_BOX_BYTES = 32;                        // bytes per slice and box
constexpr uint32_t TMA_STAGE_BYTES = 32 * TMA_BOX_BYTES;      // one box: 32 slices
constexpr uint32_t TMA_STAGES = 2;
constexpr uint32_t TMA_RING_BYTES = (SCAN_THREADS / 32) * TMA_STAGES * TMA_STAGE_BYTES;

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    // (bounded: a box that never arrives — a tensor map the driver rejected would be one — must abort the launch, not hang it)
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spins > (1u << 22)) __trap();
    }
}
// box (x bytes, y rows) of the 2-D view of the stream -> shared memory at dst, completion counted on bar
__device__ __forceinline__ void tma_load_box(uint32_t dst, const void *tmap, int32_t x, int32_t y, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(bar) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

template <typename E, bool RANGE>
__global__ void __launch_bounds__(SCAN_THREADS, 1) ac_scan_tma_kernel(const ScanArgs a, const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint8_t s_cls[256];
    __shared__ __align__(8) unsigned long long s_bar[SCAN_THREADS / 32][TMA_STAGES];
    // the ring first, on a 128-byte boundary (what the TMA unit wants of its destination), then the table window
    const uint32_t ring_base = ((uint32_t)__cvta_generic_to_shared(smem_raw) + 127u) & ~127u;
    E *s_tab = reinterpret_cast<E *>(smem_raw + (ring_base - (uint32_t)__cvta_generic_to_shared(smem_raw)) + TMA_RING_BYTES);   // (moved by stage_window)

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u, warp = tid >> 5;
    const E *gtab = static_cast<const E *>(a.table);
    const uint32_t ring = ring_base + warp * (TMA_STAGES * TMA_STAGE_BYTES);
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&s_bar[warp][0]);

    s_tab = stage_window<E>(s_tab, gtab, a.win_lo, a.win_rows, a.ncls, tid, SCAN_THREADS);
    if (tid < 256) s_cls[tid] = a.cls_map[tid];
    if (lane == 0) {
#pragma unroll
        for (uint32_t st = 0; st < TMA_STAGES; ++st) mbar_init(bar0 + st * 8u, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();      // the only CTA-wide barrier: from here on warps run independently

    Scanner<E, RANGE, false> sc;
    sc.gtab = gtab; sc.text = a.text;
    sc.ncls = a.ncls; sc.row_bytes = a.ncls * (uint32_t)sizeof(E);
    sc.win_lo = a.win_lo; sc.win_rows = a.win_rows;
    sc.fin_rel = (a.final_bound > a.win_lo) ? a.final_bound - a.win_lo + 1u : 1u;
    {
        const uint32_t t0 = (uint32_t)__cvta_generic_to_shared(s_tab);
        const uint32_t c0 = (uint32_t)__cvta_generic_to_shared(s_cls);
        asm volatile("mov.u32 %0, %1;" : "=r"(sc.s_tab) : "r"(t0));
        asm volatile("mov.u32 %0, %1;" : "=r"(sc.s_cls) : "r"(c0));
    }
    sc.lo = a.range_lo; sc.n_used = a.n_used;
    sc.final_bound = a.final_bound; sc.readable = a.readable;
    sc.out = a.out; sc.cap = a.capacity;
    sc.have_pend = false;

    const uint32_t prior = a.counters[1];
    const uint32_t n_boxes = a.chunk / TMA_BOX_BYTES;         // boxes per tile (the warm-up box not counted)
    uint32_t phase = 0;                                       // bit st = parity the next wait on stage st expects

    while (true) {
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(&a.counters[0], 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= a.n_tiles) break;

        const uint32_t row0 = a.chunk_begin + tile * 32u;
        const uint32_t chunk_id = row0 + lane;
        const bool active = chunk_id < a.chunk_end;
        uint32_t cs = 0, ce = 0, h = 0, s_cs = 0, hb = 0;
        sc.cnt = 0; sc.found = false;
        bool boxed = active;
        if (active) {
            cs = chunk_id * a.chunk;
            ce = min(cs + a.chunk, a.total);
            h = find_haystack(a, cs);
            hb = hay_begin(a, h);
            // the box serves a slice that is complete, lies inside one haystack and starts either at its haystack's
            // first byte (no warm-up) or at least one box behind it (the warm-up is the 32 bytes before the slice)
            boxed = ce - cs == a.chunk && hay_end(a, h) >= ce && (cs == hb || cs - hb >= TMA_BOX_BYTES);
        }
        boxed = __all_sync(0xffffffffu, boxed);

        if (boxed) {
            // box q = 0: the 32 bytes before every slice (the tail of the rows above); q = 1 .. n_boxes: the slices
            auto issue = [&](uint32_t q) {
                if (lane == 0) {
                    const uint32_t st = q & 1u;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the lanes' reads of this stage are over
                    mbar_expect_tx(bar0 + st * 8u, TMA_STAGE_BYTES);
                    if (q == 0) tma_load_box(ring + st * TMA_STAGE_BYTES, &tmap, (int32_t)(a.chunk - TMA_BOX_BYTES), (int32_t)row0 - 1, bar0 + st * 8u);
                    else tma_load_box(ring + st * TMA_STAGE_BYTES, &tmap, (int32_t)((q - 1u) * TMA_BOX_BYTES), (int32_t)row0, bar0 + st * 8u);
                }
            };
            issue(0);
            issue(1);
            uint32_t s = a.root;
            sc.report_from = cs;
            for (uint32_t q = 0; q <= n_boxes; ++q) {
                const uint32_t st = q & 1u;
                mbar_wait(bar0 + st * 8u, (phase >> st) & 1u);
                phase ^= 1u << st;
                const uint32_t at = ring + st * TMA_STAGE_BYTES + lane * TMA_BOX_BYTES;
                const uint4 v0 = lds128(at), v1 = lds128(at + 16u);
                __syncwarp();
                if (q + 2u <= n_boxes) issue(q + 2u);
                // ONE copy of the unrolled 16-step switch serves the warm-up box and the slice's boxes, both halves
                const uint32_t i = cs + q * TMA_BOX_BYTES - TMA_BOX_BYTES;
                if (q != 0 || cs != hb) {
#pragma unroll 1
                    for (uint32_t g = 0; g < 2u; ++g) s = sc.template walk_group<true, 0>(s, g ? v1 : v0, i + 16u * g);
                } else s = (h == 0) ? a.init_state : a.root;
                if (q == 0) s_cs = s;
            }
            if (ce == a.total) { a.counters[2] = s; if (a.host_counters) a.host_counters[2] = s; }
        } else if (active) {
            uint32_t ws = (cs - hb > a.halo) ? ((cs - a.halo) & ~15u) : hb;
            if (ws < hb) ws = hb;
            uint32_t s = (ws == hb && h == 0) ? a.init_state : a.root;
            sc.report_from = cs;
            s = sc.template walk<true, 0>(s, ws, cs);
            s_cs = s;
            s = scan_slice<0>(a, sc, s, h, cs, ce);
            if (ce == a.total) { a.counters[2] = s; if (a.host_counters) a.host_counters[2] = s; }
        }
        __syncwarp();

        // warp prefix of the per-lane event counts
        uint32_t incl = sc.cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        const unsigned long long excl = tile_lookback(a, tile, total, prior, lane);
        if (sc.cnt) {
            const uint32_t off = (uint32_t)excl + (incl - sc.cnt);
            if (sc.cnt <= 2) {
                if (off < a.capacity) a.out[off] = make_uint2(sc.e0p, sc.e0s);
                if (sc.cnt == 2 && off + 1 < a.capacity) a.out[off + 1] = make_uint2(sc.e1p, sc.e1s);
            } else if (off < a.capacity) {
                // dense slice: walk it again from the saved entry state and write in place
                const bool very_dense = sc.cnt > DENSE_PAIRED_MIN;
                sc.obase = off;
                sc.cnt = 0;
                sc.have_pend = false;
                if (very_dense) { scan_slice<2>(a, sc, s_cs, h, cs, ce); sc.flush_pending(); }
                else scan_slice<1>(a, sc, s_cs, h, cs, ce);
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------- small texts ------
//
// ahocorasick_match() on ONE short haystack (examples/benchmark.php:55-76 calls it 256 times on 8 KiB strings): what
// such a call costs is not the walk but the round trips around it.  One CTA walks the text straight from the dense
// table (no shared-memory window to stage: a few KiB of text touch a few rows), orders the events with a block
// prefix sum and writes them — with their count and the end state — into mapped pinned host memory: the call is
// one host-to-device copy, one launch, one wait.

constexpr uint32_t SMALL_TEXT_BYTES = 32u << 10;      // texts up to this size take the one-CTA path
constexpr int SMALL_THREADS = 1024;

struct SmallArgs {
    const uint8_t *text;          // device copy of the text, padded like ScanArgs::text
    uint32_t total, readable, chunk, halo;
    const void *table;
    const uint8_t *cls_map;
    uint32_t ncls, final_bound, root, range_lo, n_used, init_state;
    uint2 *out;                   // mapped host memory: events {end offset, state}, ascending
    uint32_t capacity;
    uint32_t *hdr;                // mapped host memory: [0] events [1] end state
};

template <typename E, bool RANGE>
__global__ void __launch_bounds__(SMALL_THREADS, 1) ac_small_kernel(const SmallArgs a)
{
    __shared__ uint8_t s_cls[256];
    __shared__ uint32_t s_warp[SMALL_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    for (uint32_t i = tid; i < 256u; i += blockDim.x) s_cls[i] = a.cls_map[i];     // (the CTA may be a single warp)
    __syncthreads();

    Scanner<E, RANGE, false> sc;
    sc.gtab = static_cast<const E *>(a.table); sc.text = a.text;
    sc.ncls = a.ncls; sc.row_bytes = a.ncls * (uint32_t)sizeof(E);
    sc.win_lo = a.final_bound; sc.win_rows = 0;           // every step reads the true entry (L1 / L2)
    sc.s_tab = 0; sc.fin_rel = 1;
    sc.s_cls = (uint32_t)__cvta_generic_to_shared(s_cls);
    sc.lo = a.range_lo; sc.n_used = a.n_used;
    sc.final_bound = a.final_bound; sc.readable = a.readable;
    sc.out = a.out; sc.cap = a.capacity;
    sc.found = false; sc.cnt = 0; sc.have_pend = false;
    sc.e0p = sc.e0s = sc.e1p = sc.e1s = 0;

    const uint32_t cs = min(tid * a.chunk, a.total), ce = min(cs + a.chunk, a.total);
    uint32_t s_cs = a.root;
    if (cs < ce) {
        const uint32_t ws = (cs > a.halo) ? ((cs - a.halo) & ~15u) : 0u;
        sc.report_from = cs;
        s_cs = sc.template walk<true, 0>(ws == 0u ? a.init_state : a.root, ws, cs);
        const uint32_t s_end = sc.template walk<true, 0>(s_cs, cs, ce);
        if (ce == a.total) a.hdr[1] = s_end;
    }
    // block prefix of the per-thread event counts
    uint32_t incl = sc.cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t base = 0, all = 0;
#pragma unroll
    for (int w = 0; w < SMALL_THREADS / 32; ++w) {
        const uint32_t v = (uint32_t)w < blockDim.x / 32u ? s_warp[w] : 0u;
        if ((uint32_t)w < warp) base += v;
        all += v;
    }
    if (tid == 0) a.hdr[0] = all;
    const uint32_t off = base + incl - sc.cnt;
    if (sc.cnt && off < a.capacity) {
        if (sc.cnt <= 2) {
            a.out[off] = make_uint2(sc.e0p, sc.e0s);
            if (sc.cnt == 2 && off + 1 < a.capacity) a.out[off + 1] = make_uint2(sc.e1p, sc.e1s);
        } else {
            sc.obase = off;
            sc.cnt = 0;
            sc.template walk<true, 1>(s_cs, cs, ce);
        }
    }
}

// ------------------------------------------------------- hit expansion ----
//
// The reference's callback turns every event into one record per reported pattern
// (src/php_ahocorasick.c:555-584: pos, start = pos - length, the pattern's identity).
// On the device: events -> hits {haystack index, end offset inside it, start offset,
// pattern index in acceptance order}, longest pattern first inside an event, events in order.

struct HitArgs {
    const uint2 *events;          // {end offset in the stream, state}, ascending
    uint32_t n_events;
    const uint32_t *out_off;      // per reporting state s: patterns out_idx[out_off[s-1] .. out_off[s])
    const uint32_t *out_idx;      // pattern index in acceptance order
    const uint32_t *pat_len;      // per accepted pattern: length
    const uint32_t *hay_off;      // haystack offsets (or nullptr with uniform_len)
    uint32_t n_hay, uniform_len;
    uint32_t *block_sum;          // per HIT_THREADS events: hits
    uint4 *hits;                  // out: {text_idx, end, start, pattern}
    unsigned long long capacity;
    unsigned long long *total;    // out: number of hits
};

constexpr int HIT_THREADS = 256;

// hits per block of HIT_THREADS events
__global__ void __launch_bounds__(HIT_THREADS) ac_hit_count_kernel(const HitArgs a)
{
    __shared__ uint32_t s_warp[HIT_THREADS / 32];
    const uint32_t i = blockIdx.x * HIT_THREADS + threadIdx.x;
    uint32_t c = 0;
    if (i < a.n_events) {
        const uint32_t s = a.events[i].y;
        c = a.out_off[s] - a.out_off[s - 1];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31u) == 0) s_warp[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < HIT_THREADS / 32; ++w) t += s_warp[w];
        a.block_sum[blockIdx.x] = t;
    }
}

// every CTA adds up the earlier blocks' sums itself, scans its own events' counts and writes their hits
__global__ void __launch_bounds__(HIT_THREADS) ac_hit_write_kernel(const HitArgs a)
{
    __shared__ unsigned long long s_prev[HIT_THREADS / 32];
    __shared__ uint32_t s_warp[HIT_THREADS / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    unsigned long long part = 0;
    for (uint32_t j = tid; j < blockIdx.x; j += HIT_THREADS) part += a.block_sum[j];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if (lane == 0) s_prev[warp] = part;

    const uint32_t i = blockIdx.x * HIT_THREADS + tid;
    uint32_t end = 0, s = 0, b = 0, c = 0;
    if (i < a.n_events) {
        const uint2 e = a.events[i];
        end = e.x; s = e.y;
        b = a.out_off[s - 1];
        c = a.out_off[s] - b;
    }
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long base = 0;
#pragma unroll
    for (int w = 0; w < HIT_THREADS / 32; ++w) {
        base += s_prev[w];
        if ((uint32_t)w < warp) base += s_warp[w];
    }
    unsigned long long o = base + incl - c;
    if (i == a.n_events - 1) *a.total = o + c;
    if (c == 0) return;
    // haystack of this event: the one that contains byte end-1
    uint32_t h, hb;
    if (a.uniform_len) { h = (end - 1u) / a.uniform_len; hb = h * a.uniform_len; }
    else {
        uint32_t lo = 0, hi = a.n_hay;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(a.hay_off + mid) <= end - 1u) lo = mid; else hi = mid;
        }
        h = lo; hb = __ldg(a.hay_off + h);
    }
    const uint32_t pos = end - hb;
    for (uint32_t k = 0; k < c; ++k, ++o) {
        if (o >= a.capacity) break;
        const uint32_t pid = __ldg(a.out_idx + b + k);
        a.hits[o] = make_uint4(h, pos, pos - __ldg(a.pat_len + pid), pid);
    }
}

} // namespace acb200
