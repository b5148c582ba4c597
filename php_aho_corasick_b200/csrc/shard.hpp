// shard.hpp — how one host call (ahocorasick_match / ahocorasick_match_batch) is cut into slabs and spread
// over the GPUs of the box.  Pure host arithmetic, no CUDA: tested on the CPU through acb200_plan_slabs().
//
// The haystacks of a call form one virtual byte stream (haystack i = stream bytes [off[i], off[i+1])).  The
// stream is cut into slabs of at most `slab_bytes` — a whole number of rounds over the devices — and slab i goes
// to device i mod n_dev: every device works on the same region of the stream at the same time, so the calling
// thread, which must run the callbacks in stream order, replays slab i while the devices are busy with the
// next round (one contiguous range per device left 7/8 of an eight-GPU call's events to be replayed after the
// last copy had finished: measured 74.6 ms per 8 GiB of which 28 ms were that tail).  A cut may fall anywhere — also inside a haystack: such a slab carries the (Lmax-1) bytes before
// the cut in front of its own bytes ("halo"), the device walks them from the root like any haystack start, and
// the host drops the events that end inside the halo (they belong to the slab before).  After Lmax-1 bytes the
// state reached from the root equals the state of an uninterrupted walk — the argument of scan_kernels.cuh, and
// what the reference does sequentially by carrying last_node across chunks (src/multifast/ahocorasick.c:191-194,
// 236-238).  That one rule covers a batch spread over 8 GPUs, one haystack larger than a launch can address, and
// ONE large haystack (the virus-scan and adversarial shapes) split over several GPUs.
#pragma once

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace acb200 {

struct SlabPlan {
    uint64_t b0, b1;      // the stream bytes whose event ends (b0, b1] this slab reports
    uint32_t halo;        // bytes before b0 that travel with it (0 when b0 is the start of a haystack)
    size_t h_first;       // haystack that contains byte b0
    size_t h_end;         // one past the last haystack that starts before b1
    int device_slot;      // index into the call's device list
};

// off: n + 1 ascending stream offsets (off[0] == 0).  halo_max = Lmax - 1.
inline std::vector<SlabPlan> plan_slabs(const uint64_t *off, size_t n, uint32_t halo_max, int n_dev, uint64_t slab_bytes)
{
    std::vector<SlabPlan> out;
    const uint64_t total = n ? off[n] : 0;
    if (total == 0 || n_dev < 1 || slab_bytes == 0) return out;
    // as many slabs as the size limit asks for, at least one per device, a whole number of rounds over the devices
    uint64_t k = std::max<uint64_t>((total + slab_bytes - 1) / slab_bytes, (uint64_t)n_dev);
    k = (k + (uint64_t)n_dev - 1) / (uint64_t)n_dev * (uint64_t)n_dev;
    k = std::min(k, total);
    const uint64_t nominal = total / k;
    // a cut lands on the nearest haystack boundary when one is close (no halo, equal-length batches stay uniform)
    auto snap = [&](uint64_t t, uint64_t lo) -> uint64_t {
        if (t <= lo) return lo;
        if (t >= total) return total;
        const uint64_t *p = std::lower_bound(off, off + n + 1, t);          // first boundary >= t
        uint64_t best = t, dist = nominal / 8 + 1;
        if (p != off + n + 1 && *p - t < dist && *p > lo && *p < total) { best = *p; dist = *p - t; }
        if (p != off) { const uint64_t q = *(p - 1); if (t - q < dist && q > lo && q < total) best = q; }
        return best;
    };
    auto push = [&](uint64_t lo, uint64_t hi) {
        SlabPlan p;
        p.b0 = lo; p.b1 = hi;
        p.h_first = (size_t)(std::upper_bound(off, off + n + 1, lo) - off) - 1;     // off[h] <= lo < off[h+1]
        p.h_end = (size_t)(std::lower_bound(off, off + n + 1, hi) - off);           // haystacks h < h_end start before hi
        p.halo = (uint32_t)std::min<uint64_t>(halo_max, lo - off[p.h_first]);
        p.device_slot = (int)(out.size() % (size_t)n_dev);
        out.push_back(p);
    };
    uint64_t lo = 0;
    for (uint64_t s = 0; s < k && lo < total; ++s) {
        uint64_t hi = (s == k - 1) ? total : snap(total / k * (s + 1) + total % k * (s + 1) / k, lo);
        if (hi - lo > slab_bytes + slab_bytes / 8) hi = lo + slab_bytes;          // snapping never grows a slab beyond 9/8
        if (hi == lo) continue;
        push(lo, hi);
        lo = hi;
    }
    while (lo < total) {                                                          // (what a clamped last slab left over)
        const uint64_t hi = std::min(total, lo + slab_bytes);
        push(lo, hi);
        lo = hi;
    }
    return out;
}

// Offsets of the slab's haystack pieces inside its device stream [halo bytes | bytes b0..b1): rel[0] = 0,
// rel[i] = end of piece i-1.  Piece 0 includes the halo.  rel gets h_end - h_first + 1 entries.
inline void slab_rel_offsets(const SlabPlan &p, const uint64_t *off, std::vector<uint64_t> &rel)
{
    const size_t cnt = p.h_end - p.h_first;
    rel.resize(cnt + 1);
    rel[0] = 0;
    for (size_t i = 1; i <= cnt; ++i) {
        const uint64_t e = std::min(std::max(off[p.h_first + i], p.b0), p.b1);
        rel[i] = p.halo + (e - p.b0);
    }
}

} // namespace acb200
