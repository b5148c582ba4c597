// cabi.cpp — extern "C" surface declared in include/acb200.h.
//
// Mirrors the five MultiFast entry points the reference's Zend glue calls
// (src/multifast/ahocorasick.h:73-80) and adds the batched / event-level /
// device-resident entry points.  The device does the scan; this file replays
// the compact (position, state) events through the caller's callback exactly
// as the reference's loop would have called it
// (src/multifast/ahocorasick.c:214-233).
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime_api.h>

#include "acb200.h"
#include "automaton.hpp"
#include "engine.hpp"
#include "filter_hash.hpp"
#include "helper_pool.hpp"
#include "shard.hpp"
#include "staging_copy.hpp"

using namespace acb200;

struct ac_trie {
    HostTrie trie;
    FlatAutomaton flat;
    Engine engine;             // the automaton on its primary device
    std::vector<std::unique_ptr<Engine>> replicas;   // the same automaton on further GPUs (acb200_set_devices)
    int device = 0;            // primary device, fixed when the handle is created
    bool open = true;          // patterns may still be added (reference: trie_open)
    bool device_ok = false;    // finalize reached the device
    uint32_t last_state = ROOT_STATE;   // keep=1 continuation (reference: last_node)
    size_t base_position = 0;  // keep=1 continuation (reference: base_position)
    std::vector<uint64_t> gather_off;
    std::deque<std::string> blob_arena;   // pattern bytes / string ids of an automaton loaded from a blob
    ACB200_STATS_t stats{};    // statistics of the most recent search (summed over slabs and devices)
    uint64_t slab_bytes = 64ull << 20;
    std::vector<std::unique_ptr<HelperPool>> pools;   // per device slot of a call, made on first use (pool_for)
};

static inline size_t patterns_of(const ac_trie *t, uint32_t state, const AC_PATTERN_t **p)
{
    if (state == 0 || state >= t->flat.final_bound) { if (p) *p = nullptr; return 0; }
    const uint32_t i = state - 1;
    const uint64_t b = t->flat.out_off[i], e = t->flat.out_off[i + 1];
    if (p) *p = t->flat.out_pat.data() + b;
    return (size_t)(e - b);
}


// ------------------------------------------------------------------------------------------------------------
// Replicas: the same finalized automaton on further GPUs of the box, so that ONE host call (one PHP process, one
// thread — src/php_ahocorasick.c:664-746) can spread its haystacks over all of them.
// ------------------------------------------------------------------------------------------------------------

static int set_devices(ac_trie *t, const int *devices, size_t n)
{
    if (t->open || !t->device_ok) { set_error("automaton is not finalized"); return -1; }
    // the primary always takes part; every further entry of the list is one more pipeline (an ordinal listed
    // twice gets two: two pipelines on one GPU are legal, and how a one-GPU box tests this path)
    std::vector<int> want(devices, devices + n);
    auto self = std::find(want.begin(), want.end(), t->device);
    if (self != want.end()) want.erase(self);
    t->replicas.clear();
    std::vector<std::unique_ptr<Engine>> built(want.size());
    std::vector<std::string> errs(want.size());
    std::vector<std::thread> th;
    for (size_t i = 0; i < want.size(); ++i)       // one thread per replica: the builds run side by side
        th.emplace_back([&, i] {
            std::unique_ptr<Engine> e(new (std::nothrow) Engine());
            if (e && e->build(t->flat, want[i], &t->engine)) {
                e->info.n_patterns = t->engine.info.n_patterns;
                e->info.finalized = 1;
                built[i] = std::move(e);
            } else errs[i] = get_error();
        });
    for (auto &x : th) x.join();
    cudaSetDevice(t->device);
    int rc = 0;
    for (size_t i = 0; i < want.size(); ++i) {
        if (built[i]) t->replicas.push_back(std::move(built[i]));
        else { set_error("replica on device " + std::to_string(want[i]) + ": " + errs[i]); rc = -1; }
    }
    return rc;
}

// ACB200_DEVICES=all | 0,1,2,...  — the knob a PHP deployment has (there is no INI entry in the reference either,
// src/php_ahocorasick.c:92-95); unset: one GPU.
static void replicate_from_env(ac_trie *t)
{
    const char *e = getenv("ACB200_DEVICES");
    if (!e || !*e) return;
    std::vector<int> devs;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return;
    if (!strcmp(e, "all")) { for (int d = 0; d < n; ++d) devs.push_back(d); }
    else {
        for (const char *p = e; *p;) {
            char *end = nullptr;
            const long d = strtol(p, &end, 10);
            if (end == p) break;
            if (d >= 0 && d < n) devs.push_back((int)d);
            p = (*end == ',') ? end + 1 : end;
        }
    }
    const std::string keep = get_error();
    if (set_devices(t, devs.data(), devs.size()) == 0) set_error(keep);
}

// ------------------------------------------------------------------------------------------------------------
// One host call = slabs spread over the devices (shard.hpp), one worker thread per device running a double-
// buffered pipeline (gather into pinned staging -> H2D -> kernels -> D2H of the compact events), the calling
// thread replaying the events through the callback in global order while the workers run ahead.
// ------------------------------------------------------------------------------------------------------------

namespace {

struct HaySource {
    const char *flat = nullptr;          // haystacks laid end to end ...
    const AC_TEXT_t *texts = nullptr;    // ... or scattered (ac_trie_search_batch: PHP strings)
    const uint64_t *off = nullptr;       // n + 1 stream offsets
    size_t n = 0;
    bool pinned = false;                 // flat buffer is page-locked: slabs are DMA'd straight from it
    bool stream_stores = true;           // gather with non-temporal stores (ACB200_GATHER_NT=0: plain memcpy, for A/B runs)

    // stream bytes [b, e) -> dst (pinned staging the copy engine reads next); h = a haystack that starts at or before b
    void copy(char *dst, uint64_t b, uint64_t e, size_t h) const
    {
        if (flat) { copy_to_staging(dst, flat + b, (size_t)(e - b), stream_stores); staging_fence(); return; }
        while (h + 1 <= n && off[h + 1] <= b) ++h;
        while (b < e) {
            const uint64_t he = std::min(off[h + 1], e);
            if (he > b) { copy_to_staging(dst, texts[h].astring + (b - off[h]), (size_t)(he - b), stream_stores); dst += he - b; b = he; }
            ++h;
        }
        staging_fence();
    }
};

// one event as the slab worker hands it to the calling thread: everything the callback needs already looked up
struct ResolvedEvent {
    uint64_t end;                        // exclusive end offset inside its haystack
    const AC_PATTERN_t *patterns;        // the state's output list ...
    uint32_t size;                       // ... and its length
    uint32_t state;
    size_t text_idx;
};

// ... and as a slab worker hands it over: 16 bytes.  The calling thread looks the output list up itself — what bounds an
// 8-GPU call after the copies is the ONE thread that must run 8 M callbacks in stream order, and reading 32-byte
// records written by eight other cores costs it more (6.3 ns per event, replay micro-benchmark) than two loads from
// the output-list offsets it keeps in its own cache (4.8 ns).
struct SlabEvent {
    uint64_t end;                        // exclusive end offset inside its haystack
    uint32_t state;
    uint32_t text_idx;
};

struct SlabResult {
    std::vector<SlabEvent> ev;
    ACB200_STATS_t st{};
    uint32_t end_state = 0;
    bool ready = false, ok = true;
    std::string err;
};

struct ShardRun {
    std::mutex m;
    std::condition_variable cv;
    std::vector<SlabResult> res;
    std::atomic<bool> abort{false};
};

bool is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// threads that copy for one device slot of a call: hw cores, n_dev devices at work side by side
// (16-core host, one GPU, 1 GiB of 8 KiB strings: 4 / 8 / 12 / 16 threads gather at 27.6 / 36.1 / 38.0 / 38.4 GB/s
// with non-temporal stores, 21.5 / 25.8 / 23.4 / 24.9 GB/s with memcpy, whose read-for-ownership traffic also slows
// the copy engine's reads of the slab before: 27-36 ms of H2D per GiB instead of 20 — profiles/r02_gather_probe.txt)
int copy_threads(int n_dev)
{
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    int helpers = (int)std::max(1u, std::min(12u, hw * 3u / (4u * (unsigned)std::max(1, n_dev))));
    if (const char *e = getenv("ACB200_GATHER_THREADS")) { const int v = atoi(e); if (v >= 1 && v <= 64) helpers = v; }
    return helpers;
}

// the parked helper threads of device slot `slot` (the caller of HelperPool::run is one more)
HelperPool &pool_for(ac_trie *t, size_t slot, int threads)
{
    if (t->pools.size() <= slot) t->pools.resize(slot + 1);
    // (a pool only grows: a handle that is used for one-GPU and for all-GPU calls in turn keeps its threads)
    if (!t->pools[slot] || t->pools[slot]->wanted() < threads - 1) t->pools[slot].reset(new HelperPool(threads - 1));
    t->pools[slot]->set_limit(threads - 1);
    return *t->pools[slot];
}

// haystack that contains stream byte b (off[h] <= b < off[h + 1]; empty haystacks skipped)
inline size_t haystack_at(const HaySource &src, uint64_t b)
{
    return (size_t)(std::upper_bound(src.off, src.off + src.n + 1, b) - src.off) - 1;
}

constexpr uint64_t COPY_PIECE_BYTES = 256u << 10;

// The gather of one slab into pinned staging buffer `buf` of `eng` and its upload, in pieces handed to the slot's
// helper threads (one core copies 5-10 GB/s, PCIe takes 55): every piece's DMA is queued as soon as the piece is there.
bool gather_and_upload_slab(const HaySource &src, Engine *eng, int buf, const SlabPlan &p, HelperPool *pool)
{
    const uint64_t b = p.b0 - p.halo, e = p.b1;
    const uint64_t len = e - b;
    char *dst = eng->slab_staging(buf, (size_t)len);
    if (!dst) return false;
    if (!pool || len < (1u << 20)) {
        src.copy(dst, b, e, p.h_first);
        return eng->slab_upload_async(buf, dst, (size_t)len);
    }
    if (!eng->slab_upload_begin(buf, (size_t)len)) return false;
    // two or three pieces per thread: every piece costs a cudaMemcpyAsync call, and those serialise inside the driver
    // (256 KiB pieces = 256 calls per 64 MiB slab from twelve threads: 1 GiB of strings took 38.5 ms instead of 31)
    const uint64_t per = len / (3u * (uint64_t)(pool->helpers() + 1));
    const uint64_t piece = std::max<uint64_t>(COPY_PIECE_BYTES, (per + 4095) & ~(uint64_t)4095);
    const int n_pieces = (int)((len + piece - 1) / piece);
    std::mutex em;
    std::string err;
    bool ok = true;
    pool->run(n_pieces, [&](int i) {
        const uint64_t pb = b + (uint64_t)i * piece, pe = std::min(e, pb + piece);
        src.copy(dst + (pb - b), pb, pe, src.flat ? 0 : haystack_at(src, pb));
        if (!eng->slab_upload_part(buf, (size_t)(pb - b), (size_t)(pe - pb))) {
            std::lock_guard<std::mutex> g(em);
            ok = false; err = get_error();            // (errors are per thread)
        }
    });
    if (!ok) { set_error(err); return false; }
    return eng->slab_upload_end(buf);
}

void add_stats(ACB200_STATS_t &sum, const ACB200_STATS_t &s)
{
    sum.bytes += s.bytes; sum.events += s.events; sum.kernel_launches += s.kernel_launches;
    sum.kernel_ms += s.kernel_ms; sum.filter_ms += s.filter_ms; sum.verify_ms += s.verify_ms;
    sum.reorder_ms += s.reorder_ms; sum.d2h_ms += s.d2h_ms; sum.h2d_ms += s.h2d_ms;
    sum.flagged_words += s.flagged_words; sum.dense_tiles += s.dense_tiles;
    sum.chunk_bytes = s.chunk_bytes; sum.halo_bytes = s.halo_bytes; sum.filtered = s.filtered;
}

void shard_worker(const ac_trie *t, Engine *eng, const HaySource &src, const std::vector<SlabPlan> &plans, const std::vector<size_t> &mine,
                  bool first_only, uint32_t init_state, HelperPool *pool, ShardRun &run)
{
    std::vector<uint64_t> rel;
    auto fail = [&](size_t from) {
        const std::string err = get_error();
        std::lock_guard<std::mutex> g(run.m);
        for (size_t k = from; k < mine.size(); ++k) { run.res[mine[k]].ok = false; run.res[mine[k]].err = err; run.res[mine[k]].ready = true; }
        run.cv.notify_all();
    };
    auto prepare = [&](size_t k) -> bool {
        const SlabPlan &p = plans[mine[k]];
        const size_t n_bytes = (size_t)(p.halo + (p.b1 - p.b0));
        const int buf = (int)(k & 1);
        if (src.pinned) return eng->slab_upload_async(buf, src.flat + (p.b0 - p.halo), n_bytes);
        return gather_and_upload_slab(src, eng, buf, p, pool);
    };
    if (mine.empty()) return;
    if (!prepare(0)) { fail(0); return; }
    for (size_t k = 0; k < mine.size(); ++k) {
        if (run.abort.load(std::memory_order_relaxed)) { set_error("search stopped"); fail(k); return; }
        if (k + 1 < mine.size() && !prepare(k + 1)) { fail(k); return; }
        const SlabPlan &p = plans[mine[k]];
        slab_rel_offsets(p, src.off, rel);
        // findAll=false on the device keeps one event per haystack piece; a piece that starts with a halo could
        // lose its first real event to one inside the halo, so such a slab reports everything and the host picks
        const bool fo = first_only && p.halo == 0;
        const uint32_t init = (mine[k] == 0) ? init_state : ROOT_STATE;
        if (!eng->scan_slab((int)(k & 1), rel.data(), p.h_end - p.h_first, fo, init)) { fail(k); return; }
        SlabResult r;
        {
            // the slab's events in the caller's coordinates (haystack index, offset inside it), on the worker: every
            // GPU's worker does its own in parallel, the calling thread owns the callbacks and is the serial part
            const PackedEvent *pe = eng->host_events();
            const size_t ne = eng->n_events();
            r.ev.reserve(ne);
            size_t h = p.h_first;
            const uint64_t base = p.b0 - p.halo;
            for (size_t i = 0; i < ne; ++i) {
                if (pe[i].end <= p.halo) continue;               // ends inside the halo: the slab before reported it
                const uint64_t g = base + pe[i].end;
                while (g > src.off[h + 1]) ++h;
                r.ev.push_back(SlabEvent{g - src.off[h], pe[i].state, (uint32_t)h});
            }
        }
        r.st = eng->stats;
        r.st.h2d_ms = eng->slab_h2d_ms((int)(k & 1));
        r.end_state = eng->end_state();
        r.ready = true;
        {
            std::lock_guard<std::mutex> g(run.m);
            run.res[mine[k]] = std::move(r);
        }
        run.cv.notify_all();
    }
}

// below this many bytes a further GPU costs more than it brings (16 MiB with the default 64 MiB slabs)
inline uint64_t min_bytes_per_device(const ac_trie *t) { return std::max<uint64_t>(1, t->slab_bytes / 4); }

// sink(text_idx, position, state) -> non-zero = stop (that haystack, or with stop_all the whole search).
// Returns 0, 1 (stopped, stop_all only) or -1.
template <class Sink>
int sharded_search(ac_trie *t, const HaySource &src, bool first_only, uint32_t init_state, bool stop_all,
                   uint32_t *end_state, Sink &&sink)
{
    const uint64_t total = src.off[src.n];
    if (src.n > 0xffffffffull) { set_error("more than 2^32 haystacks in one call"); return -1; }
    std::vector<Engine *> engines{&t->engine};
    for (auto &r : t->replicas) engines.push_back(r.get());
    const int n_dev = (int)std::min<uint64_t>(engines.size(), std::max<uint64_t>(1, total / min_bytes_per_device(t)));
    const uint32_t halo_max = t->flat.max_pattern_len ? t->flat.max_pattern_len - 1 : 0;
    const std::vector<SlabPlan> plans = plan_slabs(src.off, src.n, halo_max, n_dev, t->slab_bytes);
    ACB200_STATS_t sum{};
    sum.devices = (uint32_t)n_dev;
    if (end_state) *end_state = (init_state == ROOT_STATE) ? t->flat.root : init_state;
    if (plans.empty()) { t->stats = sum; return 0; }

    ShardRun run;
    run.res.resize(plans.size());
    std::vector<std::vector<size_t>> mine(n_dev);
    for (size_t i = 0; i < plans.size(); ++i) mine[plans[i].device_slot].push_back(i);
    // (the pools are made here, on the calling thread: the workers only use them)
    std::vector<HelperPool *> pools(n_dev, nullptr);
    if (!src.pinned) { const int thr = copy_threads(n_dev); for (int d = 0; d < n_dev; ++d) pools[d] = &pool_for(t, (size_t)d, thr); }
    std::vector<std::thread> workers;
    for (int d = 0; d < n_dev; ++d)
        workers.emplace_back(shard_worker, (const ac_trie *)t, engines[d], std::cref(src), std::cref(plans), std::cref(mine[d]), first_only,
                             init_state, pools[d], std::ref(run));

    int rc = 0;
    size_t stopped = (size_t)-1;
    std::vector<ACB200_STATS_t> per_dev(n_dev);
    for (size_t i = 0; i < plans.size() && rc == 0; ++i) {
        {
            std::unique_lock<std::mutex> g(run.m);
            run.cv.wait(g, [&] { return run.res[i].ready; });
        }
        SlabResult &r = run.res[i];
        if (!r.ok) { set_error(r.err); rc = -1; break; }
        const SlabPlan &p = plans[i];
        add_stats(per_dev[p.device_slot], r.st);
        if (end_state) *end_state = r.end_state;
        for (const SlabEvent &e : r.ev) {
            const size_t h = e.text_idx;
            if (h == stopped) continue;
            const AC_PATTERN_t *pats;
            const uint32_t size = (uint32_t)patterns_of(t, e.state, &pats);
            const int s = sink(ResolvedEvent{e.end, pats, size, e.state, h});
            if (s && stop_all) { rc = 1; break; }
            if (s || first_only) stopped = h;
        }
        std::vector<SlabEvent>().swap(r.ev);
    }
    if (rc != 0) run.abort.store(true);
    for (auto &w : workers) w.join();
    cudaSetDevice(t->device);
    // devices work side by side: the call's device time is that of the slowest one
    for (int d = 0; d < n_dev; ++d) {
        const ACB200_STATS_t &s = per_dev[d];
        sum.bytes += s.bytes; sum.events += s.events; sum.kernel_launches += s.kernel_launches;
        sum.flagged_words += s.flagged_words; sum.dense_tiles += s.dense_tiles;
        sum.kernel_ms = std::max(sum.kernel_ms, s.kernel_ms); sum.filter_ms = std::max(sum.filter_ms, s.filter_ms);
        sum.verify_ms = std::max(sum.verify_ms, s.verify_ms); sum.reorder_ms = std::max(sum.reorder_ms, s.reorder_ms);
        sum.h2d_ms = std::max(sum.h2d_ms, s.h2d_ms); sum.d2h_ms = std::max(sum.d2h_ms, s.d2h_ms);
        if (s.bytes) { sum.chunk_bytes = s.chunk_bytes; sum.halo_bytes = s.halo_bytes; sum.filtered = s.filtered; }
    }
    t->stats = sum;
    return rc;
}

// Inputs of up to half a slab (32 MiB by default) skip the slab machinery: one copy, one launch, no threads.
inline bool takes_direct_path(const ac_trie *t, uint64_t total)
{
    return total <= t->slab_bytes / 2;
}

} // namespace

// Events of a direct (single launch) scan -> sink, in (text_idx, position) order
template <class Sink>
static void replay_direct(ac_trie *t, const uint64_t *offsets, size_t n, int first_only, Sink &&sink)
{
    const PackedEvent *ev = t->engine.host_events();
    const size_t ne = t->engine.n_events();
    size_t h = 0;
    size_t stopped = (size_t)-1;                 // haystack whose callback asked to stop
    for (size_t i = 0; i < ne; ++i) {
        const uint64_t end = ev[i].end;
        while (h < n && end > offsets[h + 1]) ++h;
        if (h == stopped) continue;
        const AC_PATTERN_t *pats;
        const uint32_t size = (uint32_t)patterns_of(t, ev[i].state, &pats);
        const int r = sink(ResolvedEvent{end - offsets[h], pats, size, ev[i].state, h});
        if (r || first_only) stopped = h;
    }
}

// The direct path (one launch) for a source that is not page-locked — a PHP string, or the strings of a batch:
// the handle's helper threads copy it into pinned staging piece by piece (non-temporal stores) and queue each piece's
// DMA as soon as it is there, so the copy engine runs behind the cores instead of after them.  (cudaMemcpyAsync on a
// pageable pointer stages through the driver on ONE thread: 19 GB/s on the box this was measured on — a 16 MiB
// haystack spent 0.83 ms of its 0.95 ms call there; the pinned copy takes 55 GB/s.)
// Below 3 MiB waking the helpers costs more than it brings (profiles/r02_hostcall_probe.txt, one text, staged against
// the driver's pageable copy: 2 MiB 227 / 206 us, 4 MiB 278 / 312, 8 MiB 385 / 525, 16 MiB 586 / 948, 32 MiB 976 / 1,832 us;
// the strings of a batch, until now gathered by the calling thread alone: 32 MiB 997 / 2,383 us).
static uint64_t staged_min_bytes()
{
    if (const char *e = getenv("ACB200_STAGE_MIN")) { const long long v = atoll(e); if (v > 0) return (uint64_t)v; }
    return 3ull << 20;
}

static bool staged_direct_scan(ac_trie *t, const HaySource &src, uint64_t total, bool first_only, uint32_t init_state)
{
    Engine &eng = t->engine;
    char *stage = eng.slab_staging(0, (size_t)total);
    if (!stage || !eng.slab_upload_begin(0, (size_t)total)) return false;
    // two pieces per thread: every piece costs a cudaMemcpyAsync call (~3 us, serialised inside the driver), and the copy
    // engine should start well before the last core is done
    const int threads = copy_threads(1);
    const uint64_t piece = std::max<uint64_t>(64u << 10, (total / (2u * (uint64_t)threads) + 4095) & ~(uint64_t)4095);
    const int n_pieces = (int)((total + piece - 1) / piece);
    std::mutex em;
    std::string err;
    bool ok = true;
    pool_for(t, 0, threads).run(n_pieces, [&](int i) {
        const uint64_t pb = (uint64_t)i * piece, pe = std::min(total, pb + piece);
        src.copy(stage + pb, pb, pe, src.flat ? 0 : haystack_at(src, pb));
        if (!eng.slab_upload_part(0, (size_t)pb, (size_t)(pe - pb))) {
            std::lock_guard<std::mutex> g(em);
            ok = false; err = get_error();            // (errors are per thread)
        }
    });
    if (!ok) { set_error(err); return false; }
    if (!eng.slab_upload_end(0)) return false;
    if (!eng.scan_slab(0, src.off, src.n, first_only, init_state)) return false;
    eng.stats.h2d_ms = eng.slab_h2d_ms(0);           // staging and DMA together
    return true;
}

template <class Sink>
static int search_source(ac_trie *t, HaySource &src, int first_only, Sink &&sink)
{
    if (t->open) { set_error("automaton is not finalized"); return -1; }
    if (!t->device_ok) return -1;
    if (const char *e = getenv("ACB200_GATHER_NT")) src.stream_stores = atoi(e) != 0;
    const uint64_t total = src.off[src.n];
    if (takes_direct_path(t, total)) {
        const char *bytes = src.flat;
        if (total >= staged_min_bytes() && !(src.flat && is_pinned(src.flat))) {
            if (!staged_direct_scan(t, src, total, first_only != 0, ROOT_STATE)) return -1;
        } else {
            if (!bytes && total) {               // a few scattered haystacks: gathered straight into pinned staging
                char *stage = t->engine.slab_staging(0, (size_t)total);
                if (!stage) return -1;
                src.copy(stage, 0, total, 0);
                bytes = stage;
            }
            if (!t->engine.scan_host(bytes, src.off, src.n, first_only != 0, ROOT_STATE)) return -1;
        }
        t->stats = t->engine.stats; t->stats.devices = 1;
        replay_direct(t, src.off, src.n, first_only, sink);
        return 0;
    }
    src.pinned = src.flat && is_pinned(src.flat);
    return sharded_search(t, src, first_only != 0, ROOT_STATE, false, nullptr, sink) < 0 ? -1 : 0;
}

// ------------------------------------------------------------------------------------------------------------
// Mailbox gather (one process per GPU): see include/acb200.h.  One call per step — enqueue the scan with this
// rank's send buffer as its output, wait for this rank's own count (what the one-GPU call waits for too), then put
// the rows and the mailbox word on their way on a copy stream.  Nothing of it runs in the caller's language: the
// first version drove the same CUDA calls from Python and lost 0.1 ms per step to interpreter time between them.
// ------------------------------------------------------------------------------------------------------------
struct acb200_mailbox {
    ac_trie *t = nullptr;
    int rank = 0, world = 1, device = 0;
    bool collector = false;
    size_t cap = 0;
    char *rows = nullptr;              // the collector's row buffer as mapped in THIS process: [2][world][cap] x 8 bytes
    uint32_t *mbox = nullptr;          // the collector's mailboxes: [2][world][4] words, then one acknowledgement word per parity
    uint2 *send[2] = {nullptr, nullptr};   // this rank's rows of the step in flight: row 0 = {count, dense tiles}, rows 1.. events
    uint32_t *pinned = nullptr;        // [4][4] mailbox / acknowledgement sources, [2] own {count, dense}, then [2][world][4] arrived mailboxes
    cudaStream_t side = nullptr;
    cudaEvent_t copy_done[2] = {nullptr, nullptr}, arrived[2] = {nullptr, nullptr};
    uint32_t step = 0;
    bool pending = false;              // the newest step's rows are complete in send[] but not sent yet (mailbox_flush)
    uint32_t pending_step = 0, pending_count = 0, pending_dense = 0;
    static constexpr uint32_t MBOX_WORDS = 4;

    uint32_t *own_count() { return pinned + 16; }
    uint32_t *arrived_mbox(int p) { return pinned + 24 + (size_t)p * world * MBOX_WORDS; }
};

#define MB_OK(call)                                                                                   \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) { set_error(std::string(#call) + ": " + cudaGetErrorString(e_)); return -1; } \
    } while (0)

// Puts the rows of step k (already complete in send[k & 1]) and its mailbox word on their way.
static int mailbox_send(acb200_mailbox *m, uint32_t k, uint32_t count, uint32_t dense)
{
    const int p = (int)(k & 1u);
    const size_t slot = (size_t)p * m->world + m->rank;
    uint32_t *ack = m->mbox + 2u * m->world * acb200_mailbox::MBOX_WORDS + p;
    uint32_t *src = m->pinned + (k & 3u) * 4u;
    if (k >= 2) {
        // flow control: slot p still holds step k-2 until the collector has let go of it; it acknowledges here,
        // every sender waits for the word (in the collector's memory) on its copy stream
        if (m->collector) {
            src[3] = k - 1u;
            MB_OK(cudaMemcpyAsync(ack, src + 3, 4, cudaMemcpyDefault, m->side));
        } else if (!mailbox_wait_async(m->device, ack, 1, 1, k - 1u, m->side)) return -1;
    }
    if (count) MB_OK(cudaMemcpyAsync(m->rows + slot * m->cap * 8, m->send[p] + 1, (size_t)count * 8, cudaMemcpyDefault, m->side));
    src[0] = k + 1u; src[1] = count; src[2] = dense;
    MB_OK(cudaMemcpyAsync(m->mbox + slot * acb200_mailbox::MBOX_WORDS, src, 12, cudaMemcpyDefault, m->side));
    MB_OK(cudaEventRecord(m->copy_done[p], m->side));
    if (m->collector) {
        uint32_t *boxes = m->mbox + (size_t)p * m->world * acb200_mailbox::MBOX_WORDS;
        if (!mailbox_wait_async(m->device, boxes, (uint32_t)m->world, acb200_mailbox::MBOX_WORDS, k + 1u, m->side)) return -1;
        MB_OK(cudaMemcpyAsync(m->arrived_mbox(p), boxes, (size_t)m->world * acb200_mailbox::MBOX_WORDS * 4, cudaMemcpyDefault, m->side));
        MB_OK(cudaEventRecord(m->arrived[p], m->side));
    }
    return 0;
}

static int mailbox_flush(acb200_mailbox *m)
{
    if (!m->pending) return 0;
    m->pending = false;
    return mailbox_send(m, m->pending_step, m->pending_count, m->pending_dense);
}

static long mailbox_step(acb200_mailbox *m, const void *d_bytes, size_t n, size_t hay_len, void *stream)
{
    ac_trie *t = m->t;
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : cudaStreamLegacy;
    MB_OK(cudaSetDevice(m->device));
    const uint32_t k = m->step++;
    const int p = (int)(k & 1u);
    uint2 *send = m->send[p];
    if (k >= 2) MB_OK(cudaStreamWaitEvent(st, m->copy_done[p], 0));           // the copy of step k-2 has left this buffer
    uint32_t count = 0, dense = 0;
    if (t->engine.scan_device_uniform_async(d_bytes, n, hay_len, send, m->cap, (void *)st)) {
        // the kernels of step k are enqueued: the rows of step k-1 are sent off while the GPU works on them (the dozen
        // driver calls this takes would otherwise sit between two steps, with the GPU idle)
        if (mailbox_flush(m) != 0) return -1;
        MB_OK(cudaStreamSynchronize(st));                                      // this rank's own wait, as on one GPU
        count = t->engine.host_counters()[1]; dense = t->engine.host_counters()[4];     // (left in pinned memory by the emit kernel)
        t->engine.async_finish(count, dense);
    } else {                                                                   // this batch needs the synchronous call (full walk)
        if (mailbox_flush(m) != 0) return -1;
        if (!t->engine.scan_device_uniform(d_bytes, n, hay_len, false, (void *)st)) return -1;
        count = (uint32_t)t->engine.n_events();
        if (count <= m->cap && count && !t->engine.copy_events_to(send + 1, count, (void *)st)) return -1;
        MB_OK(cudaStreamSynchronize(st));
    }
    t->stats = t->engine.stats; t->stats.devices = 1;
    if (count > m->cap) {
        set_error("mailbox gather: " + std::to_string(count) + " events in one step, sized for " + std::to_string(m->cap) + " rows per rank");
        return -1;
    }
    m->pending = true; m->pending_step = k; m->pending_count = count; m->pending_dense = dense;
    return (long)count;
}

extern "C" {

const char *acb200_version(void) { return "acb200 0.1 (sm_100a)"; }
const char *acb200_last_error(void) { return get_error(); }

int acb200_set_device(int device) { set_preferred_device(device); return 0; }

int acb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

void *acb200_device_alloc(int device, size_t bytes) { return device_alloc(device, bytes); }
int acb200_device_free(int device, void *p) { return device_free(device, p) ? 0 : -1; }
int acb200_ipc_export(const void *dptr, unsigned char handle[64]) { return ipc_export(dptr, handle) ? 0 : -1; }
void *acb200_ipc_open(int device, const unsigned char handle[64]) { return ipc_open(device, handle); }
int acb200_ipc_close(int device, void *p) { return ipc_close(device, p) ? 0 : -1; }
int acb200_copy_async(void *dst, const void *src, size_t bytes, void *stream) { return copy_async(dst, src, bytes, stream) ? 0 : -1; }
int acb200_mailbox_wait_async(int device, const void *mailboxes, uint32_t n, uint32_t stride_words, uint32_t seq, void *stream)
{
    return mailbox_wait_async(device, mailboxes, n, stride_words, seq, stream) ? 0 : -1;
}

void *acb200_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        set_error("cudaMallocHost failed");
        return nullptr;
    }
    return p;
}

void acb200_host_free(void *p) { if (p) cudaFreeHost(p); }

AC_TRIE_t *ac_trie_create(void)
{
    ac_trie *t = new (std::nothrow) ac_trie();
    if (t) t->device = preferred_device();       // the device is a property of the handle, not of whoever finalizes it
    return t;
}

AC_STATUS_t ac_trie_add(AC_TRIE_t *t, AC_PATTERN_t *patt, int copy)
{
    if (!t->open) return ACERR_TRIE_CLOSED;      // src/multifast/ahocorasick.c:98-99
    return t->trie.add(patt, copy);
}

void ac_trie_finalize(AC_TRIE_t *t)
{
    if (!t->open) return;
    t->trie.flatten(t->flat);
    t->open = false;                             // src/multifast/ahocorasick.c:154
    set_error("");
    t->device_ok = t->engine.build(t->flat, t->device);
    t->engine.info.n_patterns = t->trie.n_patterns();
    t->engine.info.finalized = 1;
    t->trie.release_build_memory();
    if (t->device_ok) replicate_from_env(t);
    // the expansion inputs (a few words per state) stay: acb200_save() writes them
}

int acb200_save(const AC_TRIE_t *t, const char *path)
{
    if (t->open) { set_error("automaton is not finalized"); return -1; }
    std::string err;
    if (!save_flat(t->flat, path, err)) { set_error(err); return -1; }
    return 0;
}

AC_TRIE_t *acb200_load(const char *path)
{
    ac_trie *t = new (std::nothrow) ac_trie();
    if (!t) { set_error("out of memory"); return nullptr; }
    t->device = preferred_device();
    std::string err;
    if (!load_flat(t->flat, t->blob_arena, path, err)) { set_error(err); delete t; return nullptr; }
    build_gram_table(t->flat);
    t->open = false;
    set_error("");
    t->device_ok = t->engine.build(t->flat, t->device);
    t->engine.info.n_patterns = t->flat.accepted.size();
    t->engine.info.finalized = 1;
    if (t->device_ok) replicate_from_env(t);
    return t;
}

int ac_trie_search(AC_TRIE_t *t, AC_TEXT_t *text, int keep, AC_MATCH_CALBACK_f callback, void *user)
{
    if (t->open) return -1;                      // src/multifast/ahocorasick.c:183-184
    if (!t->device_ok) return -1;
    if (!keep) { t->last_state = ROOT_STATE; t->base_position = 0; }   // ac_trie_reset, ahocorasick.c:330-335
    const uint64_t offs[2] = {0, (uint64_t)text->length};
    const size_t base = t->base_position;
    auto fire = [&](uint64_t position, uint32_t state) -> int {
        const AC_PATTERN_t *pats;
        AC_MATCH_t m;
        m.size = patterns_of(t, state, &pats);
        m.patterns = const_cast<AC_PATTERN_t *>(pats);
        m.position = (size_t)position + base;
        return callback(&m, user);
    };
    uint32_t end_state;
    if (takes_direct_path(t, offs[1])) {
        if (offs[1] >= staged_min_bytes() && !is_pinned(text->astring)) {
            HaySource src;
            src.flat = text->astring; src.off = offs; src.n = 1;
            if (const char *e = getenv("ACB200_GATHER_NT")) src.stream_stores = atoi(e) != 0;
            if (!staged_direct_scan(t, src, offs[1], false, t->last_state)) return -1;
        } else if (!t->engine.scan_host(text->astring, offs, 1, false, t->last_state)) return -1;
        t->stats = t->engine.stats; t->stats.devices = 1;
        const PackedEvent *ev = t->engine.host_events();
        const size_t n = t->engine.n_events();
        for (size_t i = 0; i < n; ++i)
            if (fire(ev[i].end, ev[i].state)) return 1;      // state is not saved on a stop (ahocorasick.c:226-232)
        end_state = t->engine.end_state();
    } else {
        // a long text: slabs with a halo, over every GPU of the handle; a text of any size_t length is scanned
        HaySource src;
        src.flat = text->astring; src.off = offs; src.n = 1; src.pinned = is_pinned(text->astring);
        const int rc = sharded_search(t, src, false, t->last_state, true, &end_state, [&](const ResolvedEvent &e) {
            AC_MATCH_t m;
            m.patterns = const_cast<AC_PATTERN_t *>(e.patterns);
            m.size = e.size;
            m.position = (size_t)e.end + base;
            return callback(&m, user);
        });
        if (rc != 0) return rc;
    }
    t->last_state = end_state;                   // ahocorasick.c:236-238
    t->base_position += text->length;
    return 0;
}

int ac_trie_search_flat(AC_TRIE_t *t, const char *bytes, const uint64_t *offsets, size_t n,
                        int first_only, ACB200_BATCH_CALLBACK_f callback, void *user)
{
    if (offsets[0] != 0) { set_error("offsets[0] must be 0"); return -1; }
    HaySource src;
    src.flat = bytes ? bytes : ""; src.off = offsets; src.n = n;
    return search_source(t, src, first_only, [&](const ResolvedEvent &e) {
        AC_MATCH_t m;
        m.patterns = const_cast<AC_PATTERN_t *>(e.patterns);
        m.size = e.size;
        m.position = (size_t)e.end;
        return callback(e.text_idx, &m, user);
    });
}

int ac_trie_search_batch(AC_TRIE_t *t, const AC_TEXT_t *texts, size_t n, int first_only,
                         ACB200_BATCH_CALLBACK_f callback, void *user)
{
    t->gather_off.resize(n + 1);
    uint64_t total = 0;
    for (size_t i = 0; i < n; ++i) { t->gather_off[i] = total; total += texts[i].length; }
    t->gather_off[n] = total;
    HaySource src;
    src.texts = texts; src.off = t->gather_off.data(); src.n = n;
    return search_source(t, src, first_only, [&](const ResolvedEvent &e) {
        AC_MATCH_t m;
        m.patterns = const_cast<AC_PATTERN_t *>(e.patterns);
        m.size = e.size;
        m.position = (size_t)e.end;
        return callback(e.text_idx, &m, user);
    });
}

int acb200_search_events(AC_TRIE_t *t, const char *bytes, const uint64_t *offsets, size_t n,
                         int first_only, ACB200_EVENT_t *events, size_t cap, size_t *n_events)
{
    if (offsets[0] != 0) { set_error("offsets[0] must be 0"); return -1; }
    HaySource src;
    src.flat = bytes ? bytes : ""; src.off = offsets; src.n = n;
    size_t w = 0;
    const int rc = search_source(t, src, first_only, [&](const ResolvedEvent &e) {
        if (w < cap) {
            events[w].end = e.end;
            events[w].state = e.state;
            events[w].text_idx = (uint32_t)e.text_idx;
        }
        ++w;
        return 0;
    });
    if (n_events) *n_events = w;
    return rc;
}

int acb200_set_devices(AC_TRIE_t *t, const int *devices, size_t n) { return set_devices(t, devices, n); }

int acb200_set_slab_bytes(AC_TRIE_t *t, uint64_t bytes)
{
    if (bytes && (bytes < 4096 || bytes > (1ull << 31))) { set_error("slab size must lie in [4 KiB, 2 GiB]"); return -1; }
    t->slab_bytes = bytes ? bytes : (64ull << 20);
    return 0;
}

int acb200_plan_slabs(const uint64_t *offsets, size_t n, uint32_t halo_max, int n_devices, uint64_t slab_bytes,
                      ACB200_SLAB_t *out, size_t cap, size_t *n_slabs)
{
    const std::vector<SlabPlan> plans = plan_slabs(offsets, n, halo_max, n_devices, slab_bytes ? slab_bytes : (64ull << 20));
    for (size_t i = 0; i < plans.size() && i < cap; ++i) {
        out[i].begin = plans[i].b0; out[i].end = plans[i].b1; out[i].halo = plans[i].halo;
        out[i].device_slot = (uint32_t)plans[i].device_slot;
        out[i].first_text = plans[i].h_first; out[i].end_text = plans[i].h_end;
    }
    if (n_slabs) *n_slabs = plans.size();
    return 0;
}

int acb200_search_hits(AC_TRIE_t *t, const char *bytes, const uint64_t *offsets, size_t n,
                       ACB200_HIT_t *hits, size_t cap, size_t *n_hits)
{
    if (t->open) { set_error("automaton is not finalized"); return -1; }
    if (!t->device_ok) return -1;
    if (offsets[0] != 0) { set_error("offsets[0] must be 0"); return -1; }
    if (!t->engine.scan_host(bytes, offsets, n, false, ROOT_STATE, true)) return -1;
    size_t total = 0;
    const bool ok = t->engine.expand_hits_to_host(n, hits, cap, &total);
    t->stats = t->engine.stats; t->stats.devices = 1;
    if (!ok) return -1;
    if (n_hits) *n_hits = total;
    return 0;
}

int acb200_last_hits(AC_TRIE_t *t, ACB200_HIT_t *hits, size_t cap, size_t *n_hits)
{
    if (t->open || !t->device_ok) { set_error("automaton is not finalized"); return -1; }
    size_t total = 0;
    if (!t->engine.expand_hits_to_host(1, hits, cap, &total)) return -1;
    if (n_hits) *n_hits = total;
    return 0;
}

const AC_PATTERN_t *acb200_pattern(const AC_TRIE_t *t, size_t index)
{
    if (t->open || index >= t->flat.accepted.size()) return nullptr;
    return &t->flat.accepted[index];
}

int acb200_search_device(AC_TRIE_t *t, const void *d_bytes, const uint64_t *offsets, size_t n,
                         int first_only, void *stream, const void **d_events, size_t *n_events)
{
    if (t->open) { set_error("automaton is not finalized"); return -1; }
    if (!t->device_ok) return -1;
    if (offsets[0] != 0) { set_error("offsets[0] must be 0"); return -1; }
    const bool ok = t->engine.scan_device(d_bytes, offsets, n, first_only != 0, ROOT_STATE, stream);
    t->stats = t->engine.stats; t->stats.devices = 1;
    if (!ok) return -1;
    if (d_events) *d_events = t->engine.device_events();
    if (n_events) *n_events = t->engine.n_events();
    return 0;
}

int acb200_search_device_uniform(AC_TRIE_t *t, const void *d_bytes, size_t n, size_t hay_len,
                                 int first_only, void *stream, const void **d_events, size_t *n_events)
{
    if (t->open) { set_error("automaton is not finalized"); return -1; }
    if (!t->device_ok) return -1;
    const bool ok = t->engine.scan_device_uniform(d_bytes, n, hay_len, first_only != 0, stream);
    t->stats = t->engine.stats; t->stats.devices = 1;
    if (!ok) return -1;
    if (d_events) *d_events = t->engine.device_events();
    if (n_events) *n_events = t->engine.n_events();
    return 0;
}

int acb200_search_device_uniform_async(AC_TRIE_t *t, const void *d_bytes, size_t n, size_t hay_len, void *d_rows,
                                       size_t max_events, void *stream)
{
    if (t->open) { set_error("automaton is not finalized"); return -1; }
    if (!t->device_ok) return -1;
    return t->engine.scan_device_uniform_async(d_bytes, n, hay_len, d_rows, max_events, stream) ? 0 : -1;
}

int acb200_async_finish(AC_TRIE_t *t, size_t n_events, size_t dense_tiles)
{
    if (t->open || !t->device_ok) return -1;
    t->engine.async_finish(n_events, dense_tiles);
    t->stats = t->engine.stats; t->stats.devices = 1;
    return 0;
}

long acb200_copy_events(AC_TRIE_t *t, void *d_dst, size_t max_events, void *stream)
{
    if (t->open || !t->device_ok) { set_error("automaton is not finalized"); return -1; }
    const size_t n = std::min(max_events, t->engine.n_events());
    if (n && !t->engine.copy_events_to(d_dst, n, stream)) return -1;
    return (long)n;
}

static inline uint64_t tally_fold(uint64_t h, uint64_t v)
{
    h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
    return h * 0xff51afd7ed558ccdULL;
}

int acb200_tally_cb(size_t text_idx, AC_MATCH_t *m, void *tally)
{
    // order-sensitive, one multiply on the dependent chain per event (a consumer that costs more than the replay
    // loop around it would be what an end-to-end number measures)
    ACB200_TALLY_t *t = static_cast<ACB200_TALLY_t *>(tally);
    t->events++;
    t->hits += m->size;
    uint64_t v = ((uint64_t)m->position * 0x9e3779b97f4a7c15ULL) ^ ((uint64_t)text_idx << 20) ^ ((uint64_t)m->size << 52);
    if (m->size)
        v ^= (uint64_t)(uintptr_t)m->patterns[0].aux * 0xc2b2ae3d27d4eb4fULL + (uint64_t)(uintptr_t)m->patterns[m->size - 1].aux;
    t->hash = (((t->hash << 7) | (t->hash >> 57)) ^ v) * 0xff51afd7ed558ccdULL;
    return 0;
}

int acb200_tally_match_cb(AC_MATCH_t *m, void *tally) { return acb200_tally_cb(0, m, tally); }

int acb200_event_digest(const AC_TRIE_t *t, const ACB200_EVENT_t *events, size_t n_events, size_t n_texts,
                        uint64_t *counts, uint64_t *hashes)
{
    if (t->open) { set_error("automaton is not finalized"); return -1; }
    for (size_t h = 0; h < n_texts; ++h) { counts[h] = 0; hashes[h] = 0; }
    for (size_t i = 0; i < n_events; ++i) {
        const ACB200_EVENT_t &e = events[i];
        if (e.text_idx >= n_texts) { set_error("event names a haystack outside the batch"); return -1; }
        const AC_PATTERN_t *pats;
        const size_t size = patterns_of(t, e.state, &pats);
        uint64_t h = tally_fold(hashes[e.text_idx], e.end);
        h = tally_fold(h, (uint64_t)size);
        if (size) {
            h = tally_fold(h, (uint64_t)(uintptr_t)pats[0].aux);
            h = tally_fold(h, (uint64_t)(uintptr_t)pats[size - 1].aux);
        }
        hashes[e.text_idx] = h;
        counts[e.text_idx]++;
    }
    return 0;
}

size_t acb200_state_patterns(const AC_TRIE_t *t, uint32_t state, const AC_PATTERN_t **patterns)
{
    if (t->open) { if (patterns) *patterns = nullptr; return 0; }
    return patterns_of(t, state, patterns);
}

int acb200_info(const AC_TRIE_t *t, ACB200_INFO_t *out)
{
    if (!out) return -1;
    *out = t->engine.info;
    out->n_patterns = t->open ? t->trie.n_patterns() : t->flat.accepted.size();
    out->finalized = t->open ? 0 : 1;
    if (t->open || !t->device_ok) {
        out->n_states = t->open ? t->trie.n_nodes() : t->flat.n_states;
        out->device = -1;
    }
    if (!t->open) {                              // host facts of a finalized automaton hold with or without a device
        out->n_classes = t->flat.n_classes;
        out->max_pattern_len = t->flat.max_pattern_len;
        out->final_bound = t->flat.final_bound;
        out->root = t->flat.root;
        out->filter_word = (int32_t)t->flat.filter_w;
        out->min_pattern_len = t->flat.min_pattern_len;
        out->filter_l1_fill = (float)t->flat.l1_fill;
        out->filter_l2_log2 = t->flat.l2_log2;
        out->direct_keys = (uint32_t)t->flat.gt_keys;
        out->direct_walk_keys = (uint32_t)t->flat.gt_walk_keys;
    }
    return 0;
}

int acb200_last_stats(const AC_TRIE_t *t, ACB200_STATS_t *out)
{
    if (!out) return -1;
    *out = t->stats;
    return 0;
}

int acb200_set_tuning(AC_TRIE_t *t, uint32_t chunk_bytes, uint32_t smem_table_bytes)
{
    t->engine.tune_chunk = chunk_bytes;
    t->engine.tune_smem_bytes = smem_table_bytes;
    return 0;
}

int acb200_set_tma(AC_TRIE_t *t, int mode)
{
    t->engine.tune_tma = mode;
    return 0;
}

int acb200_set_filter(AC_TRIE_t *t, int mode)
{
    t->engine.tune_filter = mode;
    return 0;
}

int acb200_filter_probe(const AC_TRIE_t *t, uint64_t word, unsigned next_byte)
{
    const FlatAutomaton &f = t->flat;
    if (t->open || f.filter_w == 0) return -1;
    const uint32_t lo = (uint32_t)word, hi = (f.filter_w == 8) ? (uint32_t)(word >> 32) : 0u;
    const uint32_t tt = filter_mix1(lo, hi, next_byte);
    const uint32_t w = f.l1[filter_l1_word(tt, next_byte == FILTER_NEXT_UNKNOWN)];
    if (!((w >> filter_bit1(tt)) & (w >> filter_bit2(tt)) & 1u)) return 0;
    if (f.l2_log2) {
        const uint32_t i3 = filter_l2_index(tt, f.l2_log2);
        if (!((f.l2[i3 >> 5] >> (i3 & 31)) & 1u)) return 0;
    }
    return 1;
}

int acb200_set_direct(AC_TRIE_t *t, int mode)
{
    t->engine.tune_direct = mode;
    return 0;
}

int acb200_direct_probe(const AC_TRIE_t *t, const char *bytes, size_t length, size_t hay_begin, size_t word_index, uint32_t *end, uint32_t *state)
{
    const FlatAutomaton &f = t->flat;
    if (t->open || f.filter_w == 0 || f.gt_log2 == 0) return -1;
    const uint32_t W = f.filter_w;
    const uint32_t warm = (f.max_pattern_len - 1 + W - 1) / W * W;
    const uint64_t rs = (uint64_t)(word_index + 1) * W;
    if (rs < warm || rs + W > length || rs + W > 0xffffffffull) return -1;     // window clipped by the stream: the kernel walks
    const uint32_t hb = (uint32_t)std::min<uint64_t>(hay_begin, rs);
    const uint8_t *text = (const uint8_t *)bytes;
    auto slot = [&](uint32_t i) { return f.gt_slots[i]; };
    auto st = [&](uint32_t i) { return f.gt_pat[i]; };
    uint32_t e = 0, s = 0;
    int v;
    if (W == 8) {
        auto txt = [&](uint32_t i) { uint64_t c; memcpy(&c, text + i, 8); return c; };
        auto pat = [&](uint32_t i) { uint64_t c; memcpy(&c, f.gt_pat.data() + i, 8); return c; };
        v = (int)gram_verify<8>((uint32_t)rs, warm, hb, f.gt_log2, txt, slot, pat, st, &e, &s);
    } else {
        auto txt = [&](uint32_t i) { uint32_t c; memcpy(&c, text + i, 4); return c; };
        auto pat = [&](uint32_t i) { return f.gt_pat[i]; };
        v = (int)gram_verify<4>((uint32_t)rs, warm, hb, f.gt_log2, txt, slot, pat, st, &e, &s);
    }
    if (v == GRAM_EVENT) { if (end) *end = e; if (state) *state = s; }
    return v;
}

ACB200_MAILBOX_t *acb200_mailbox_create(AC_TRIE_t *t, int rank, int world, int collector, size_t cap_rows, void *rows, void *mailboxes)
{
    if (t->open || !t->device_ok) { set_error("automaton is not finalized"); return nullptr; }
    acb200_mailbox *m = new (std::nothrow) acb200_mailbox();
    if (!m) { set_error("out of memory"); return nullptr; }
    m->t = t; m->rank = rank; m->world = world; m->collector = collector != 0; m->cap = cap_rows;
    m->device = t->engine.device();
    m->rows = static_cast<char *>(rows); m->mbox = static_cast<uint32_t *>(mailboxes);
    bool ok = cudaSetDevice(m->device) == cudaSuccess;
    for (int p = 0; ok && p < 2; ++p) {
        ok = cudaMalloc((void **)&m->send[p], (cap_rows + 1) * 8) == cudaSuccess && cudaMemset(m->send[p], 0, (cap_rows + 1) * 8) == cudaSuccess &&
             cudaEventCreateWithFlags(&m->copy_done[p], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&m->arrived[p], cudaEventDisableTiming) == cudaSuccess;
    }
    ok = ok && cudaDeviceSynchronize() == cudaSuccess;          // the zero-fills ran on the legacy stream
    ok = ok && cudaMallocHost((void **)&m->pinned, (24 + 2 * (size_t)world * acb200_mailbox::MBOX_WORDS) * 4) == cudaSuccess &&
         cudaStreamCreateWithFlags(&m->side, cudaStreamNonBlocking) == cudaSuccess;
    if (!ok) { set_error(std::string("mailbox setup: ") + cudaGetErrorString(cudaGetLastError())); acb200_mailbox_free(m); return nullptr; }
    memset(m->pinned, 0, (24 + 2 * (size_t)world * acb200_mailbox::MBOX_WORDS) * 4);
    return m;
}

long acb200_mailbox_step(ACB200_MAILBOX_t *m, const void *d_bytes, size_t n, size_t hay_len, void *stream)
{
    return mailbox_step(m, d_bytes, n, hay_len, stream);
}

int acb200_mailbox_result(ACB200_MAILBOX_t *m, uint32_t step, uint32_t *counts)
{
    if (!m->collector) { set_error("only the collector holds the rows"); return -1; }
    if (step >= m->step || step + 2 < m->step) { set_error("mailbox gather: that step's rows are not held (any more)"); return -1; }
    const int p = (int)(step & 1u);
    MB_OK(cudaSetDevice(m->device));
    if (m->pending && m->pending_step == step && mailbox_flush(m) != 0) return -1;
    MB_OK(cudaEventSynchronize(m->arrived[p]));
    for (int r = 0; r < m->world; ++r) counts[r] = m->arrived_mbox(p)[(size_t)r * acb200_mailbox::MBOX_WORDS + 1];
    return 0;
}

int acb200_mailbox_drain(ACB200_MAILBOX_t *m, void *stream)
{
    MB_OK(cudaSetDevice(m->device));
    if (mailbox_flush(m) != 0) return -1;
    for (int p = 0; p < 2; ++p) MB_OK(cudaStreamWaitEvent(stream ? static_cast<cudaStream_t>(stream) : cudaStreamLegacy, m->copy_done[p], 0));
    return 0;
}

void acb200_mailbox_free(ACB200_MAILBOX_t *m)
{
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->side) { cudaStreamSynchronize(m->side); cudaStreamDestroy(m->side); }
    for (int p = 0; p < 2; ++p) {
        cudaFree(m->send[p]);
        if (m->copy_done[p]) cudaEventDestroy(m->copy_done[p]);
        if (m->arrived[p]) cudaEventDestroy(m->arrived[p]);
    }
    if (m->pinned) cudaFreeHost(m->pinned);
    delete m;
}

void ac_trie_release(AC_TRIE_t *t)
{
    delete t;
}

} // extern "C"
