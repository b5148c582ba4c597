// cabi.cpp — extern "C" surface declared in include/acb200.h.
//
// Mirrors the five MultiFast entry points the reference's Zend glue calls
// (src/multifast/ahocorasick.h:73-80) and adds the batched / event-level /
// device-resident entry points.  The device does the scan; this file replays
// the compact (position, state) events through the caller's callback exactly
// as the reference's loop would have called it
// (src/multifast/ahocorasick.c:214-233).
#include <algorithm>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <cuda_runtime_api.h>

#include "acb200.h"
#include "automaton.hpp"
#include "engine.hpp"
#include "filter_hash.hpp"

using namespace acb200;

struct ac_trie {
    HostTrie trie;
    FlatAutomaton flat;
    Engine engine;
    bool open = true;          // patterns may still be added (reference: trie_open)
    bool device_ok = false;    // finalize reached the device
    uint32_t last_state = ROOT_STATE;   // keep=1 continuation (reference: last_node)
    size_t base_position = 0;  // keep=1 continuation (reference: base_position)
    std::vector<char> gather;  // batch gather buffer
    std::vector<uint64_t> gather_off;
    std::deque<std::string> blob_arena;   // pattern bytes / string ids of an automaton loaded from a blob
};

static inline size_t patterns_of(const ac_trie *t, uint32_t state, const AC_PATTERN_t **p)
{
    if (state == 0 || state >= t->flat.final_bound) { if (p) *p = nullptr; return 0; }
    const uint32_t i = state - 1;
    const uint64_t b = t->flat.out_off[i], e = t->flat.out_off[i + 1];
    if (p) *p = t->flat.out_pat.data() + b;
    return (size_t)(e - b);
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
    return new (std::nothrow) ac_trie();
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
    t->device_ok = t->engine.build(t->flat);
    t->engine.info.n_patterns = t->trie.n_patterns();
    t->engine.info.finalized = 1;
    t->trie.release_build_memory();
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
    std::string err;
    if (!load_flat(t->flat, t->blob_arena, path, err)) { set_error(err); delete t; return nullptr; }
    build_gram_table(t->flat);
    t->open = false;
    set_error("");
    t->device_ok = t->engine.build(t->flat);
    t->engine.info.n_patterns = t->flat.accepted.size();
    t->engine.info.finalized = 1;
    return t;
}

int ac_trie_search(AC_TRIE_t *t, AC_TEXT_t *text, int keep, AC_MATCH_CALBACK_f callback, void *user)
{
    if (t->open) return -1;                      // src/multifast/ahocorasick.c:183-184
    if (!t->device_ok) return -1;
    if (!keep) { t->last_state = ROOT_STATE; t->base_position = 0; }   // ac_trie_reset, ahocorasick.c:330-335
    const uint64_t offs[2] = {0, (uint64_t)text->length};
    if (!t->engine.scan_host(text->astring, offs, 1, false, t->last_state)) return -1;
    const PackedEvent *ev = t->engine.host_events();
    const size_t n = t->engine.n_events();
    for (size_t i = 0; i < n; ++i) {
        const AC_PATTERN_t *pats;
        AC_MATCH_t m;
        m.size = patterns_of(t, ev[i].state, &pats);
        m.patterns = const_cast<AC_PATTERN_t *>(pats);
        m.position = (size_t)ev[i].end + t->base_position;
        if (callback(&m, user)) return 1;        // state is not saved on a stop (ahocorasick.c:226-232)
    }
    t->last_state = t->engine.end_state();       // ahocorasick.c:236-238
    t->base_position += text->length;
    return 0;
}

static int replay_batch(ac_trie *t, const uint64_t *offsets, size_t n, int first_only,
                        ACB200_BATCH_CALLBACK_f callback, void *user)
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
        AC_MATCH_t m;
        m.size = patterns_of(t, ev[i].state, &pats);
        m.patterns = const_cast<AC_PATTERN_t *>(pats);
        m.position = (size_t)(end - offsets[h]);
        const int r = callback(h, &m, user);
        if (r || first_only) stopped = h;
    }
    return 0;
}

// Large batches are cut into slabs at haystack boundaries and pipelined: while slab i is scanned on the
// device and its events are replayed through the callback on the host, slab i+1 is already crossing PCIe.
static constexpr uint64_t SLAB_BYTES = 64ull << 20;

static int search_flat_pipelined(ac_trie *t, const char *bytes, const uint64_t *offsets, size_t n,
                                 int first_only, ACB200_BATCH_CALLBACK_f callback, void *user)
{
    // slab s covers haystacks [cut[s], cut[s+1])
    std::vector<size_t> cut{0};
    for (size_t h = 0; h < n;) {
        size_t e = h;
        while (e < n && offsets[e + 1] - offsets[h] <= SLAB_BYTES) ++e;
        if (e == h) e = h + 1;                   // a single haystack larger than a slab travels alone
        if (offsets[e] - offsets[h] >= 0xffffff00ull) { set_error("haystack exceeds 4 GiB"); return -1; }
        cut.push_back(e);
        h = e;
    }
    const size_t n_slabs = cut.size() - 1;
    Engine &eng = t->engine;
    ACB200_STATS_t sum{};
    std::vector<uint64_t> rel;
    auto upload = [&](size_t s) {
        const uint64_t b0 = offsets[cut[s]], b1 = offsets[cut[s + 1]];
        return eng.slab_upload_async((int)(s & 1), bytes + b0, (size_t)(b1 - b0));
    };
    if (!upload(0)) return -1;
    for (size_t s = 0; s < n_slabs; ++s) {
        if (s + 1 < n_slabs && !upload(s + 1)) return -1;
        const size_t h0 = cut[s], h1 = cut[s + 1];
        rel.resize(h1 - h0 + 1);
        for (size_t i = 0; i <= h1 - h0; ++i) rel[i] = offsets[h0 + i] - offsets[h0];
        if (!eng.scan_slab((int)(s & 1), rel.data(), h1 - h0, first_only != 0)) return -1;
        sum.bytes += eng.stats.bytes; sum.events += eng.stats.events; sum.kernel_launches += eng.stats.kernel_launches;
        sum.kernel_ms += eng.stats.kernel_ms; sum.filter_ms += eng.stats.filter_ms; sum.verify_ms += eng.stats.verify_ms;
        sum.reorder_ms += eng.stats.reorder_ms; sum.d2h_ms += eng.stats.d2h_ms; sum.h2d_ms += eng.slab_h2d_ms((int)(s & 1));
        sum.flagged_words += eng.stats.flagged_words; sum.dense_tiles += eng.stats.dense_tiles;
        sum.chunk_bytes = eng.stats.chunk_bytes; sum.halo_bytes = eng.stats.halo_bytes;
        sum.filtered = eng.stats.filtered;
        // replay this slab's events; haystack indices are those of the whole batch
        const PackedEvent *ev = eng.host_events();
        const size_t ne = eng.n_events();
        size_t h = 0, stopped = (size_t)-1;
        for (size_t i = 0; i < ne; ++i) {
            const uint64_t end = ev[i].end;
            while (h < h1 - h0 && end > rel[h + 1]) ++h;
            if (h == stopped) continue;
            const AC_PATTERN_t *pats;
            AC_MATCH_t m;
            m.size = patterns_of(t, ev[i].state, &pats);
            m.patterns = const_cast<AC_PATTERN_t *>(pats);
            m.position = (size_t)(end - rel[h]);
            const int r = callback(h0 + h, &m, user);
            if (r || first_only) stopped = h;
        }
    }
    eng.stats = sum;
    return 0;
}

int ac_trie_search_flat(AC_TRIE_t *t, const char *bytes, const uint64_t *offsets, size_t n,
                        int first_only, ACB200_BATCH_CALLBACK_f callback, void *user)
{
    if (t->open) { set_error("automaton is not finalized"); return -1; }
    if (!t->device_ok) return -1;
    if (offsets[0] != 0) { set_error("offsets[0] must be 0"); return -1; }
    if (n > 1 && offsets[n] > 2 * SLAB_BYTES) return search_flat_pipelined(t, bytes, offsets, n, first_only, callback, user);
    if (!t->engine.scan_host(bytes, offsets, n, first_only != 0, ROOT_STATE)) return -1;
    return replay_batch(t, offsets, n, first_only, callback, user);
}

int ac_trie_search_batch(AC_TRIE_t *t, const AC_TEXT_t *texts, size_t n, int first_only,
                         ACB200_BATCH_CALLBACK_f callback, void *user)
{
    if (t->open) { set_error("automaton is not finalized"); return -1; }
    if (!t->device_ok) return -1;
    t->gather_off.resize(n + 1);
    uint64_t total = 0;
    for (size_t i = 0; i < n; ++i) { t->gather_off[i] = total; total += texts[i].length; }
    t->gather_off[n] = total;
    if (t->gather.size() < total) t->gather.resize(total);
    for (size_t i = 0; i < n; ++i)
        if (texts[i].length) memcpy(t->gather.data() + t->gather_off[i], texts[i].astring, texts[i].length);
    if (!t->engine.scan_host(t->gather.data(), t->gather_off.data(), n, first_only != 0, ROOT_STATE)) return -1;
    return replay_batch(t, t->gather_off.data(), n, first_only, callback, user);
}

int acb200_search_events(AC_TRIE_t *t, const char *bytes, const uint64_t *offsets, size_t n,
                         int first_only, ACB200_EVENT_t *events, size_t cap, size_t *n_events)
{
    if (t->open) { set_error("automaton is not finalized"); return -1; }
    if (!t->device_ok) return -1;
    if (offsets[0] != 0) { set_error("offsets[0] must be 0"); return -1; }
    if (!t->engine.scan_host(bytes, offsets, n, first_only != 0, ROOT_STATE)) return -1;
    const PackedEvent *ev = t->engine.host_events();
    const size_t ne = t->engine.n_events();
    size_t h = 0, w = 0;
    size_t last_h = (size_t)-1;
    for (size_t i = 0; i < ne; ++i) {
        const uint64_t end = ev[i].end;
        while (h < n && end > offsets[h + 1]) ++h;
        if (first_only) { if (h == last_h) continue; last_h = h; }
        if (w < cap) {
            events[w].end = end - offsets[h];
            events[w].state = ev[i].state;
            events[w].text_idx = (uint32_t)h;
        }
        ++w;
    }
    if (n_events) *n_events = w;
    return 0;
}

int acb200_search_hits(AC_TRIE_t *t, const char *bytes, const uint64_t *offsets, size_t n,
                       ACB200_HIT_t *hits, size_t cap, size_t *n_hits)
{
    if (t->open) { set_error("automaton is not finalized"); return -1; }
    if (!t->device_ok) return -1;
    if (offsets[0] != 0) { set_error("offsets[0] must be 0"); return -1; }
    if (!t->engine.scan_host(bytes, offsets, n, false, ROOT_STATE)) return -1;
    size_t total = 0;
    if (!t->engine.expand_hits_to_host(n, hits, cap, &total)) return -1;
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
    if (!t->engine.scan_device(d_bytes, offsets, n, first_only != 0, ROOT_STATE, stream)) return -1;
    if (d_events) *d_events = t->engine.device_events();
    if (n_events) *n_events = t->engine.n_events();
    return 0;
}

int acb200_search_device_uniform(AC_TRIE_t *t, const void *d_bytes, size_t n, size_t hay_len,
                                 int first_only, void *stream, const void **d_events, size_t *n_events)
{
    if (t->open) { set_error("automaton is not finalized"); return -1; }
    if (!t->device_ok) return -1;
    if (!t->engine.scan_device_uniform(d_bytes, n, hay_len, first_only != 0, stream)) return -1;
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
    ACB200_TALLY_t *t = static_cast<ACB200_TALLY_t *>(tally);
    t->events++;
    t->hits += m->size;
    uint64_t h = tally_fold(t->hash, (uint64_t)text_idx);
    h = tally_fold(h, (uint64_t)m->position);
    h = tally_fold(h, (uint64_t)m->size);
    if (m->size) {
        h = tally_fold(h, (uint64_t)(uintptr_t)m->patterns[0].aux);
        h = tally_fold(h, (uint64_t)(uintptr_t)m->patterns[m->size - 1].aux);
    }
    t->hash = h;
    return 0;
}

int acb200_tally_match_cb(AC_MATCH_t *m, void *tally) { return acb200_tally_cb(0, m, tally); }

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
    *out = t->engine.stats;
    return 0;
}

int acb200_set_tuning(AC_TRIE_t *t, uint32_t chunk_bytes, uint32_t smem_table_bytes)
{
    t->engine.tune_chunk = chunk_bytes;
    t->engine.tune_smem_bytes = smem_table_bytes;
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
        const uint32_t i3 = filter_mix3(lo, hi, next_byte) >> (32 - f.l2_log2);
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

void ac_trie_release(AC_TRIE_t *t)
{
    delete t;
}

} // extern "C"
