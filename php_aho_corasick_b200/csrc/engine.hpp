// engine.hpp — device side of one automaton handle: the dense table in HBM,
// per-handle scratch (haystack staging, event buffer, look-back words) and the
// launch logic of the scan.  Host code only; CUDA types stay behind void*.
#pragma once

#include <cstdint>
#include <cstddef>
#include <string>
#include <vector>

#include "acb200.h"
#include "automaton.hpp"

namespace acb200 {

struct PackedEvent { uint32_t end; uint32_t state; };   // as written by the kernel
constexpr uint32_t ROOT_STATE = 0xffffffffu;            // init_state value meaning "start at the root"
// One launch addresses its stream with 32-bit offsets; the bound leaves room for slice ends (offset + slice
// length, slice <= 1 MiB) to stay below 2^32.  Host callers never see it: their streams are cut into slabs.
constexpr uint64_t MAX_STREAM_BYTES = 0xffe00000ull;

void set_error(const std::string &msg);
const char *get_error();
int preferred_device();
void set_preferred_device(int d);

// Peer-memory plumbing of the multi-GPU event gather (one process per GPU): device buffers that other processes map
// through CUDA IPC, copies into them by the copy engines, and the wait for every sender's "step done" mailbox.
void *device_alloc(int device, size_t bytes);
bool device_free(int device, void *p);
bool ipc_export(const void *dptr, unsigned char handle[64]);
void *ipc_open(int device, const unsigned char handle[64]);
bool ipc_close(int device, void *p);
bool copy_async(void *dst, const void *src, size_t bytes, void *stream);
bool mailbox_wait_async(int device, const void *mailboxes, uint32_t n, uint32_t stride_words, uint32_t seq, void *stream);

class Engine {
public:
    Engine();
    ~Engine();

    // Uploads `flat` to CUDA device `device` and expands the dense table there; a replica (table_src != nullptr)
    // copies the expanded table from the primary's HBM instead (peer copy over NVLink).  false on error.
    bool build(const FlatAutomaton &flat, int device, const Engine *table_src = nullptr);
    int device() const { return device_; }

    // Scans a flat haystack stream that lives in HOST memory.  Events are left in
    // host_events() sorted by stream offset; returns false on error.
    // (events_stay_on_device: the caller goes on with the device copy of the events — expand_hits_to_host())
    bool scan_host(const char *bytes, const uint64_t *offsets, size_t n, bool first_only,
                   uint32_t init_state, bool events_stay_on_device = false);
    // Pipelined variant for large batches: the caller cuts the batch into slabs at haystack boundaries,
    // uploads slab i+1 (slab_upload_async) while slab i is scanned (scan_slab) and replayed on the host.
    // `bytes` is page-locked memory: the caller's own (acb200_host_alloc) or slab_staging(buf, n).
    char *slab_staging(int buf, size_t n_bytes);
    bool slab_upload_async(int buf, const char *bytes, size_t n_bytes);
    // ... or piece by piece while the staging buffer is still being filled: begin, then any number of parts of
    // slab_staging(buf) — slab_upload_part may be called from several threads at once —, then end.
    bool slab_upload_begin(int buf, size_t n_bytes);
    bool slab_upload_part(int buf, size_t offset, size_t n_bytes);
    bool slab_upload_end(int buf);
    bool scan_slab(int buf, const uint64_t *offsets, size_t n, bool first_only, uint32_t init_state = ROOT_STATE);
    float slab_h2d_ms(int buf);
    // Same for a stream already resident in device memory; events stay on the device.
    bool scan_device(const void *d_bytes, const uint64_t *offsets, size_t n, bool first_only,
                     uint32_t init_state, void *stream);

    // n haystacks of equal length laid end to end
    bool scan_device_uniform(const void *d_bytes, size_t n, size_t hay_len, bool first_only, void *stream);

    // Asynchronous variant of scan_device_uniform for callers that chain more device work behind the scan (the
    // multi-GPU event gather): nothing is waited for.  Row 0 of d_rows (8-byte rows) receives the event count,
    // rows 1 .. max_events the first max_events events.  Only the prefilter path can do this (false + error
    // otherwise: use the synchronous call); async_finish() completes the bookkeeping once the caller has waited.
    bool scan_device_uniform_async(const void *d_bytes, size_t n, size_t hay_len, void *d_rows, size_t max_events, void *stream);
    void async_finish(size_t n_events, size_t dense_tiles);
    bool copy_events_to(void *d_dst, size_t n, void *stream);
    // Expands the events of the most recent scan_host() into hits on the device and copies up to `cap` of them
    // to `hits` (host).  *n_hits receives the total.
    bool expand_hits_to_host(size_t n_hay, ACB200_HIT_t *hits, size_t cap, size_t *n_hits);
    const PackedEvent *host_events() const { return last_host_events_; }
    const void *device_events() const { return d_events_; }
    size_t n_events() const { return n_events_; }
    uint32_t end_state() const { return end_state_; }

    ACB200_STATS_t stats{};
    ACB200_INFO_t info{};
    uint32_t tune_chunk = 0;
    uint32_t tune_smem_bytes = 0;
    int tune_tma = 0;              // full walk: 0 auto / 1 haystack text staged by the TMA unit (ac_scan_tma_kernel), -1 plain loads
    int tune_filter = 0;           // 0 auto, 1 always use the gram prefilter when the dictionary allows, -1 never
    int tune_direct = 0;           // 0 auto (= 1), 1 direct verification of flagged words inside the walk kernel,
                                   // -1 every flagged word is walked

    // counters of the last prefilter-path call as its last kernel handed them to pinned host memory: [1] events,
    // [4] densely flagged tiles — valid once the call's stream has been waited for (asynchronous calls included)
    const uint32_t *host_counters() const { return h_counters_; }

private:
    bool ensure_text(size_t bytes);
    bool ensure_events(size_t n);
    bool ensure_offsets(size_t n);
    bool ensure_tiles(size_t n);
    bool ensure_mask(size_t words);
    bool ensure_verify_scratch(size_t n_tiles);
    bool launch_filtered(const void *d_text, uint32_t total, uint32_t readable, size_t n_hay, uint32_t uniform_len,
                         void *stream);
    void window_for(size_t smem_budget, uint32_t *win_lo, uint32_t *win_rows) const;
    bool ensure_host_events(size_t n);
    bool scan_small(const char *bytes, uint32_t total, uint32_t init_state);
    bool upload_offsets(const uint64_t *offsets, size_t n, uint32_t *uniform_len);
    bool launch_scan(const void *d_text, uint32_t total, uint32_t readable, size_t n_hay, uint32_t uniform_len,
                     bool first_only, uint32_t init_state, void *stream);
    uint32_t pick_chunk(uint64_t total) const;
    void release();

    int device_ = -1;
    int n_sms_ = 0;
    int max_smem_optin_ = 0;
    void *stream_ = nullptr;       // cudaStream_t
    void *ev_[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

    // automaton
    void *d_table_ = nullptr;
    uint8_t *d_cls_ = nullptr;
    uint32_t ncls_ = 1, final_bound_ = 1, root_ = 1, n_states_ = 1, n_rows_ = 2, halo_ = 0;
    uint32_t range_lo_ = 0, n_used_ = 0;
    bool range_map_ = true;
    int entry_bytes_ = 2;
    uint64_t table_entries_ = 1;

    // gram prefilter
    uint32_t filter_w_ = 0, l1_bits_ = 0, l2_log2_ = 0;
    // output lists on the device (hit expansion)
    uint32_t *d_out_off_ = nullptr, *d_out_idx_ = nullptr, *d_pat_len_ = nullptr;
    uint32_t *d_hit_sums_ = nullptr; size_t hit_sums_cap_ = 0;
    void *d_hits_ = nullptr; size_t hits_cap_ = 0;
    unsigned long long *d_hit_total_ = nullptr;
    uint32_t last_uniform_len_ = 0;            // batch shape of the most recent scan (hit expansion needs it)
    uint32_t *d_l1_ = nullptr, *d_l2_ = nullptr;
    void *d_gt_slots_ = nullptr;    // exact gram table (gram_table.hpp)
    uint32_t *d_gt_pat_ = nullptr;
    uint32_t gt_log2_ = 0;
    uint32_t *d_mask_ = nullptr;  size_t mask_cap_ = 0;
    uint32_t *d_items_ = nullptr;   // work items of the verify kernels
    uint32_t *d_recs_ = nullptr;    // per item {first event state, count << 16 | relative end}
    uint32_t *d_desc_ = nullptr;    // per 16 KiB tile {offset into items, count}
    uint32_t *d_tile_len_ = nullptr;// per tile: events
    size_t verify_tiles_cap_ = 0;
    double last_dense_frac_ = 0.0; // tiles the verify kernel had to walk completely, previous filtered scan

    // scratch
    uint8_t *d_text_ = nullptr;   size_t text_cap_ = 0;
    uint32_t *d_off_ = nullptr;   size_t off_cap_ = 0;
    uint32_t *d_first_ = nullptr; size_t first_cap_ = 0;
    void *d_events_ = nullptr;    size_t events_cap_ = 0;
    unsigned long long *d_tiles_ = nullptr; size_t tiles_cap_ = 0;
    uint32_t *h_counters_ = nullptr;          // pinned
    PackedEvent *h_events_ = nullptr; size_t h_events_cap_ = 0;   // pinned
    const PackedEvent *last_host_events_ = nullptr;               // where the most recent host-side scan left its events
    uint8_t *h_small_ = nullptr;                                   // pinned + mapped: [64-byte header | events | text staging] of the one-CTA path
    void *h_slab_[2] = {nullptr, nullptr}; size_t h_slab_cap_[2] = {0, 0};   // pinned staging of the slabs (pageable / scattered input)
    uint8_t *d_slab_[2] = {nullptr, nullptr}; size_t slab_cap_[2] = {0, 0};   // double-buffered haystack slabs
    void *copy_stream_ = nullptr;                                  // cudaStream_t of the slab uploads
    void *ev_slab_[4] = {nullptr, nullptr, nullptr, nullptr};     // per buffer: upload started / finished
    std::vector<uint32_t> off32_;

    void *async_rows_ = nullptr;   // set while scan_device_uniform_async runs launch_scan
    size_t async_cap_ = 0;
    bool async_pending_ = false;
    uint32_t async_tiles_ = 0;
    double last_density_ = 0.0;    // events per byte of the previous scan (kernel choice)
    size_t n_events_ = 0;
    uint32_t end_state_ = 0;
};

} // namespace acb200
