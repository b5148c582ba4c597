// engine.cu — see engine.hpp.
#include "engine.hpp"
#include "scan_kernels.cuh"
#include "filter_kernels.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace acb200 {

// ------------------------------------------------------------ error state --

static thread_local std::string g_error;
static thread_local int g_device = -2;   // -2: not chosen yet

void set_error(const std::string &msg) { g_error = msg; }
const char *get_error() { return g_error.c_str(); }

int preferred_device()
{
    if (g_device == -2) {
        const char *e = getenv("ACB200_DEVICE");
        if (e && *e) g_device = atoi(e);
        else {
            int d = 0;
            if (cudaGetDevice(&d) != cudaSuccess) d = 0;
            g_device = d;
        }
    }
    return g_device;
}
void set_preferred_device(int d) { g_device = d; }

#define CU_OK(call)                                                                        \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                 \
            return false;                                                                  \
        }                                                                                  \
    } while (0)

static inline cudaStream_t S(void *p) { return static_cast<cudaStream_t>(p); }

// 2-D view of a haystack stream for the TMA unit: rows of `chunk` bytes (row r = slice r), boxes of TMA_BOX_BYTES x 32 rows.
// cuTensorMapEncodeTiled is a driver-API call; the library links the runtime only, so it is looked up once.
static bool make_slice_map(CUtensorMap *map, const void *text, uint32_t chunk, uint32_t rows)
{
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    static bool looked = false;
    if (!looked) {
        looked = true;
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            encode = reinterpret_cast<encode_fn>(fn);
        else cudaGetLastError();
    }
    if (!encode) return false;
    const cuuint64_t dims[2] = {chunk, rows};
    const cuuint64_t strides[1] = {chunk};                    // bytes between rows
    const cuuint32_t box[2] = {TMA_BOX_BYTES, 32};
    const cuuint32_t elem[2] = {1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(text), dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static inline cudaEvent_t EV(void *p) { return static_cast<cudaEvent_t>(p); }

// ------------------------------------------------------------ peer memory --

// Waits (on the device, one warp) until the first word of each of n mailboxes has reached `seq` (step numbers only
// grow): every sender writes its mailbox after its events, through the same copy-engine stream, so a mailbox at
// `seq` means its rows have landed.  The same kernel makes a sender wait for the collector's acknowledgement word
// (which may live in a peer's memory) before it overwrites a slot.
__global__ void ac_mailbox_wait_kernel(const volatile uint32_t *mailboxes, uint32_t n, uint32_t stride_words, uint32_t seq)
{
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
        while (mailboxes[(size_t)i * stride_words] < seq) __nanosleep(200);
}

void *device_alloc(int device, size_t bytes)
{
    void *p = nullptr;
    if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess ||
        cudaMemset(p, 0, bytes ? bytes : 1) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
        set_error(std::string("device_alloc: ") + cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}

bool device_free(int device, void *p)
{
    CU_OK(cudaSetDevice(device));
    CU_OK(cudaFree(p));
    return true;
}

bool ipc_export(const void *dptr, unsigned char handle[64])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    cudaIpcMemHandle_t h;
    CU_OK(cudaIpcGetMemHandle(&h, const_cast<void *>(dptr)));
    memcpy(handle, &h, 64);
    return true;
}

void *ipc_open(int device, const unsigned char handle[64])
{
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void *p = nullptr;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { set_error(std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)); return nullptr; }
    return p;
}

bool ipc_close(int device, void *p)
{
    CU_OK(cudaSetDevice(device));
    CU_OK(cudaIpcCloseMemHandle(p));
    return true;
}

bool copy_async(void *dst, const void *src, size_t bytes, void *stream)
{
    if (bytes) CU_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, stream ? S(stream) : cudaStreamLegacy));
    return true;
}

bool mailbox_wait_async(int device, const void *mailboxes, uint32_t n, uint32_t stride_words, uint32_t seq, void *stream)
{
    CU_OK(cudaSetDevice(device));
    ac_mailbox_wait_kernel<<<1, 32, 0, stream ? S(stream) : cudaStreamLegacy>>>((const volatile uint32_t *)mailboxes, n, stride_words, seq);
    CU_OK(cudaGetLastError());
    return true;
}

// --------------------------------------------------------------- lifetime --

Engine::Engine() {}
Engine::~Engine() { release(); }

void Engine::release()
{
    if (device_ < 0) return;
    cudaSetDevice(device_);
    if (stream_) cudaStreamSynchronize(S(stream_));
    cudaFree(d_table_); cudaFree(d_cls_); cudaFree(d_text_); cudaFree(d_off_);
    cudaFree(d_first_); cudaFree(d_events_); cudaFree(d_tiles_);
    cudaFree(d_l1_); cudaFree(d_l2_); cudaFree(d_mask_);
    cudaFree(d_gt_slots_); cudaFree(d_gt_pat_); d_gt_slots_ = nullptr; d_gt_pat_ = nullptr; gt_log2_ = 0;
    cudaFree(d_out_off_); cudaFree(d_out_idx_); cudaFree(d_pat_len_); cudaFree(d_hit_sums_); cudaFree(d_hits_); cudaFree(d_hit_total_);
    d_out_off_ = nullptr; d_out_idx_ = nullptr; d_pat_len_ = nullptr; d_hit_sums_ = nullptr; d_hits_ = nullptr; d_hit_total_ = nullptr;
    hit_sums_cap_ = 0; hits_cap_ = 0;
    cudaFree(d_items_); cudaFree(d_recs_); cudaFree(d_desc_); cudaFree(d_tile_len_);
    d_l1_ = nullptr; d_l2_ = nullptr; d_mask_ = nullptr; mask_cap_ = 0;
    d_items_ = nullptr; d_recs_ = nullptr; d_desc_ = nullptr; d_tile_len_ = nullptr; verify_tiles_cap_ = 0;
    if (h_counters_) cudaFreeHost(h_counters_);
    if (h_events_) cudaFreeHost(h_events_);
    if (h_small_) cudaFreeHost(h_small_);
    h_small_ = nullptr; last_host_events_ = nullptr;
    for (int b = 0; b < 2; ++b) { if (h_slab_[b]) cudaFreeHost(h_slab_[b]); h_slab_[b] = nullptr; h_slab_cap_[b] = 0; }
    for (auto &e : ev_) if (e) { cudaEventDestroy(EV(e)); e = nullptr; }
    for (auto &e : ev_slab_) if (e) { cudaEventDestroy(EV(e)); e = nullptr; }
    for (int b = 0; b < 2; ++b) { cudaFree(d_slab_[b]); d_slab_[b] = nullptr; slab_cap_[b] = 0; }
    if (copy_stream_) { cudaStreamDestroy(S(copy_stream_)); copy_stream_ = nullptr; }
    if (stream_) cudaStreamDestroy(S(stream_));
    d_table_ = nullptr; d_cls_ = nullptr; d_text_ = nullptr; d_off_ = nullptr; d_first_ = nullptr;
    d_events_ = nullptr; d_tiles_ = nullptr;
    h_counters_ = nullptr; h_events_ = nullptr; stream_ = nullptr;
    device_ = -1;
}

// --------------------------------------------------------------- finalize --

template <typename E>
static bool expand_table(E *table, const FlatAutomaton &f, cudaStream_t st, uint64_t *launches)
{
    const uint32_t N = f.n_states;
    const uint32_t NR = f.n_rows;
    uint32_t *d_order = nullptr, *d_fail = nullptr, *d_src = nullptr, *d_dst = nullptr;
    uint16_t *d_cls = nullptr;
    const size_t ne = f.edge_src.size();
    bool ok = true;
    auto fail_with = [&](const char *what, cudaError_t e) {
        set_error(std::string(what) + ": " + cudaGetErrorString(e));
        ok = false;
    };
    cudaError_t e;
    if ((e = cudaMalloc(&d_order, sizeof(uint32_t) * N)) != cudaSuccess) fail_with("cudaMalloc(order)", e);
    if (ok && (e = cudaMalloc(&d_fail, sizeof(uint32_t) * NR)) != cudaSuccess) fail_with("cudaMalloc(fail)", e);
    if (ok && ne) {
        if ((e = cudaMalloc(&d_src, sizeof(uint32_t) * ne)) != cudaSuccess) fail_with("cudaMalloc(edge_src)", e);
        if (ok && (e = cudaMalloc(&d_dst, sizeof(uint32_t) * ne)) != cudaSuccess) fail_with("cudaMalloc(edge_dst)", e);
        if (ok && (e = cudaMalloc(&d_cls, sizeof(uint16_t) * ne)) != cudaSuccess) fail_with("cudaMalloc(edge_cls)", e);
    }
    auto upload = [&](void *dst, const void *src, size_t bytes, const char *what) {
        if (!ok || !bytes) return;
        if ((e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st)) != cudaSuccess) fail_with(what, e);
    };
    upload(d_order, f.bfs_order.data(), sizeof(uint32_t) * N, "upload(order)");
    upload(d_fail, f.fail.data(), sizeof(uint32_t) * NR, "upload(fail)");
    upload(d_src, f.edge_src.data(), sizeof(uint32_t) * ne, "upload(edge_src)");
    upload(d_dst, f.edge_dst.data(), sizeof(uint32_t) * ne, "upload(edge_dst)");
    upload(d_cls, f.edge_cls.data(), sizeof(uint16_t) * ne, "upload(edge_cls)");
    if (ok) {
        const size_t n_levels = f.level_off.size() - 1;
        for (size_t d = 0; d < n_levels; ++d) {
            const uint32_t lb = f.level_off[d], le = f.level_off[d + 1];
            const unsigned long long cells = (unsigned long long)(le - lb) * f.n_classes;
            if (cells) {
                const unsigned blocks = (unsigned)((cells + EXPAND_THREADS - 1) / EXPAND_THREADS);
                expand_inherit_kernel<E><<<blocks, EXPAND_THREADS, 0, st>>>(table, d_order, d_fail, lb, le, f.n_classes, f.root);
                ++*launches;
            }
            const uint32_t eb = f.level_edge_off[d], ee = f.level_edge_off[d + 1];
            if (ee > eb) {
                const unsigned blocks = (ee - eb + EXPAND_THREADS - 1) / EXPAND_THREADS;
                expand_edges_kernel<E><<<blocks, EXPAND_THREADS, 0, st>>>(table, d_src, d_dst, d_cls, eb, ee, f.n_classes);
                ++*launches;
            }
        }
        if ((e = cudaGetLastError()) != cudaSuccess) fail_with("expand launch", e);
        if (ok && (e = cudaStreamSynchronize(st)) != cudaSuccess) fail_with("expand sync", e);
    }
    cudaFree(d_order); cudaFree(d_fail); cudaFree(d_src); cudaFree(d_dst); cudaFree(d_cls);
    return ok;
}

template <typename E, bool RANGE, bool FIRST>
static cudaError_t set_smem_attr(int bytes)
{
    return cudaFuncSetAttribute(ac_scan_kernel<E, RANGE, FIRST>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

bool Engine::build(const FlatAutomaton &f, int dev, const Engine *table_src)
{
    int n_dev = 0;
    cudaError_t ce = cudaGetDeviceCount(&n_dev);
    if (ce != cudaSuccess || n_dev == 0) {
        set_error(std::string("no CUDA device available: ") + cudaGetErrorString(ce));
        return false;
    }
    if (dev < 0 || dev >= n_dev) { set_error("ACB200 device ordinal out of range"); return false; }
    CU_OK(cudaSetDevice(dev));
    device_ = dev;
    cudaDeviceProp prop;
    CU_OK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) {
        set_error("libacb200 is built for sm_100a (B200) only; found " + std::string(prop.name));
        return false;
    }
    n_sms_ = prop.multiProcessorCount;
    max_smem_optin_ = (int)prop.sharedMemPerBlockOptin;

    cudaStream_t st;
    CU_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    stream_ = st;
    for (auto &e : ev_) { cudaEvent_t x; CU_OK(cudaEventCreate(&x)); e = x; }

    n_states_ = f.n_states; n_rows_ = f.n_rows; ncls_ = f.n_classes; final_bound_ = f.final_bound; root_ = f.root;
    halo_ = f.max_pattern_len ? f.max_pattern_len - 1 : 0;
    range_map_ = f.range_map; range_lo_ = f.range_lo; n_used_ = f.n_used_bytes;
    table_entries_ = (uint64_t)n_rows_ * ncls_;
    if (table_entries_ >= (1ull << 32)) {
        set_error("automaton too large: states x classes must stay below 2^32 table entries");
        return false;
    }
    entry_bytes_ = (n_rows_ <= 65536u) ? 2 : 4;
    const size_t table_bytes = (size_t)table_entries_ * entry_bytes_;

    CU_OK(cudaMalloc(&d_table_, table_bytes + 16));
    CU_OK(cudaMemsetAsync(d_table_, 0, table_bytes + 16, st));
    CU_OK(cudaMalloc(&d_cls_, 256));
    CU_OK(cudaMemcpyAsync(d_cls_, f.cls_map, 256, cudaMemcpyHostToDevice, st));
    CU_OK(cudaMallocHost(&h_counters_, 64));

    uint64_t launches = 0;
    bool ok = false;
    if (table_src && table_src->device_ >= 0 && table_src->d_table_ && table_src->table_entries_ == table_entries_ &&
        table_src->entry_bytes_ == entry_bytes_) {
        // a replica: the expanded table comes from the primary's HBM over NVLink instead of being expanded again
        ok = cudaMemcpyPeerAsync(d_table_, dev, table_src->d_table_, table_src->device_, table_bytes, st) == cudaSuccess &&
             cudaStreamSynchronize(st) == cudaSuccess;
        if (!ok) cudaGetLastError();
    }
    if (!ok)
        ok = (entry_bytes_ == 2) ? expand_table<uint16_t>((uint16_t *)d_table_, f, st, &launches)
                                 : expand_table<uint32_t>((uint32_t *)d_table_, f, st, &launches);
    if (!ok) return false;

    const int dyn = max_smem_optin_ - 2048;   // static shared memory of the kernel + slack
    CU_OK((set_smem_attr<uint16_t, true, false>(dyn)));
    CU_OK((set_smem_attr<uint16_t, true, true>(dyn)));
    CU_OK((set_smem_attr<uint16_t, false, false>(dyn)));
    CU_OK((set_smem_attr<uint16_t, false, true>(dyn)));
    CU_OK((set_smem_attr<uint32_t, true, false>(dyn)));
    CU_OK((set_smem_attr<uint32_t, true, true>(dyn)));
    CU_OK((set_smem_attr<uint32_t, false, false>(dyn)));
    CU_OK((set_smem_attr<uint32_t, false, true>(dyn)));
    CU_OK(cudaFuncSetAttribute(ac_scan_tma_kernel<uint16_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    CU_OK(cudaFuncSetAttribute(ac_scan_tma_kernel<uint16_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    CU_OK(cudaFuncSetAttribute(ac_scan_tma_kernel<uint32_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    CU_OK(cudaFuncSetAttribute(ac_scan_tma_kernel<uint32_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));

    // output lists for the device-side hit expansion
    {
        std::vector<uint32_t> off32(f.out_off.size());
        for (size_t i = 0; i < f.out_off.size(); ++i) off32[i] = (uint32_t)f.out_off[i];
        std::vector<uint32_t> len32(f.accepted.size());
        for (size_t i = 0; i < f.accepted.size(); ++i) len32[i] = (uint32_t)f.accepted[i].ptext.length;
        CU_OK(cudaMalloc(&d_out_off_, std::max<size_t>(1, off32.size()) * sizeof(uint32_t)));
        CU_OK(cudaMalloc(&d_out_idx_, std::max<size_t>(1, f.out_idx.size()) * sizeof(uint32_t)));
        CU_OK(cudaMalloc(&d_pat_len_, std::max<size_t>(1, len32.size()) * sizeof(uint32_t)));
        CU_OK(cudaMalloc(&d_hit_total_, sizeof(unsigned long long)));
        CU_OK(cudaMemcpyAsync(d_out_off_, off32.data(), off32.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        if (!f.out_idx.empty()) CU_OK(cudaMemcpyAsync(d_out_idx_, f.out_idx.data(), f.out_idx.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        if (!len32.empty()) CU_OK(cudaMemcpyAsync(d_pat_len_, len32.data(), len32.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        CU_OK(cudaStreamSynchronize(st));
    }

    // gram prefilter tables
    filter_w_ = f.filter_w; l1_bits_ = f.l1_bits; l2_log2_ = f.l2_log2;
    if (filter_w_) {
        CU_OK(cudaMalloc(&d_l1_, f.l1.size() * sizeof(uint32_t)));
        CU_OK(cudaMemcpyAsync(d_l1_, f.l1.data(), f.l1.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        if (l2_log2_) {
            CU_OK(cudaMalloc(&d_l2_, f.l2.size() * sizeof(uint32_t)));
            CU_OK(cudaMemcpyAsync(d_l2_, f.l2.data(), f.l2.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        }
        if (f.gt_log2 && !f.gt_slots.empty() && !f.gt_pat.empty()) {
            CU_OK(cudaMalloc(&d_gt_slots_, f.gt_slots.size() * sizeof(GramSlot)));
            CU_OK(cudaMemcpyAsync(d_gt_slots_, f.gt_slots.data(), f.gt_slots.size() * sizeof(GramSlot), cudaMemcpyHostToDevice, st));
            CU_OK(cudaMalloc(&d_gt_pat_, f.gt_pat.size() * sizeof(uint32_t)));
            CU_OK(cudaMemcpyAsync(d_gt_pat_, f.gt_pat.data(), f.gt_pat.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            gt_log2_ = f.gt_log2;
        }
        CU_OK(cudaStreamSynchronize(st));
        CU_OK(cudaFuncSetAttribute(ac_filter_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FILTER_L1_BYTES));
        CU_OK(cudaFuncSetAttribute(ac_filter_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FILTER_L1_BYTES));
        CU_OK(cudaFuncSetAttribute(ac_filter_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FILTER_L1_BYTES));
        CU_OK(cudaFuncSetAttribute(ac_filter_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FILTER_L1_BYTES));
    }
    info.filter_word = (int32_t)filter_w_;
    info.min_pattern_len = f.min_pattern_len;
    info.filter_l1_fill = (float)f.l1_fill;
    info.filter_l2_log2 = l2_log2_;

    info.n_states = n_states_;
    info.n_classes = ncls_;
    info.entry_bytes = (uint32_t)entry_bytes_;
    info.max_pattern_len = f.max_pattern_len;
    info.final_bound = final_bound_;
    info.root = root_;
    info.table_bytes = table_bytes;
    info.device = device_;
    stats = ACB200_STATS_t{};
    stats.kernel_launches = launches;
    return true;
}

// ---------------------------------------------------------------- scratch --

bool Engine::ensure_text(size_t bytes)
{
    if (bytes <= text_cap_) return true;
    cudaFree(d_text_); d_text_ = nullptr; text_cap_ = 0;
    const size_t cap = std::max(bytes + bytes / 4, (size_t)1 << 20);
    CU_OK(cudaMalloc(&d_text_, cap));
    // (the scan prefetches 16-byte groups past the end of the stream, never used.)  ON the handle's stream: cudaMemset runs
    // on the legacy default stream, which this non-blocking stream does not wait for — the zeroes could land after the
    // haystack that the next line of the caller copies in (seen once a call got short enough: the one-CTA path).
    CU_OK(cudaMemsetAsync(d_text_, 0, cap, S(stream_)));
    text_cap_ = cap;
    return true;
}

bool Engine::ensure_events(size_t n)
{
    if (n <= events_cap_) return true;
    cudaFree(d_events_); d_events_ = nullptr; events_cap_ = 0;
    CU_OK(cudaMalloc(&d_events_, n * sizeof(PackedEvent)));
    events_cap_ = n;
    return true;
}

bool Engine::ensure_host_events(size_t n)
{
    if (n <= h_events_cap_) return true;
    if (h_events_) cudaFreeHost(h_events_);
    if (h_small_) cudaFreeHost(h_small_);
    h_small_ = nullptr; last_host_events_ = nullptr;
    h_events_ = nullptr; h_events_cap_ = 0;
    const size_t cap = std::max(n + n / 4, (size_t)4096);
    CU_OK(cudaMallocHost(&h_events_, cap * sizeof(PackedEvent)));
    h_events_cap_ = cap;
    return true;
}

bool Engine::ensure_offsets(size_t n)
{
    if (n > off_cap_) {
        cudaFree(d_off_); d_off_ = nullptr; off_cap_ = 0;
        const size_t cap = std::max(n + n / 4, (size_t)1024);
        CU_OK(cudaMalloc(&d_off_, cap * sizeof(uint32_t)));
        off_cap_ = cap;
    }
    return true;
}

bool Engine::ensure_mask(size_t words)
{
    if (words <= mask_cap_) return true;
    cudaFree(d_mask_); d_mask_ = nullptr; mask_cap_ = 0;
    const size_t cap = std::max(words + words / 4, (size_t)4096);
    CU_OK(cudaMalloc(&d_mask_, cap * sizeof(uint32_t)));
    mask_cap_ = cap;
    return true;
}

// scratch of the verify kernels, sized by the number of 16 KiB tiles
bool Engine::ensure_verify_scratch(size_t n_tiles)
{
    if (n_tiles <= verify_tiles_cap_) return true;
    cudaFree(d_items_); cudaFree(d_recs_); cudaFree(d_desc_); cudaFree(d_tile_len_);
    d_items_ = nullptr; d_recs_ = nullptr; d_desc_ = nullptr; d_tile_len_ = nullptr; verify_tiles_cap_ = 0;
    const size_t cap = std::max(n_tiles + n_tiles / 4, (size_t)256);
    CU_OK(cudaMalloc(&d_items_, cap * VER_DENSE_MAX * sizeof(uint32_t)));
    CU_OK(cudaMalloc(&d_recs_, cap * VER_DENSE_MAX * 2 * sizeof(uint32_t)));
    CU_OK(cudaMalloc(&d_desc_, cap * 2 * sizeof(uint32_t)));
    // one block: [16 counters | block sums | events per tile | offsets per tile] — the first three are zeroed by ONE memset
    CU_OK(cudaMalloc(&d_tile_len_, (16 + (cap / EMIT_THREADS + 16) + 2 * cap) * sizeof(uint32_t)));
    verify_tiles_cap_ = cap;
    return true;
}

bool Engine::ensure_tiles(size_t n)
{
    if (n <= tiles_cap_) return true;
    cudaFree(d_tiles_); d_tiles_ = nullptr; tiles_cap_ = 0;
    const size_t cap = std::max(n + n / 4, (size_t)1024);
    CU_OK(cudaMalloc(&d_tiles_, (cap + 4) * sizeof(unsigned long long)));      // + the full walk's eight counters behind its tiles
    tiles_cap_ = cap;
    return true;
}

// Converts the caller's 64-bit offsets to the 32-bit device array.  Batches whose
// haystacks all have one length need no array at all (*uniform_len > 0).
bool Engine::upload_offsets(const uint64_t *offsets, size_t n, uint32_t *uniform_len)
{
    *uniform_len = 0;
    const uint64_t total = offsets[n];
    if (n <= 1 || total == 0) { *uniform_len = (uint32_t)std::max<uint64_t>(total, 1); return true; }
    const uint64_t L = offsets[1] - offsets[0];
    bool uniform = L > 0;
    for (size_t i = 1; uniform && i < n; ++i) uniform = (offsets[i + 1] - offsets[i]) == L;
    if (uniform) { *uniform_len = (uint32_t)L; return true; }
    off32_.resize(n + 1);
    for (size_t i = 0; i <= n; ++i) off32_[i] = (uint32_t)offsets[i];
    if (!ensure_offsets(n + 1)) return false;
    CU_OK(cudaMemcpyAsync(d_off_, off32_.data(), (n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, S(stream_)));
    return true;
}

// ------------------------------------------------------------------- scan --

uint32_t Engine::pick_chunk(uint64_t total) const
{
    if (tune_chunk) return std::min(1u << 20, std::max(16u, (tune_chunk + 15u) & ~15u));
    auto up16 = [](uint64_t v) -> uint64_t { return (v + 15) & ~(uint64_t)15; };
    // Steady state: 512-byte slices measured best on B200 (adjacent lanes stay within a few DRAM
    // pages, the (Lmax-1)-byte halo re-read stays below ~12%); long patterns need longer slices.
    const uint64_t ideal = std::max<uint64_t>(512, up16(8ull * (halo_ + 1)));
    // Small inputs: shorter slices so that every warp of every SM still gets a tile, but never so
    // short that the halo dominates.
    // (2x the halo: measured on the adversarial shape — Lmax 1024, every byte an event — 2 KiB slices give 129 GB/s,
    // 4 KiB 96, 8 KiB 62: with long patterns the walk needs the lanes more than it minds re-reading the halo)
    // (and 32 bytes at least: a 0.25 / 1 / 4 MiB batch of config 2 walks in 25.5 / 28.7 / 32.8 us with 32-byte slices,
    // 31 / 33 / 37 us with 64-byte ones and 23.5 / 28.7 / 45.9 us with 16-byte ones — profiles/r02_midsize_calls.txt)
    const uint64_t floor_ = std::max<uint64_t>(32, up16(2ull * halo_));
    const uint64_t want = (uint64_t)n_sms_ * SCAN_THREADS;
    uint64_t c = ideal;
    if (total / ideal < want) c = std::max(floor_, up16(total / std::max<uint64_t>(want, 1)));
    return (uint32_t)std::min<uint64_t>(c, ideal);
}

// Rows of the dense table that go to shared memory: a contiguous id window around final_bound.
void Engine::window_for(size_t smem_budget, uint32_t *win_lo, uint32_t *win_rows) const
{
    const size_t row_bytes = (size_t)ncls_ * entry_bytes_;
    // one row of the budget is the staged window's sink row; window-relative ids must fit an entry
    uint64_t rows_fit = smem_budget / row_bytes;
    rows_fit = rows_fit ? rows_fit - 1 : 0;
    if (entry_bytes_ == 2) rows_fit = std::min<uint64_t>(rows_fit, 65534);
    const uint64_t n_final = final_bound_ - 1, n_plain = n_rows_ - final_bound_;
    uint64_t B = std::min<uint64_t>(n_final, rows_fit / 4);
    uint64_t A = std::min<uint64_t>(n_plain, rows_fit - B);
    B = std::min<uint64_t>(n_final, rows_fit - A);
    // a window that covers only a sliver of a huge automaton cannot pay for itself
    if (!tune_smem_bytes && A * 64 < n_plain) { A = 0; B = 0; }
    if (A == 0) B = 0;
    *win_lo = final_bound_ - (uint32_t)B;
    *win_rows = (uint32_t)(A + B);
}

template <typename E, bool RANGE, bool FIRST>
static void launch_kernel(const ScanArgs &a, unsigned grid, size_t smem, cudaStream_t st)
{
    ac_scan_kernel<E, RANGE, FIRST><<<grid, SCAN_THREADS, smem, st>>>(a);
}

bool Engine::launch_scan(const void *d_text, uint32_t total, uint32_t readable, size_t n_hay, uint32_t uniform_len,
                         bool first_only, uint32_t init_state, void *stream)
{
    cudaStream_t st = stream ? S(stream) : S(stream_);
    n_events_ = 0;
    end_state_ = (init_state == ROOT_STATE) ? root_ : init_state;
    stats.bytes = total; stats.events = 0; stats.kernel_launches = 0; stats.kernel_ms = 0;
    stats.halo_bytes = halo_;
    stats.filtered = 0; stats.filter_ms = 0; stats.verify_ms = 0; stats.flagged_words = 0; stats.dense_tiles = 0;
    stats.reorder_ms = 0; stats.expand_ms = 0;
    if (total == 0) { stats.chunk_bytes = 0; return true; }

    // Gram prefilter: needs an eligible dictionary and a walk that starts at the root.  Automatic mode
    // skips it for inputs too small to amortise a second launch and after a scan whose tiles were mostly
    // walked completely anyway.
    // findAll=false may use it too while events are sparse: the prefilter path returns every event and the host
    // keeps the first one per haystack; with dense events the FIRST kernel's early exit wins.
    if (filter_w_ && tune_filter >= 0 && (init_state == ROOT_STATE || init_state == root_) &&
        (!first_only || last_density_ < 1.0 / 2048)) {
        const bool want = tune_filter > 0 || (total >= (8u << 20) && last_dense_frac_ < 0.5);
        if (want) return launch_filtered(d_text, total, readable, n_hay, uniform_len, stream);
        last_dense_frac_ *= 0.5;      // re-probe the prefilter now and then
    }

    if (async_rows_) { set_error("asynchronous search needs the prefilter path (dictionary, batch size or density rule it out)"); return false; }
    const uint32_t chunk = pick_chunk(total);
    const uint32_t n_chunks = (uint32_t)(((uint64_t)total + chunk - 1) / chunk);
    stats.chunk_bytes = chunk;
    if (events_cap_ == 0 && !ensure_events(std::max<size_t>(1 << 16, total / 64))) return false;
    if (first_only) {
        if (n_hay > first_cap_) {
            cudaFree(d_first_); d_first_ = nullptr; first_cap_ = 0;
            CU_OK(cudaMalloc(&d_first_, (n_hay + n_hay / 4 + 16) * sizeof(uint32_t)));
            first_cap_ = n_hay + n_hay / 4 + 16;
        }
    }

    // Text through the TMA unit (ac_scan_tma_kernel) where its box fits: patterns of up to 33 bytes (the warm-up is one
    // box), slices that are a whole number of boxes, enough complete rows to matter.  tune_tma: 0 auto, 1 on, -1 off.
    CUtensorMap tmap;
    const uint32_t full_rows = total / chunk;
    // (automatic: only after a call with few events — the ring takes 64 KB from the table window, and a text full of
    // needles leaves a smaller window that often: 1 GiB of config 2, one needle per KiB 1.41 vs 1.27 ms, none 0.73 vs 0.98 ms)
    const bool tma = (tune_tma > 0 || (tune_tma == 0 && last_density_ < 1.0 / 8192)) && !first_only && halo_ <= TMA_BOX_BYTES && chunk % TMA_BOX_BYTES == 0 && chunk >= 2 * TMA_BOX_BYTES &&
                     full_rows >= 64 && (((uintptr_t)d_text) & 15u) == 0 && make_slice_map(&tmap, d_text, chunk, full_rows);

    // Hot window in shared memory: B shallowest finals + A shallowest non-finals (root first).
    const int dyn_max = max_smem_optin_ - 2048 - (tma ? (int)TMA_RING_BYTES + 768 : 0);
    size_t smem_budget = (size_t)dyn_max - 16;     // (16 bytes: stage_window's alignment shift)
    if (tune_smem_bytes) smem_budget = std::min<size_t>(smem_budget, tune_smem_bytes);
    uint32_t win_lo = 0, win_rows = 0;
    window_for(smem_budget, &win_lo, &win_rows);
    const size_t row_bytes = (size_t)ncls_ * entry_bytes_;
    const size_t smem_bytes = std::max<size_t>(16, ((size_t)win_rows + 1) * row_bytes) + 16;     // (+ stage_window's alignment shift)

    const uint32_t n_tiles = (n_chunks + 31u) / 32u;

    if (!ensure_tiles(n_tiles)) return false;

    ScanArgs a{};
    a.text = (const uint8_t *)d_text;
    a.hay_off = uniform_len ? nullptr : d_off_;
    a.n_hay = (uint32_t)n_hay;
    a.uniform_len = uniform_len;
    a.total = total;
    a.readable = readable;
    a.chunk = chunk;
    a.halo = halo_;
    a.chunk_begin = 0;
    a.chunk_end = n_chunks;
    a.n_tiles = n_tiles;
    a.table = d_table_;
    a.cls_map = d_cls_;
    a.ncls = ncls_;
    a.final_bound = final_bound_;
    a.root = root_;
    a.win_lo = win_lo;
    a.win_rows = win_rows;
    a.range_lo = range_lo_;
    a.n_used = n_used_;
    a.init_state = (init_state == ROOT_STATE) ? root_ : init_state;
    a.tile_status = d_tiles_;
    a.counters = reinterpret_cast<uint32_t *>(d_tiles_ + n_tiles);       // right behind the tile status words: ONE memset clears both
    a.first_end = d_first_;
    a.host_counters = h_counters_;               // pinned memory, mapped into the device's address space (UVA)
    const unsigned warps_per_cta = SCAN_THREADS / 32;
    // Tiles are handed out by ticket, so any grid is correct.  One CTA per 32 tiles (a tile per warp) left most SMs idle
    // for mid-size inputs and long patterns (config 5: 256 MiB = 1,024 tiles of 256 KiB -> 32 of 148 SMs): spread
    // the tiles over the SMs instead, at least four per CTA so that a CTA's table staging is shared by some work.
    // (Round 2, with the 2 KiB slices pick_chunk() now gives long patterns: config 5 went from 20 to 127-129 GB/s read.)
    const unsigned grid = std::min<uint32_t>(std::max<uint32_t>((n_tiles + warps_per_cta - 1) / warps_per_cta, (n_tiles + 3u) / 4u),
                                             (uint32_t)n_sms_);

    for (int attempt = 0; attempt < 2; ++attempt) {
        a.out = (uint2 *)d_events_;
        a.capacity = (uint32_t)std::min<size_t>(events_cap_, 0xffffffffu);
        CU_OK(cudaMemsetAsync(d_tiles_, 0, ((size_t)n_tiles + 4) * sizeof(unsigned long long), st));
        if (first_only) CU_OK(cudaMemsetAsync(d_first_, 0xff, n_hay * sizeof(uint32_t), st));
        h_counters_[1] = 0; h_counters_[2] = 0;     // (a findAll=false call may skip its last slice and with it the end state)
        CU_OK(cudaEventRecord(EV(ev_[0]), st));
        if (tma) {
            const size_t smem_tma = smem_bytes + TMA_RING_BYTES + 128;      // + alignment slack of the ring
            if (entry_bytes_ == 2) {
                if (range_map_) ac_scan_tma_kernel<uint16_t, true><<<grid, SCAN_THREADS, smem_tma, st>>>(a, tmap);
                else ac_scan_tma_kernel<uint16_t, false><<<grid, SCAN_THREADS, smem_tma, st>>>(a, tmap);
            } else {
                if (range_map_) ac_scan_tma_kernel<uint32_t, true><<<grid, SCAN_THREADS, smem_tma, st>>>(a, tmap);
                else ac_scan_tma_kernel<uint32_t, false><<<grid, SCAN_THREADS, smem_tma, st>>>(a, tmap);
            }
        } else if (entry_bytes_ == 2) {
            if (range_map_) { if (first_only) launch_kernel<uint16_t, true, true>(a, grid, smem_bytes, st); else launch_kernel<uint16_t, true, false>(a, grid, smem_bytes, st); }
            else            { if (first_only) launch_kernel<uint16_t, false, true>(a, grid, smem_bytes, st); else launch_kernel<uint16_t, false, false>(a, grid, smem_bytes, st); }
        } else {
            if (range_map_) { if (first_only) launch_kernel<uint32_t, true, true>(a, grid, smem_bytes, st); else launch_kernel<uint32_t, true, false>(a, grid, smem_bytes, st); }
            else            { if (first_only) launch_kernel<uint32_t, false, true>(a, grid, smem_bytes, st); else launch_kernel<uint32_t, false, false>(a, grid, smem_bytes, st); }
        }
        CU_OK(cudaGetLastError());
        CU_OK(cudaEventRecord(EV(ev_[1]), st));
        stats.kernel_launches += 1;
        CU_OK(cudaStreamSynchronize(st));          // (the kernel has written the event total and the end state to h_counters_)
        float ms = 0;
        cudaEventElapsedTime(&ms, EV(ev_[0]), EV(ev_[1]));
        stats.kernel_ms += ms;
        const size_t found = h_counters_[1];
        end_state_ = h_counters_[2];
        if (found <= events_cap_) {
            n_events_ = found; stats.events = found;
            last_density_ = (double)found / (double)total;
            return true;
        }
        // event buffer too small: grow to the exact need and scan again
        if (!ensure_events(found + found / 16 + 1024)) return false;
    }
    set_error("event buffer overflow persisted after regrow");
    return false;
}

template <int W>
static void launch_filter_k(const FilterArgs &fa, bool l2, unsigned grid, cudaStream_t st)
{
    if (l2) ac_filter_kernel<W, true><<<grid, SCAN_THREADS, FILTER_L1_BYTES, st>>>(fa);
    else ac_filter_kernel<W, false><<<grid, SCAN_THREADS, FILTER_L1_BYTES, st>>>(fa);
}

// Programmatic dependent launch: the kernel is put on the SMs while the one before it in the stream (which calls
// cudaTriggerProgrammaticLaunchCompletion at its start) is still running, stages its class map, and waits in
// cudaGridDependencySynchronize() for that kernel's results — the launch latency and the prologue of the two short
// kernels of a step that follow another kernel directly (walk after collect, emit after offsets) leave the critical path.
template <typename K>
static void launch_dependent(K kernel, const VerifyArgs &a, unsigned grid, unsigned block, cudaStream_t st)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kernel, a) != cudaSuccess) {
        cudaGetLastError();
        kernel<<<grid, block, 0, st>>>(a);          // (a driver that refuses the attribute: the plain, fully ordered launch)
    }
}

template <typename E, int W>
static void launch_walk_k(const VerifyArgs &a, bool range, unsigned grid, cudaStream_t st)
{
    if (range) launch_dependent(ac_walk_kernel<E, true, W>, a, grid, WALK_THREADS, st);
    else launch_dependent(ac_walk_kernel<E, false, W>, a, grid, WALK_THREADS, st);
}

template <typename E, int W>
static void launch_emit_k(const VerifyArgs &a, bool range, unsigned grid, cudaStream_t st)
{
    if (range) launch_dependent(ac_emit_kernel<E, true, W>, a, grid, COUNT_THREADS, st);
    else launch_dependent(ac_emit_kernel<E, false, W>, a, grid, COUNT_THREADS, st);
}

// ahocorasick_match() through the gram prefilter: ac_filter_kernel streams the haystack and flags aligned
// words; ac_collect_kernel / ac_walk_kernel / ac_tile_count_kernel / ac_emit_kernel walk the automaton
// around the flagged words only and write the ordered event list.
bool Engine::launch_filtered(const void *d_text, uint32_t total, uint32_t readable, size_t n_hay,
                             uint32_t uniform_len, void *stream)
{
    cudaStream_t st = stream ? S(stream) : S(stream_);
    const uint32_t W = filter_w_;
    const uint32_t NB = 16 / W;
    const uint32_t n_spans = (uint32_t)(((uint64_t)total + SPAN_BYTES - 1) / SPAN_BYTES);
    const uint32_t n_tiles = (n_spans + 31) / 32;      // 16 KiB tiles
    stats.chunk_bytes = SPAN_BYTES;
    stats.filtered = 1;
    if (events_cap_ == 0 && !ensure_events(std::max<size_t>(1 << 16, total / 64))) return false;
    if (!ensure_mask((size_t)n_spans * NB)) return false;

    const uint32_t warm = (halo_ + W - 1) / W * W;
    // walking a whole tile costs ~(512 + halo) steps per lane, a flagged word (warm + W) steps on one lane
    const uint32_t dense_max = std::max<uint32_t>(32u, std::min<uint32_t>(VER_DENSE_MAX, 32u * (SPAN_BYTES + halo_) / (warm + W)));
    if (!ensure_verify_scratch(n_tiles)) return false;

    FilterArgs fa{};
    fa.text = (const uint8_t *)d_text;
    fa.total = total;
    fa.l1 = d_l1_; fa.l1_bits = l1_bits_;
    fa.l2 = d_l2_; fa.l2_shift = l2_log2_ ? 32 - l2_log2_ : 0;
    fa.mask = d_mask_;
    fa.n_spans = n_spans;
    uint32_t *const vcounters = d_tile_len_;                              // counters of the prefilter path
    uint32_t *const vblock_sum = d_tile_len_ + 16;
    uint32_t *const vtile_len = vblock_sum + (verify_tiles_cap_ / EMIT_THREADS + 16);

    VerifyArgs va{};
    ScanArgs &a = va.s;
    a.text = (const uint8_t *)d_text;
    a.hay_off = uniform_len ? nullptr : d_off_;
    a.n_hay = (uint32_t)n_hay;
    a.uniform_len = uniform_len;
    a.total = total;
    a.readable = readable;
    a.chunk = SPAN_BYTES;
    a.halo = halo_;
    a.chunk_begin = 0;
    a.chunk_end = n_spans;
    a.n_tiles = n_tiles;
    a.table = d_table_;
    a.cls_map = d_cls_;
    a.ncls = ncls_;
    a.final_bound = final_bound_;
    a.root = root_;
    a.win_lo = final_bound_;
    a.win_rows = 0;
    a.range_lo = range_lo_;
    a.n_used = n_used_;
    a.init_state = root_;
    a.tile_status = nullptr;
    a.counters = vcounters;
    a.first_end = nullptr;
    va.mask = d_mask_;
    va.n_spans = n_spans;
    va.n_tiles = n_tiles;
    va.dense_max = dense_max;
    va.warm = warm;
    va.want_end_state = (n_hay == 1) ? 1u : 0u;
    const bool direct = gt_log2_ && tune_direct >= 0;
    va.gt_slots = direct ? (const uint4 *)d_gt_slots_ : nullptr;
    va.gt_pat = d_gt_pat_;
    va.gt_log2 = direct ? gt_log2_ : 0u;
    va.items = d_items_;
    va.desc = (uint2 *)d_desc_;
    va.recs = (uint2 *)d_recs_;
    va.tile_len = vtile_len;
    va.tile_off = vtile_len + verify_tiles_cap_;
    va.block_sum = vblock_sum;
    va.host_counters = h_counters_;              // pinned memory, mapped into the device's address space (UVA)

    // tune_direct: 0 / 1 flagged words are settled by one comparison inside ac_walk_kernel where the gram table allows;
    // -1 every flagged word is walked.
    const unsigned warps_per_cta = SCAN_THREADS / 32;
    const unsigned tiles_per_cta = COLLECT_THREADS / 32;
    // grid-stride kernels, measured on B200: 8 / 12 / 16 / 24 / 40 CTAs per SM for the walk give 0.108 / 0.106 / 0.102 /
    // 0.101 / 0.101 ms (five are resident: more, smaller CTAs even out the tail; exactly one resident wave was slower),
    // 4 / 8 / 16 / 32 for emit make no difference
    const unsigned grid_w = (unsigned)n_sms_ * 16u;
    const unsigned grid_e = (n_tiles + EMIT_THREADS - 1) / EMIT_THREADS;
    const unsigned grid_n = std::min<uint32_t>((n_tiles + COUNT_THREADS / 32 - 1) / (COUNT_THREADS / 32), (uint32_t)n_sms_ * 8u);

    const bool async = async_rows_ != nullptr;
    for (int attempt = 0; attempt < 2; ++attempt) {
        a.out = (uint2 *)d_events_;
        a.capacity = (uint32_t)std::min<size_t>(events_cap_, 0xffffffffu);
        if (async) {                 // events go straight into the caller's rows; what does not fit is counted, not written
            a.out = (uint2 *)async_rows_ + 1;
            va.count_row = (uint2 *)async_rows_;
            a.capacity = (uint32_t)std::min<size_t>(async_cap_, 0xffffffffu);
        }
        // counters, block sums and events per tile (the walk kernel adds to both) start at zero: on the first attempt the
        // filter kernel's CTA 0 clears the first two and ac_collect_kernel every tile's count — no memset in front of a call
        fa.zero = vcounters; fa.n_zero = (uint32_t)(vtile_len - vcounters);
        va.clear_tile_len = attempt == 0 ? 1u : 0u;
        if (attempt != 0)
            CU_OK(cudaMemsetAsync(vcounters, 0, (size_t)((vtile_len - vcounters) + n_tiles) * sizeof(uint32_t), st));
        CU_OK(cudaEventRecord(EV(ev_[0]), st));
        if (attempt == 0) {      // the bit planes survive a regrow of the event buffer
            fa.span_begin = 0;
            fa.span_end = n_spans;
            const unsigned grid_f = std::min<uint32_t>((n_spans + warps_per_cta - 1) / warps_per_cta, (uint32_t)n_sms_);
            if (W == 8) launch_filter_k<8>(fa, d_l2_ != nullptr, grid_f, st);
            else launch_filter_k<4>(fa, d_l2_ != nullptr, grid_f, st);
            stats.kernel_launches += 1;
        }
        CU_OK(cudaEventRecord(EV(ev_[4]), st));
        va.tile_begin = 0; va.tile_end = n_tiles;
        va.item_base = 0;
        va.counter_slot = 8u;
        {
            const unsigned grid_c = std::min<uint32_t>((n_tiles + tiles_per_cta - 1) / tiles_per_cta, (uint32_t)n_sms_ * 2u);
            if (W == 8) ac_collect_kernel<8><<<grid_c, COLLECT_THREADS, 0, st>>>(va);
            else ac_collect_kernel<4><<<grid_c, COLLECT_THREADS, 0, st>>>(va);
            if (entry_bytes_ == 2) {
                if (W == 8) launch_walk_k<uint16_t, 8>(va, range_map_, grid_w, st);
                else launch_walk_k<uint16_t, 4>(va, range_map_, grid_w, st);
            } else {
                if (W == 8) launch_walk_k<uint32_t, 8>(va, range_map_, grid_w, st);
                else launch_walk_k<uint32_t, 4>(va, range_map_, grid_w, st);
            }
            stats.kernel_launches += 2;
        }
        CU_OK(cudaEventRecord(EV(ev_[5]), st));
        ac_offsets_kernel<<<grid_e, EMIT_THREADS, 0, st>>>(va);
        if (entry_bytes_ == 2) {
            if (W == 8) launch_emit_k<uint16_t, 8>(va, range_map_, grid_n, st);
            else launch_emit_k<uint16_t, 4>(va, range_map_, grid_n, st);
        } else {
            if (W == 8) launch_emit_k<uint32_t, 8>(va, range_map_, grid_n, st);
            else launch_emit_k<uint32_t, 4>(va, range_map_, grid_n, st);
        }
        CU_OK(cudaGetLastError());
        CU_OK(cudaEventRecord(EV(ev_[1]), st));
        stats.kernel_launches += 2;
        if (async) {
            // row 0 = {event count, dense tiles}; the caller waits (after its own work on the stream) and calls async_finish()
            async_pending_ = true;
            async_tiles_ = n_tiles;
            return true;
        }
        CU_OK(cudaStreamSynchronize(st));          // (the emit kernel has written the counters to h_counters_)
        float ms_f = 0, ms_v = 0, ms_r = 0;
        cudaEventElapsedTime(&ms_f, EV(ev_[0]), EV(ev_[4]));
        cudaEventElapsedTime(&ms_v, EV(ev_[4]), EV(ev_[5]));
        cudaEventElapsedTime(&ms_r, EV(ev_[5]), EV(ev_[1]));
        stats.filter_ms += ms_f; stats.verify_ms += ms_v; stats.reorder_ms += ms_r;
        stats.kernel_ms += ms_f + ms_v + ms_r;
        const size_t found = h_counters_[1];
        end_state_ = h_counters_[2];
        if (attempt == 0) stats.flagged_words = h_counters_[3];
        stats.dense_tiles = h_counters_[4];
        if (found <= events_cap_) {
            n_events_ = found; stats.events = found;
            last_density_ = (double)found / (double)total;
            last_dense_frac_ = (double)stats.dense_tiles / (double)n_tiles;
            return true;
        }
        if (!ensure_events(found + found / 16 + 1024)) return false;
    }
    set_error("event buffer overflow persisted after regrow");
    return false;
}

bool Engine::expand_hits_to_host(size_t n_hay, ACB200_HIT_t *hits, size_t cap, size_t *n_hits)
{
    CU_OK(cudaSetDevice(device_));
    cudaStream_t st = S(stream_);
    *n_hits = 0;
    stats.expand_ms = 0;
    if (n_events_ == 0) return true;
    const uint32_t n_blocks = (uint32_t)((n_events_ + HIT_THREADS - 1) / HIT_THREADS);
    if (n_blocks > hit_sums_cap_) {
        cudaFree(d_hit_sums_); d_hit_sums_ = nullptr; hit_sums_cap_ = 0;
        CU_OK(cudaMalloc(&d_hit_sums_, (size_t)(n_blocks + n_blocks / 4 + 64) * sizeof(uint32_t)));
        hit_sums_cap_ = n_blocks + n_blocks / 4 + 64;
    }
    HitArgs ha{};
    ha.events = (const uint2 *)d_events_;
    ha.n_events = (uint32_t)n_events_;
    ha.out_off = d_out_off_; ha.out_idx = d_out_idx_; ha.pat_len = d_pat_len_;
    ha.hay_off = last_uniform_len_ ? nullptr : d_off_;
    ha.n_hay = (uint32_t)n_hay; ha.uniform_len = last_uniform_len_;
    ha.block_sum = d_hit_sums_;
    ha.total = d_hit_total_;
    unsigned long long total = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        ha.hits = (uint4 *)d_hits_;
        ha.capacity = hits_cap_;
        CU_OK(cudaEventRecord(EV(ev_[2]), st));
        ac_hit_count_kernel<<<n_blocks, HIT_THREADS, 0, st>>>(ha);
        ac_hit_write_kernel<<<n_blocks, HIT_THREADS, 0, st>>>(ha);
        CU_OK(cudaGetLastError());
        CU_OK(cudaEventRecord(EV(ev_[3]), st));
        CU_OK(cudaMemcpyAsync(&total, d_hit_total_, sizeof(total), cudaMemcpyDeviceToHost, st));
        CU_OK(cudaStreamSynchronize(st));
        float ms = 0;
        cudaEventElapsedTime(&ms, EV(ev_[2]), EV(ev_[3]));
        stats.expand_ms += ms;
        stats.kernel_launches += 2;
        if (total <= hits_cap_) break;
        // hit buffer too small: grow to what the caller can take (plus the exact need if that is smaller)
        const size_t want = (size_t)std::min<unsigned long long>(total, std::max<unsigned long long>(cap, 1));
        if (want <= hits_cap_) break;            // the caller's buffer is the limit: a truncated expansion is enough
        cudaFree(d_hits_); d_hits_ = nullptr; hits_cap_ = 0;
        CU_OK(cudaMalloc(&d_hits_, want * sizeof(ACB200_HIT_t)));
        hits_cap_ = want;
    }
    *n_hits = (size_t)total;
    const size_t n_copy = (size_t)std::min<unsigned long long>(std::min<unsigned long long>(total, cap), hits_cap_);
    if (n_copy) {
        CU_OK(cudaMemcpyAsync(hits, d_hits_, n_copy * sizeof(ACB200_HIT_t), cudaMemcpyDeviceToHost, st));
        CU_OK(cudaStreamSynchronize(st));
    }
    return true;
}

bool Engine::copy_events_to(void *d_dst, size_t n, void *stream)
{
    CU_OK(cudaSetDevice(device_));
    cudaStream_t st = stream ? S(stream) : S(stream_);
    CU_OK(cudaMemcpyAsync(d_dst, d_events_, n * sizeof(PackedEvent), cudaMemcpyDeviceToDevice, st));
    return true;
}

bool Engine::scan_device(const void *d_bytes, const uint64_t *offsets, size_t n, bool first_only,
                         uint32_t init_state, void *stream)
{
    if (device_ < 0) { set_error("automaton has no device table (finalize failed?)"); return false; }
    CU_OK(cudaSetDevice(device_));
    const uint64_t total = offsets[n];
    if (total >= MAX_STREAM_BYTES) { set_error("haystack stream exceeds 4 GiB per call"); return false; }
    if (((uintptr_t)d_bytes & 15u) != 0) { set_error("device haystack pointer must be 16-byte aligned"); return false; }
    uint32_t uniform_len = 0;
    if (!upload_offsets(offsets, n, &uniform_len)) return false;
    if (!uniform_len && stream && S(stream) != S(stream_)) CU_OK(cudaStreamSynchronize(S(stream_)));
    stats.h2d_ms = 0; stats.d2h_ms = 0;
    return launch_scan(d_bytes, (uint32_t)total, (uint32_t)total, n, uniform_len, first_only, init_state, stream);
}

bool Engine::scan_device_uniform(const void *d_bytes, size_t n, size_t hay_len, bool first_only, void *stream)
{
    if (device_ < 0) { set_error("automaton has no device table (finalize failed?)"); return false; }
    CU_OK(cudaSetDevice(device_));
    const uint64_t total = (uint64_t)n * hay_len;
    if (total >= MAX_STREAM_BYTES) { set_error("haystack stream exceeds 4 GiB per call"); return false; }
    if (((uintptr_t)d_bytes & 15u) != 0) { set_error("device haystack pointer must be 16-byte aligned"); return false; }
    stats.h2d_ms = 0; stats.d2h_ms = 0;
    const uint32_t uniform_len = (n <= 1 || total == 0) ? (uint32_t)std::max<uint64_t>(total, 1) : (uint32_t)hay_len;
    return launch_scan(d_bytes, (uint32_t)total, (uint32_t)total, std::max<size_t>(n, 1), uniform_len, first_only, ROOT_STATE, stream);
}

bool Engine::scan_device_uniform_async(const void *d_bytes, size_t n, size_t hay_len, void *d_rows, size_t max_events, void *stream)
{
    if (device_ < 0) { set_error("automaton has no device table (finalize failed?)"); return false; }
    CU_OK(cudaSetDevice(device_));
    const uint64_t total = (uint64_t)n * hay_len;
    if (total == 0 || total >= MAX_STREAM_BYTES) { set_error("asynchronous search: empty batch or stream beyond 4 GiB"); return false; }
    if (((uintptr_t)d_bytes & 15u) != 0 || ((uintptr_t)d_rows & 7u) != 0) { set_error("device pointers must be 16-byte (haystack) / 8-byte (rows) aligned"); return false; }
    stats.h2d_ms = 0; stats.d2h_ms = 0;
    const uint32_t uniform_len = (n <= 1) ? (uint32_t)total : (uint32_t)hay_len;
    async_rows_ = d_rows; async_cap_ = max_events; async_pending_ = false;
    // The point of this call is ordering with the caller's later work on ITS stream: a NULL handle here means the
    // legacy default stream itself (the synchronous calls read NULL as "the library's own stream").
    if (!stream) stream = (void *)cudaStreamLegacy;
    const bool ok = launch_scan(d_bytes, (uint32_t)total, (uint32_t)total, std::max<size_t>(n, 1), uniform_len, false, ROOT_STATE, stream);
    async_rows_ = nullptr; async_cap_ = 0;
    return ok && async_pending_;
}

// The caller has waited for the stream: times from the recorded events, counts from what it read in row 0.
void Engine::async_finish(size_t n_events, size_t dense_tiles)
{
    if (!async_pending_) return;
    async_pending_ = false;
    cudaSetDevice(device_);
    float ms_f = 0, ms_v = 0, ms_r = 0;
    if (cudaEventElapsedTime(&ms_f, EV(ev_[0]), EV(ev_[4])) != cudaSuccess) ms_f = 0;
    if (cudaEventElapsedTime(&ms_v, EV(ev_[4]), EV(ev_[5])) != cudaSuccess) ms_v = 0;
    if (cudaEventElapsedTime(&ms_r, EV(ev_[5]), EV(ev_[1])) != cudaSuccess) ms_r = 0;
    cudaGetLastError();
    stats.filter_ms = ms_f; stats.verify_ms = ms_v; stats.reorder_ms = ms_r;
    stats.kernel_ms = ms_f + ms_v + ms_r;
    stats.events = n_events;
    n_events_ = 0;                           // the events live in the caller's rows, not in the library's buffer
    last_density_ = stats.bytes ? (double)n_events / (double)stats.bytes : 0.0;
    stats.dense_tiles = dense_tiles;
    if (async_tiles_) last_dense_frac_ = (double)dense_tiles / (double)async_tiles_;      // feeds the automatic kernel choice
}

// Pinned staging memory of slab buffer `buf` for callers whose haystacks are pageable or scattered (the gather
// of ac_trie_search_batch): fill it, then slab_upload_async(buf, <this pointer>, n).
char *Engine::slab_staging(int buf, size_t n_bytes)
{
    if (device_ < 0) { set_error("automaton has no device table (finalize failed?)"); return nullptr; }
    if (cudaSetDevice(device_) != cudaSuccess) { set_error("cudaSetDevice failed"); return nullptr; }
    if (n_bytes > h_slab_cap_[buf]) {
        if (h_slab_[buf]) cudaFreeHost(h_slab_[buf]);
        h_slab_[buf] = nullptr; h_slab_cap_[buf] = 0;
        const size_t cap = n_bytes + n_bytes / 8 + 4096;
        if (cudaMallocHost(&h_slab_[buf], cap) != cudaSuccess) { set_error("cudaMallocHost(slab staging) failed"); return nullptr; }
        h_slab_cap_[buf] = cap;
    }
    return (char *)h_slab_[buf];
}

bool Engine::slab_upload_begin(int buf, size_t n_bytes)
{
    if (device_ < 0) { set_error("automaton has no device table (finalize failed?)"); return false; }
    CU_OK(cudaSetDevice(device_));
    if (!copy_stream_) {
        cudaStream_t cs;
        CU_OK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        copy_stream_ = cs;
        for (auto &e : ev_slab_) { cudaEvent_t x; CU_OK(cudaEventCreate(&x)); e = x; }
    }
    if (n_bytes + 64 > slab_cap_[buf]) {
        cudaFree(d_slab_[buf]); d_slab_[buf] = nullptr; slab_cap_[buf] = 0;
        const size_t cap = n_bytes + n_bytes / 8 + 4096;
        CU_OK(cudaMalloc(&d_slab_[buf], cap));
        CU_OK(cudaMemsetAsync(d_slab_[buf], 0, cap, S(copy_stream_)));       // on the stream that uploads into it
        slab_cap_[buf] = cap;
    }
    CU_OK(cudaEventRecord(EV(ev_slab_[2 * buf]), S(copy_stream_)));
    return true;
}

// bytes [offset, offset + n_bytes) of slab_staging(buf) -> the same bytes of the device slab.  Any thread.
bool Engine::slab_upload_part(int buf, size_t offset, size_t n_bytes)
{
    if (!n_bytes) return true;
    CU_OK(cudaSetDevice(device_));
    CU_OK(cudaMemcpyAsync((char *)d_slab_[buf] + offset, (const char *)h_slab_[buf] + offset, n_bytes, cudaMemcpyHostToDevice, S(copy_stream_)));
    return true;
}

bool Engine::slab_upload_end(int buf)
{
    CU_OK(cudaSetDevice(device_));
    CU_OK(cudaEventRecord(EV(ev_slab_[2 * buf + 1]), S(copy_stream_)));
    return true;
}

bool Engine::slab_upload_async(int buf, const char *bytes, size_t n_bytes)
{
    if (!slab_upload_begin(buf, n_bytes)) return false;
    if (n_bytes) CU_OK(cudaMemcpyAsync(d_slab_[buf], bytes, n_bytes, cudaMemcpyHostToDevice, S(copy_stream_)));
    return slab_upload_end(buf);
}

float Engine::slab_h2d_ms(int buf)
{
    float ms = 0;
    if (ev_slab_[2 * buf] && cudaEventElapsedTime(&ms, EV(ev_slab_[2 * buf]), EV(ev_slab_[2 * buf + 1])) != cudaSuccess) ms = 0;
    return ms;
}

// Scans the slab uploaded into buffer `buf` (offsets are relative to the slab) and brings its events to
// host_events().  Returns after the events have arrived; the other buffer's upload keeps running meanwhile.
bool Engine::scan_slab(int buf, const uint64_t *offsets, size_t n, bool first_only, uint32_t init_state)
{
    CU_OK(cudaSetDevice(device_));
    cudaStream_t st = S(stream_);
    const uint64_t total = offsets[n];
    if (total >= MAX_STREAM_BYTES) { set_error("haystack stream exceeds 4 GiB per call"); return false; }
    uint32_t uniform_len = 0;
    if (!upload_offsets(offsets, n, &uniform_len)) return false;
    CU_OK(cudaStreamWaitEvent(st, EV(ev_slab_[2 * buf + 1]), 0));
    if (!launch_scan(d_slab_[buf], (uint32_t)total, (uint32_t)std::min<uint64_t>(slab_cap_[buf], 0xffffffffu), n, uniform_len,
                     first_only, init_state, nullptr)) return false;
    last_uniform_len_ = uniform_len;
    stats.d2h_ms = 0;
    if (n_events_) {
        if (!ensure_host_events(n_events_)) return false;
        CU_OK(cudaEventRecord(EV(ev_[2]), st));
        last_host_events_ = h_events_;
        CU_OK(cudaMemcpyAsync(h_events_, d_events_, n_events_ * sizeof(PackedEvent), cudaMemcpyDeviceToHost, st));
        CU_OK(cudaEventRecord(EV(ev_[3]), st));
        CU_OK(cudaStreamSynchronize(st));
        float ms = 0;
        cudaEventElapsedTime(&ms, EV(ev_[2]), EV(ev_[3]));
        stats.d2h_ms = ms;
    }
    return true;
}

// One short text: host-to-device copy from pinned staging, ONE launch of a one-CTA kernel that writes the ordered
// events, their count and the end state into mapped host memory, one wait.
bool Engine::scan_small(const char *bytes, uint32_t total, uint32_t init_state)
{
    cudaStream_t st = S(stream_);
    constexpr size_t HDR = 64, EV_BYTES = (size_t)SMALL_TEXT_BYTES * sizeof(PackedEvent);
    if (!h_small_) {
        CU_OK(cudaHostAlloc(&h_small_, HDR + EV_BYTES + SMALL_TEXT_BYTES, cudaHostAllocMapped));
        if (!ensure_text(SMALL_TEXT_BYTES + 64)) return false;
    }
    uint32_t *hdr = reinterpret_cast<uint32_t *>(h_small_);
    PackedEvent *ev = reinterpret_cast<PackedEvent *>(h_small_ + HDR);
    uint8_t *stage = h_small_ + HDR + EV_BYTES;
    memcpy(stage, bytes, total);
    CU_OK(cudaMemcpyAsync(d_text_, stage, total, cudaMemcpyHostToDevice, st));
    const uint32_t want = (total + 15u) / 16u;
    const uint32_t threads = std::min<uint32_t>(SMALL_THREADS, (want + 31u) & ~31u);
    SmallArgs a{};
    a.text = d_text_; a.total = total; a.readable = (uint32_t)std::min<size_t>(text_cap_, 0xffffffffu);
    a.chunk = (((total + threads - 1) / threads) + 15u) & ~15u;
    a.halo = halo_;
    a.table = d_table_; a.cls_map = d_cls_; a.ncls = ncls_; a.final_bound = final_bound_; a.root = root_;
    a.range_lo = range_lo_; a.n_used = n_used_;
    a.init_state = (init_state == ROOT_STATE) ? root_ : init_state;
    a.out = reinterpret_cast<uint2 *>(ev); a.capacity = SMALL_TEXT_BYTES;      // at most one event per byte
    a.hdr = hdr;
    hdr[1] = a.init_state;
    if (entry_bytes_ == 2) {
        if (range_map_) ac_small_kernel<uint16_t, true><<<1, threads, 0, st>>>(a);
        else ac_small_kernel<uint16_t, false><<<1, threads, 0, st>>>(a);
    } else {
        if (range_map_) ac_small_kernel<uint32_t, true><<<1, threads, 0, st>>>(a);
        else ac_small_kernel<uint32_t, false><<<1, threads, 0, st>>>(a);
    }
    CU_OK(cudaGetLastError());
    CU_OK(cudaStreamSynchronize(st));
    n_events_ = hdr[0];
    end_state_ = hdr[1];
    last_host_events_ = ev;
    last_uniform_len_ = total;
    stats = ACB200_STATS_t{};
    stats.bytes = total; stats.events = n_events_; stats.kernel_launches = 1;
    stats.chunk_bytes = a.chunk; stats.halo_bytes = halo_;
    last_density_ = (double)n_events_ / (double)total;
    return true;
}

bool Engine::scan_host(const char *bytes, const uint64_t *offsets, size_t n, bool first_only,
                       uint32_t init_state, bool events_stay_on_device)
{
    if (device_ < 0) { set_error("automaton has no device table (finalize failed?)"); return false; }
    CU_OK(cudaSetDevice(device_));
    cudaStream_t st = S(stream_);
    const uint64_t total = offsets[n];
    // (events of a findAll=false call are cut to the first one by the caller; tune_chunk forces the general path)
    if (n == 1 && total > 0 && total <= SMALL_TEXT_BYTES && !events_stay_on_device && !tune_chunk && !tune_smem_bytes && tune_filter <= 0)
        return scan_small(bytes, (uint32_t)total, init_state);
    if (total >= MAX_STREAM_BYTES) { set_error("haystack stream exceeds 4 GiB per call"); return false; }
    if (!ensure_text(total + 64)) return false;
    uint32_t uniform_len = 0;
    if (!upload_offsets(offsets, n, &uniform_len)) return false;
    last_uniform_len_ = uniform_len;
    CU_OK(cudaEventRecord(EV(ev_[2]), st));
    if (total) CU_OK(cudaMemcpyAsync(d_text_, bytes, total, cudaMemcpyHostToDevice, st));
    CU_OK(cudaEventRecord(EV(ev_[3]), st));
    if (!launch_scan(d_text_, (uint32_t)total, (uint32_t)std::min<uint64_t>(text_cap_, 0xffffffffu), n, uniform_len, first_only, init_state, nullptr)) return false;
    float ms = 0;
    cudaEventElapsedTime(&ms, EV(ev_[2]), EV(ev_[3]));
    stats.h2d_ms = ms;
    stats.d2h_ms = 0;
    if (n_events_) {
        if (!ensure_host_events(n_events_)) return false;
        CU_OK(cudaEventRecord(EV(ev_[2]), st));
        last_host_events_ = h_events_;
        CU_OK(cudaMemcpyAsync(h_events_, d_events_, n_events_ * sizeof(PackedEvent), cudaMemcpyDeviceToHost, st));
        CU_OK(cudaEventRecord(EV(ev_[3]), st));
        CU_OK(cudaStreamSynchronize(st));
        cudaEventElapsedTime(&ms, EV(ev_[2]), EV(ev_[3]));
        stats.d2h_ms = ms;
    }
    return true;
}

} // namespace acb200
