// waxvs_engine.cu -- host side of libwaxvs_cuda.so: the engine object behind include/wax_vs_cuda.h.
//
// What it mirrors (all /root/reference paths): the state and behaviour of
//   actor MetalVectorEngine            Sources/WaxVectorSearch/MetalVectorEngine.swift:17-893
// with USearchVectorEngine's metric coverage (USearchVectorEngine.swift:44-67) -- the corpus matrix resident
// on the device (here: HBM, row-major fp32), a frameIds side array, a pool of per-search scratch contexts
// (the transient buffer pool, :84-121), upsert/ordered-remove mutation semantics (:330-444), the MV2V
// encoding=2 blob (:682-815) -- re-designed for a discrete 180 GB GPU: id->row hash instead of the O(N)
// firstIndex(of:) scan, explicit pinned staging, one fused kernel launch per query.
//
// There is NO CPU fallback in this file or anywhere in the product path: without a CUDA device every entry
// point that needs one returns WAX_VS_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include <chrono>
#include <random>
#include <thread>
#include <unistd.h>
#include <cmath>
#include <vector>

#include "../../include/wax_vs_cuda.h"
#include "waxvs_common.cuh"
#include "waxvs_scan.cuh"
#include "waxvs_select.cuh"
#include "waxvs_synth.cuh"
#include "waxvs_batch.cuh"

#include <cudaTypedefs.h>

using namespace waxvs;

// ---------------------------------------------------------------------------------------------------------
// errors
static thread_local char g_last_error[512] = "";

static int32_t fail(int32_t code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof g_last_error, fmt, ap);
    va_end(ap);
    return code;
}
#define CUDA_TRY(expr)                                                                            \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return fail(WAX_VS_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));         \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
        if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ---------------------------------------------------------------------------------------------------------
// id -> row: open addressing, linear probing.  Replaces frameIds.firstIndex(of:) (MetalVectorEngine.swift:334,385,426).
struct IdMap {
    std::vector<uint64_t> keys;
    std::vector<uint32_t> vals;
    size_t mask = 0, used = 0;
    static uint64_t mix(uint64_t x) {
        x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31;
        return x;
    }
    void reset(size_t expect) {
        size_t cap = 64;
        while (cap < expect * 2 + 16) cap <<= 1;
        keys.assign(cap, 0);
        vals.assign(cap, 0xFFFFFFFFu);
        mask = cap - 1;
        used = 0;
    }
    void prefetch(uint64_t id) const {
        if (vals.empty()) return;
        const size_t i = mix(id) & mask;
        __builtin_prefetch(&keys[i]);
        __builtin_prefetch(&vals[i]);
    }
    uint32_t find(uint64_t id) const {
        if (vals.empty()) return 0xFFFFFFFFu;
        for (size_t i = mix(id) & mask;; i = (i + 1) & mask) {
            if (vals[i] == 0xFFFFFFFFu) return 0xFFFFFFFFu;
            if (keys[i] == id) return vals[i];
        }
    }
    void put(uint64_t id, uint32_t row) {
        if (vals.empty() || (used + 1) * 2 > keys.size()) grow();
        for (size_t i = mix(id) & mask;; i = (i + 1) & mask) {
            if (vals[i] == 0xFFFFFFFFu) { keys[i] = id; vals[i] = row; ++used; return; }
            if (keys[i] == id) { vals[i] = row; return; }
        }
    }
    void grow() {
        std::vector<uint64_t> ok; std::vector<uint32_t> ov;
        ok.swap(keys); ov.swap(vals);
        reset(std::max<size_t>(used * 2, 32));
        for (size_t i = 0; i < ov.size(); ++i) if (ov[i] != 0xFFFFFFFFu) put(ok[i], ov[i]);
    }
};

// ---------------------------------------------------------------------------------------------------------
struct Tuning {
    int variant = 0;      // 0 auto, 1 TMA-staged, 2 direct LDG
    int rows_per_step = 0;  // 0 auto
    int stages = 0;       // 0 auto
    int warps = 0;        // 0 auto
    int grid = 0;         // 0 = one CTA per SM
    int l2_hint = 0;
    int ldg_ctas_per_sm = 4;
    int chunk_steps = -1;   // dynamic scheduling granularity of the TMA kernel: -1 auto (8 steps, fewer when the corpus gives
                            // each warp only a few steps: 10 K rows = 2 500 steps over 1 184 warps), 0 = static round-robin
    int fused_k_max = 128;  // k <= this stays in the single fused launch (register lists); larger k: emit + radix select
    int batch_tensor = 1;   // 1: batches take the tcgen05 TF32 nominate + exact re-score path when eligible
    int batch_min = 4;      // smallest batch routed to the tensor path
    int time_overlap = 0;   // wax_vs_debug_time_search: alternate consecutive queries over two streams
    int batch_pair = 0;     // 1: cta_group::2 CTA pairs for the SS shapes (validated; no net gain)
    int batch_ts = 0;       // 1: queries in TMEM + CTA pairs (dims <= 384, dims % 128 == 0)
    uint32_t tma_max_dims = 4096;   // generic TMA shape up to this row length, the direct-load kernel above
    int batch_large_k = 1;  // batches with 128 < k <= 1024 take the tensor-core levels (0: loop the single-query emit + select path)
    int batch_heap = 0;     // 0 auto (cost model + adaptive bump), 16 / 24 / 32 / 64: nominee heap size per (slice, query) = kernel shape
    int batch_noinsert = 0; // instrumentation: GEMM pipeline only (results meaningless)
    int batch_bf16 = 1;     // 1: nominate from a bf16 shadow of the corpus when HBM allows (kind::f16 MMAs, 2x the TF32 rate; +dims*2 B/row)
    int batch_ares = 1;     // with batch_bf16: keep the CTA's queries resident in shared memory when they fit (dims <= 512)
    int batch_rescore = 0;  // 0 auto; else nominees re-scored exactly per query (256, 512 or 1024)
    int batch_retry = 1;    // queries level 1 cannot prove go through the filter level (TF32, complete by construction) before an exact scan
    int filter_bf16 = 1;    // unproven queries first get a filter pass over the bf16 shadow (half the bytes of an exact scan)
    int filter_cap = 8192;  // candidates per query the filter level may collect (power of two <= 16384); overflow -> exact scan
    int inline_query = 1;   // host entry points: a query of <= 512 floats travels in the kernel parameters (no H2D copy)
    int host_delivery = 1;  // host entry points: the kernel stores the result in mapped host memory + flag (no D2H copy / sync)
    int tail_select = 1;    // TMA-staged kernels: radix-selection tail instead of pairwise list merges (same results)
    int shard_fused = 1;    // sharded search: exchange + merge inside the scan launch (0: separate 1-CTA launch)
    int single_shadow = 0;  // 1: single queries / batches below batch_min also take the bf16-shadow nominations
                            // (half the HBM bytes per query: 1.19 vs 2.04 ms at 10 M x 384, same results); off by
                            // default: the plain single-query path is the fused fp32 scan BASELINE's north_star names
};

// Per-search scratch: the analogue of TransientBuffers (MetalVectorEngine.swift:36-41, :84-117).
struct SearchCtx {
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    float *d_queries = nullptr; size_t d_queries_cap = 0;        // floats
    wax_vs_candidate *d_out = nullptr; size_t d_out_cap = 0;      // candidates
    float *h_queries = nullptr; size_t h_queries_cap = 0;        // pinned
    wax_vs_candidate *h_out = nullptr; size_t h_out_cap = 0;      // pinned
    uint64_t *d_block_keys = nullptr; size_t block_keys_cap = 0;  // u64
    uint32_t *d_ticket = nullptr;
    uint32_t *d_dist_keys = nullptr; size_t dist_keys_cap = 0;    // large-k path
    SelectState *d_select = nullptr;
    uint64_t *d_sel_keys = nullptr;                               // 16384 u64
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint64_t *d_heaps = nullptr; size_t heaps_cap = 0;            // batched path: nominee heaps
    uint32_t *d_ok = nullptr; size_t ok_cap = 0;                  // batched path: per-query proof flags
    uint32_t *h_ok = nullptr; size_t h_ok_cap = 0;                // pinned
    uint32_t *d_tau = nullptr; size_t tau_cap = 0;                // batched path: shared per-query thresholds
    __nv_bfloat16 *d_queries_bf16 = nullptr; size_t queries_bf16_cap = 0;   // batched bf16 path: converted queries
    float *d_retry_q = nullptr; size_t retry_q_cap = 0;           // bf16 -> TF32 retry: compacted queries
    wax_vs_candidate *d_retry_out = nullptr; size_t retry_out_cap = 0;
    uint32_t *d_retry_ok = nullptr; size_t retry_ok_cap = 0;
    float *d_tau_star = nullptr; size_t tau_star_cap = 0;         // level 1 -> filter level: per-query thresholds
    float *h_tau_star = nullptr; size_t h_tau_star_cap = 0;       // pinned
    float *d_filter_tau = nullptr; size_t filter_tau_cap = 0;     // compacted thresholds of the unproven queries
    float *h_filter_tau = nullptr; size_t h_filter_tau_cap = 0;   // pinned
    uint32_t *d_cand_count = nullptr; size_t cand_count_cap = 0;
    uint32_t *d_cand_rows = nullptr; size_t cand_rows_cap = 0;
    uint64_t *d_cand_keys = nullptr; size_t cand_keys_cap = 0;
    uint32_t *d_mask = nullptr; size_t mask_cap = 0;              // filtered search: row bitset / listed rows
    uint64_t *d_gather_keys = nullptr; size_t gather_cap = 0;     // filtered search: keys of the listed rows
    wax_vs_candidate *d_shard_local = nullptr;                    // sharded search: this rank's list before the exchange [kShardKCap]
    unsigned long long *h_flag = nullptr;                         // mapped pinned: host-delivery completion flag
    unsigned long long host_seq = 0;                              // last value the flag was asked to take
};

struct wax_vs_engine {
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    uint32_t dims = 0;
    uint8_t similarity = 0;

    float *d_corpus = nullptr;
    uint64_t cap_rows = 0, n_rows = 0;

    // frame ids: implicit (id_base + row) after fill_synthetic until the first mutation, else explicit.
    bool ids_identity = true;
    uint64_t id_base = 0;
    std::vector<uint64_t> ids;
    IdMap map;
    bool map_valid = true;
    // Frame ids are handed out in increasing order by the store, so the id array is normally SORTED: then a lookup is a
    // binary search in it and bulk appends touch no hash table at all; the table is built only once an out-of-order id
    // arrives (and kept from then on).  Order-preserving removes and in-place upserts keep the array sorted.
    bool ids_sorted = true;
    uint64_t *d_ids = nullptr; size_t d_ids_cap = 0; bool d_ids_dirty = true;
    std::mutex ids_mu;

    std::shared_mutex rw;  // readers: search / serialize; writer: mutators (AsyncReadWriteLock, :56-80)
    std::mutex pool_mu;
    std::vector<SearchCtx *> pool;
    std::unordered_map<void *, SearchCtx *> stream_ctx;  // wax_vs_search_device: one ctx per caller stream
    uint64_t pool_allocs = 0, pool_reuses = 0;
    Tuning tune;

    // cached per corpus version for the batched path: 1/|v| per row and max |v|
    float *d_inv_norm = nullptr; size_t inv_norm_cap = 0;
    uint32_t *d_max_norm = nullptr;
    uint64_t norms_rows = 0;       // rows [0, norms_rows) of d_inv_norm are valid (appends extend it, other mutations reset it)
    std::mutex norms_mu;
    // bf16 shadow of the corpus for the batched bf16 nominations (cached per corpus version, guarded by norms_mu)
    __nv_bfloat16 *d_shadow = nullptr; size_t shadow_cap = 0;
    uint64_t shadow_rows = 0;      // rows [0, shadow_rows) of d_shadow are valid; shadow_valid = covers every live row
    bool shadow_valid = false, shadow_unavailable = false;
    uint64_t batch_tensor_queries = 0, batch_fallback_queries = 0;   // instrumentation
    uint64_t batch_bf16_queries = 0, batch_retry_queries = 0, batch_tf32_queries = 0, batch_filter_bf16_queries = 0;
    // Adaptive level choice: when more than a quarter of a batch fails the coarse bf16 bound (tightly clustered
    // neighbours), the next 16 batches nominate in TF32 straight away, then bf16 is probed again.
    uint32_t bf16_skip_batches = 0;
    // adaptive nominee-heap size (bf16 level 1): a batch that left queries unproven makes the next `heap_bump_ttl` batches
    // use one size more than the model picks; a failure right after probing back down doubles the time-out
    uint32_t heap_bump = 0, heap_bump_ttl = 0, heap_backoff = 256;
    bool heap_probing = false;
    uint32_t last_heap = 0;
    // Bulk ingest / export staging: two pinned buffers so that the host-side copy of chunk i+1 overlaps
    // the DMA of chunk i, one copy stream, a device staging area for upserts and for the compaction of removes.
    struct Ingest {
        cudaStream_t stream = nullptr;
        cudaEvent_t ev[2] = {nullptr, nullptr};
        uint8_t *pin[2] = {nullptr, nullptr};
        size_t pin_bytes = 0;
        float *d_stage = nullptr; size_t d_stage_cap = 0;        // floats
        uint32_t *d_index = nullptr; size_t d_index_cap = 0;     // u32
        int threads = 1;
    } ing;
    uint64_t ingest_h2d_bytes = 0, ingest_d2h_bytes = 0;         // instrumentation
    std::mutex ingest_mu;                                        // readers that use the staging (serialize)
    std::mutex attr_mu;            // cudaFuncSetAttribute bookkeeping (per engine = per device)
    bool sort_attr_set = false, gather_attr_set = false, batch_attr_set = false;
    std::unordered_map<const void *, int> smem_granted;   // opt-in shared memory already granted, per kernel (attr_mu)
    // Device-path searches (wax_vs_search_device & co.) return while their kernels are still in flight on the
    // caller's stream.  Mutators must not touch the corpus under them: every mutator drains the device first when
    // this flag says something was enqueued since the last drain.
    std::atomic<bool> async_pending{false};
    unsigned long long *debug_trace = nullptr;   // wax_vs_debug_phase_trace: device buffer the scan kernels stamp (else nullptr)

    // Row-sharded search (wax_vs_shard_*): this engine is rank `rank` of `world`; box[r] = rank r's mailbox.
    struct Shard {
        bool open = false, connected = false;
        int rank = 0, world = 0;
        uint64_t row_offset = 0;
        ShardMailbox *box[kShardMaxRanks] = {};
        bool ipc[kShardMaxRanks] = {};            // mapped with cudaIpcOpenMemHandle (closed in shard_close)
        unsigned long long seq = 0;               // collective calls issued so far (same on every rank)
        unsigned long long timeout_ns = 20ull * 1000 * 1000 * 1000;
        std::mutex mu;                            // one collective call at a time per rank: seq order = issue order
        SearchCtx *ctx = nullptr;                 // host entry point: stream + scratch
        wax_vs_candidate *d_final = nullptr;      // [kShardKCap]
        wax_vs_candidate *h_final = nullptr;      // mapped pinned [kShardKCap]: the kernel writes the merged result here
        unsigned long long *h_flag = nullptr;     // mapped pinned: seq when h_final is complete
    } shard;
};

extern "C" { static void shard_teardown(wax_vs_engine *e, bool free_own); static void ingest_free(wax_vs_engine *e); }

// Called by every mutator after it has taken the write lock (and selected the device).
static void drain_device_path(wax_vs_engine *e) {
    if (e->async_pending.exchange(false)) cudaDeviceSynchronize();
}
// Derived per-row caches (1/|v|, bf16 shadow) after a mutation.  keep_prefix: rows [0, keep_prefix) are untouched (a pure
// append keeps everything it had); 0 = rebuild from scratch on the next batched search.
static void invalidate_row_caches(wax_vs_engine *e, uint64_t keep_prefix) {
    e->norms_rows = std::min(e->norms_rows, keep_prefix);
    e->shadow_rows = std::min(e->shadow_rows, keep_prefix);
    if (e->shadow_rows == 0) e->shadow_valid = false;
}

// ---------------------------------------------------------------------------------------------------------
// scratch contexts
static void ctx_free(SearchCtx *c) {
    if (!c) return;
    if (c->d_queries) cudaFree(c->d_queries);
    if (c->d_out) cudaFree(c->d_out);
    if (c->h_queries) cudaFreeHost(c->h_queries);
    if (c->h_out) cudaFreeHost(c->h_out);
    if (c->d_block_keys) cudaFree(c->d_block_keys);
    if (c->d_ticket) cudaFree(c->d_ticket);
    if (c->d_dist_keys) cudaFree(c->d_dist_keys);
    if (c->d_select) cudaFree(c->d_select);
    if (c->d_sel_keys) cudaFree(c->d_sel_keys);
    if (c->d_heaps) cudaFree(c->d_heaps);
    if (c->d_ok) cudaFree(c->d_ok);
    if (c->d_tau) cudaFree(c->d_tau);
    if (c->d_queries_bf16) cudaFree(c->d_queries_bf16);
    if (c->d_retry_q) cudaFree(c->d_retry_q);
    if (c->d_retry_out) cudaFree(c->d_retry_out);
    if (c->d_retry_ok) cudaFree(c->d_retry_ok);
    if (c->d_tau_star) cudaFree(c->d_tau_star);
    if (c->h_tau_star) cudaFreeHost(c->h_tau_star);
    if (c->d_filter_tau) cudaFree(c->d_filter_tau);
    if (c->h_filter_tau) cudaFreeHost(c->h_filter_tau);
    if (c->d_cand_count) cudaFree(c->d_cand_count);
    if (c->d_cand_rows) cudaFree(c->d_cand_rows);
    if (c->d_cand_keys) cudaFree(c->d_cand_keys);
    if (c->d_mask) cudaFree(c->d_mask);
    if (c->d_gather_keys) cudaFree(c->d_gather_keys);
    if (c->d_shard_local) cudaFree(c->d_shard_local);
    if (c->h_flag) cudaFreeHost(c->h_flag);
    if (c->h_ok) cudaFreeHost(c->h_ok);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

static int32_t ctx_new(wax_vs_engine *e, SearchCtx **out, bool with_stream) {
    SearchCtx *c = new (std::nothrow) SearchCtx();
    if (!c) return fail(WAX_VS_ERR_CUDA, "out of host memory");
    auto bail = [&](int32_t rc) { ctx_free(c); return rc; };
    if (with_stream) {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess)
            return bail(fail(WAX_VS_ERR_CUDA, "cudaStreamCreate failed"));
        c->own_stream = true;
    }
    c->block_keys_cap = static_cast<size_t>(std::max(e->sm_count * 8, 2048)) * 128;
    if (cudaMalloc(&c->d_block_keys, c->block_keys_cap * sizeof(uint64_t)) != cudaSuccess ||
        cudaMalloc(&c->d_ticket, 4 * sizeof(uint32_t)) != cudaSuccess ||      // [0] ticket, [1] work counter
        cudaMemset(c->d_ticket, 0, 4 * sizeof(uint32_t)) != cudaSuccess ||
        cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess)
        return bail(fail(WAX_VS_ERR_CUDA, "failed to allocate search scratch: %s",
                         cudaGetErrorString(cudaGetLastError())));
    *out = c;
    return WAX_VS_OK;
}

static int32_t ctx_acquire(wax_vs_engine *e, SearchCtx **out) {
    {
        std::lock_guard<std::mutex> g(e->pool_mu);
        if (!e->pool.empty()) {
            *out = e->pool.back();
            e->pool.pop_back();
            ++e->pool_reuses;
            return WAX_VS_OK;
        }
        ++e->pool_allocs;
    }
    return ctx_new(e, out, true);
}
static void ctx_release(wax_vs_engine *e, SearchCtx *c) {
    std::lock_guard<std::mutex> g(e->pool_mu);
    e->pool.push_back(c);
}
// The scratch context bound to a caller-owned stream (device-path entry points): find-or-create in ONE critical
// section, so two threads that first use the same stream cannot both insert (and leak) a context.
static int32_t ctx_for_stream(wax_vs_engine *e, void *cuda_stream, SearchCtx **out) {
    std::lock_guard<std::mutex> pg(e->pool_mu);
    auto it = e->stream_ctx.find(cuda_stream);
    if (it != e->stream_ctx.end()) { *out = it->second; return WAX_VS_OK; }
    SearchCtx *c = nullptr;
    int32_t rc = ctx_new(e, &c, false);
    if (rc) return rc;
    c->stream = static_cast<cudaStream_t>(cuda_stream);
    e->stream_ctx[cuda_stream] = c;
    ++e->pool_allocs;
    *out = c;
    return WAX_VS_OK;
}

template <typename T>
static int32_t ensure_dev(T **p, size_t *cap, size_t need, const char *what) {
    if (need <= *cap) return WAX_VS_OK;
    if (*p) { cudaFree(*p); *p = nullptr; *cap = 0; }
    if (cudaMalloc(p, need * sizeof(T)) != cudaSuccess)
        return fail(WAX_VS_ERR_CUDA, "failed to allocate %s (%zu bytes): %s", what, need * sizeof(T),
                    cudaGetErrorString(cudaGetLastError()));
    *cap = need;
    return WAX_VS_OK;
}
template <typename T>
static int32_t ensure_pinned(T **p, size_t *cap, size_t need, const char *what) {
    if (need <= *cap) return WAX_VS_OK;
    if (*p) { cudaFreeHost(*p); *p = nullptr; *cap = 0; }
    // mapped + portable: kernels may write results straight into it (host delivery), any device may use it
    if (cudaHostAlloc(reinterpret_cast<void **>(p), need * sizeof(T), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess)
        return fail(WAX_VS_ERR_CUDA, "failed to allocate pinned %s (%zu bytes): %s", what, need * sizeof(T),
                    cudaGetErrorString(cudaGetLastError()));
    *cap = need;
    return WAX_VS_OK;
}

// ---------------------------------------------------------------------------------------------------------
// kernel dispatch
struct TmaConfig { int C, R, warps, stages; size_t smem; };

static bool pick_tma_config(const wax_vs_engine *e, TmaConfig *cfg, int mode = 0) {
    const uint32_t d = e->dims;
    if (d % 4u != 0) return false;                      // rows must be 16-byte multiples for the bulk copy
    const size_t budget = (e->smem_optin ? e->smem_optin : 232448) - 4096;   // minus the kernels' static shared memory
    int C = 0;
    if (d % 128u == 0) {
        const int c = static_cast<int>(d / 128u);
        if (c == 1 || c == 2 || c == 3 || c == 4 || c == 6 || c == 8 || c == 12) C = c;   // unrolled shapes (12: the 1536-dim embeddings)
    }
    if (C == 0 && d < 32) return false;                  // a few floats per row: the direct-load kernel
    if (C == 0 && d > e->tune.tma_max_dims) return false;   // very long rows: too few warps fit beside two stages; the
                                                             // direct-load kernel streams them at 6.9 TB/s (sweep_long_r02t)
    // rows per step / warps per CTA by row length (profiles/dims_sweep_r01_call17.json): keep a step at >= 4-12 KB
    // and give short rows more warps (their bound is per-row instruction latency, not bytes in flight)
    int R, warps_default = 8;
    if (C == 0) {                                        // generic shape: run-time chunk count, query in shared memory
        // keep a step at >= 2-8 KB: short generic rows (dims < 128, 160, 300, 400, ...) take 8 or 4 rows per step and
        // more warps, like the unrolled C <= 2 shapes (profiles/small_dims_sweep_r02*.jsonl)
        // (long rows: as many rows as keep a step at <= 32 KB -- the few warps that then fit still hold ~190 KB in flight,
        // profiles/small_dims_sweep_r02h.jsonl / sweep_big_r02i.jsonl: 1000 dims 6.1 -> 7.3 TB/s, 2048 dims 6.5 -> 7.3)
        int auto_r = d > 256 ? 4 : 8;
        if (d > 640) { auto_r = 8; while (auto_r > 1 && static_cast<size_t>(auto_r) * d * 4 > 32768) auto_r >>= 1; }
        if (d > 3072) auto_r = 1;       // 12-16 KB rows: one per step keeps six warps in flight (3584: 7.36, 4096: 7.24 TB/s)
        const int want_r = e->tune.rows_per_step;
        R = (want_r == 1 || want_r == 2 || want_r == 4 || want_r == 8) ? want_r : auto_r;
        if (R == 8) warps_default = 16;
        else if (R == 4) warps_default = 12;
    } else if (C == 12) {
        R = e->tune.rows_per_step == 1 || e->tune.rows_per_step == 2 ? e->tune.rows_per_step : 2;
    } else if (C >= 6) {
        R = e->tune.rows_per_step == 2 || e->tune.rows_per_step == 4 ? e->tune.rows_per_step : 2;
    } else {
        R = e->tune.rows_per_step == 4 || e->tune.rows_per_step == 8 ? e->tune.rows_per_step : (C <= 2 ? 8 : 4);
        if (C == 1) warps_default = 16;
        else if (C == 2) warps_default = 12;
        // wide lists (33 <= k <= 128, four keys per lane): 16 warps per CTA spread the list upkeep and beat 8 warps up to
        // a few GB of corpus (profiles/shape_sweep_r02.jsonl, k = 72: 72 vs 87 us at 174 K rows, 296 vs 315 us at
        // 1.25 M, 555 vs 577 at 2.5 M; at 10 M rows the 8-warp shape's 48 KB in flight wins again)
        else if (mode == 1 && static_cast<uint64_t>(e->n_rows) * d * 4 < (8ull << 30)) warps_default = 16;
    }
    // Default ring depth 2: measured best on B200 (profiles/sweep_r01_call2.json: ~48 KB in flight per SM beats
    // deeper rings by 5-10 %).
    const int stages = e->tune.stages > 0 ? e->tune.stages : 2;
    const size_t stage_bytes = static_cast<size_t>(R) * d * 4;
    const size_t query_bytes = C == 0 ? (static_cast<size_t>(d) * 4 + 512 + 32) : 0;
    auto smem_for = [&](int w) { return static_cast<size_t>(w) * stages * (stage_bytes + 8 + 4) + static_cast<size_t>(w) * 1024 + 16 + query_bytes; };
    int warps = e->tune.warps ? e->tune.warps : warps_default;
    warps = std::max(1, std::min(16, warps));
    if (!e->tune.warps) while (warps > 2 && smem_for(warps) > budget) --warps;   // long rows: fewer warps per CTA
    if (smem_for(warps) > budget) return false;
    cfg->C = C; cfg->R = R; cfg->warps = warps; cfg->stages = stages; cfg->smem = smem_for(warps);
    return true;
}

// The opt-in shared-memory limit is per function and per device: set it once per engine (= per device), and again
// only if a larger ring is requested, instead of on every launch -- it costs more host time than a 10 K-row scan.
template <typename K>
static cudaError_t grant_smem(wax_vs_engine *e, K kernel, size_t bytes) {
    std::lock_guard<std::mutex> g(e->attr_mu);
    int &have = e->smem_granted[reinterpret_cast<const void *>(kernel)];
    if (have >= static_cast<int>(bytes)) return cudaSuccess;
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
    if (err == cudaSuccess) have = static_cast<int>(bytes);
    return err;
}

template <int C, int R, int M, int E, bool EMIT>
static cudaError_t launch_tma_inst(wax_vs_engine *e, const ScanParams &p, int grid, const TmaConfig &cfg, cudaStream_t s) {
    cudaError_t err = grant_smem(e, scan_tma_kernel<C, R, M, E, EMIT>, cfg.smem);
    if (err != cudaSuccess) return err;
    scan_tma_kernel<C, R, M, E, EMIT><<<grid, cfg.warps * 32, cfg.smem, s>>>(p);
    return cudaGetLastError();
}
// mode: 0 = fused list k <= 32, 1 = fused list k <= 128, 2 = emit distance keys
template <int C, int R>
static cudaError_t launch_tma_cr(wax_vs_engine *e, const ScanParams &p, int grid, const TmaConfig &cfg, int metric, int mode,
                                 cudaStream_t s) {
    switch (metric * 3 + mode) {
        case 0: return launch_tma_inst<C, R, kCosine, 1, false>(e, p, grid, cfg, s);
        case 1: return launch_tma_inst<C, R, kCosine, 4, false>(e, p, grid, cfg, s);
        case 2: return launch_tma_inst<C, R, kCosine, 1, true>(e, p, grid, cfg, s);
        case 3: return launch_tma_inst<C, R, kDot, 1, false>(e, p, grid, cfg, s);
        case 4: return launch_tma_inst<C, R, kDot, 4, false>(e, p, grid, cfg, s);
        case 5: return launch_tma_inst<C, R, kDot, 1, true>(e, p, grid, cfg, s);
        case 6: return launch_tma_inst<C, R, kL2, 1, false>(e, p, grid, cfg, s);
        case 7: return launch_tma_inst<C, R, kL2, 4, false>(e, p, grid, cfg, s);
        default: return launch_tma_inst<C, R, kL2, 1, true>(e, p, grid, cfg, s);
    }
}
static cudaError_t launch_tma(wax_vs_engine *e, const ScanParams &p, int grid, const TmaConfig &cfg, int metric, int mode,
                              cudaStream_t s) {
#define WAXVS_CASE(Cv, Rv) if (cfg.C == Cv && cfg.R == Rv) return launch_tma_cr<Cv, Rv>(e, p, grid, cfg, metric, mode, s)
    WAXVS_CASE(1, 4); WAXVS_CASE(1, 8); WAXVS_CASE(2, 4); WAXVS_CASE(2, 8);
    WAXVS_CASE(3, 4); WAXVS_CASE(3, 8); WAXVS_CASE(4, 4); WAXVS_CASE(4, 8);
    WAXVS_CASE(6, 2); WAXVS_CASE(6, 4); WAXVS_CASE(8, 2); WAXVS_CASE(8, 4);
    WAXVS_CASE(12, 1); WAXVS_CASE(12, 2);
    WAXVS_CASE(0, 1); WAXVS_CASE(0, 2); WAXVS_CASE(0, 4); WAXVS_CASE(0, 8);
#undef WAXVS_CASE
    return cudaErrorInvalidValue;
}
static cudaError_t launch_ldg(const ScanParams &p, int grid, int metric, int mode, cudaStream_t s) {
    switch (metric * 3 + mode) {
        case 0: scan_ldg_kernel<kCosine, 1, false><<<grid, 256, 0, s>>>(p); break;
        case 1: scan_ldg_kernel<kCosine, 4, false><<<grid, 256, 0, s>>>(p); break;
        case 2: scan_ldg_kernel<kCosine, 1, true><<<grid, 256, 0, s>>>(p); break;
        case 3: scan_ldg_kernel<kDot, 1, false><<<grid, 256, 0, s>>>(p); break;
        case 4: scan_ldg_kernel<kDot, 4, false><<<grid, 256, 0, s>>>(p); break;
        case 5: scan_ldg_kernel<kDot, 1, true><<<grid, 256, 0, s>>>(p); break;
        case 6: scan_ldg_kernel<kL2, 1, false><<<grid, 256, 0, s>>>(p); break;
        case 7: scan_ldg_kernel<kL2, 4, false><<<grid, 256, 0, s>>>(p); break;
        default: scan_ldg_kernel<kL2, 1, true><<<grid, 256, 0, s>>>(p); break;
    }
    return cudaGetLastError();
}

// Enqueue one query's scan + top-k on `stream`.  k_eff <= 10000.  Adds the number of kernels launched.
// Wait for a kernel's host-visible completion flag (mapped pinned memory): the result is usable a few microseconds
// after the kernel stored it, without an event / stream synchronisation.  The stream is polled now and then so that a
// launch failure surfaces.  Returns WAX_VS_OK, 1 when the flag carries the error bit, or a negative code.
static int32_t wait_host_flag(cudaStream_t stream, unsigned long long *flag_ptr, unsigned long long seq,
                              unsigned long long timeout_ns) {
    volatile unsigned long long *flag = flag_ptr;
    const auto t0 = std::chrono::steady_clock::now();
    uint32_t spins = 0;
    for (;;) {
        const unsigned long long v = *flag;
        if ((v & ~kShardErrorBit) == seq) {
            std::atomic_thread_fence(std::memory_order_acquire);
            return (v & kShardErrorBit) ? 1 : WAX_VS_OK;
        }
        if ((++spins & 0x3FFu) == 0) {
            const cudaError_t q = cudaStreamQuery(stream);
            if (q != cudaSuccess && q != cudaErrorNotReady)
                return fail(WAX_VS_ERR_CUDA, "search failed on the device: %s", cudaGetErrorString(q));
            if (q == cudaSuccess && ((*flag) & ~kShardErrorBit) != seq)
                return fail(WAX_VS_ERR_CUDA, "search finished without publishing its result");
            if (std::chrono::steady_clock::now() - t0 > std::chrono::nanoseconds(timeout_ns) + std::chrono::seconds(5))
                return fail(WAX_VS_ERR_CUDA, "search did not complete");
        }
    }
}

// Host delivery (the synchronous host entry points, fused top-k only): h_query travels in the kernel parameters when the
// kernel can take it (else it is copied H2D here), and the kernel stores the result into host_out + raises host_flag.
struct HostDelivery {
    const float *h_query;               // host query (dims floats); d_query is ignored when set
    wax_vs_candidate *host_out;         // mapped pinned [k], or nullptr
    unsigned long long *host_flag;      // mapped pinned
    unsigned long long seq;
    bool delivered = false;             // out: the launched kernel will raise host_flag (else: copy d_out back yourself)
};

// `shard` (optional): the row-sharded form -- d_out receives the result MERGED over all ranks; the exchange runs inside
// the scan launch when the kernel's shared-memory lists can hold the merge keys, else as one extra 1-CTA launch.
static int32_t enqueue_search(wax_vs_engine *e, SearchCtx *c, const float *d_query, uint32_t k_eff,
                              uint64_t row_offset, wax_vs_candidate *d_out, const uint64_t *d_ids,
                              cudaStream_t stream, uint64_t *launches, const uint32_t *d_mask = nullptr,
                              const ShardParams *shard = nullptr, const HostDelivery *host = nullptr) {
    wax_vs_candidate *d_merged = nullptr;
    if (shard) {
        if (k_eff > static_cast<uint32_t>(kShardKCap))
            return fail(WAX_VS_ERR_UNSUPPORTED, "sharded search supports top_k <= %d (got %u)", kShardKCap, k_eff);
        if (!c->d_shard_local) CUDA_TRY(cudaMalloc(&c->d_shard_local, kShardKCap * sizeof(wax_vs_candidate)));
        d_merged = d_out;
        d_out = c->d_shard_local;           // the scan produces the LOCAL list; the exchange writes d_merged
    }
    auto exchange_standalone = [&]() -> int32_t {
        ShardParams sp = *shard;
        sp.final_out = d_merged;
        shard_exchange_kernel<<<1, 256, 0, stream>>>(sp, d_out, k_eff);
        CUDA_TRY(cudaGetLastError());
        ++*launches;
        return WAX_VS_OK;
    };
    auto stage_query = [&]() -> int32_t {            // host query -> pinned staging -> device, on `stream`
        int32_t rc = ensure_dev(&c->d_queries, &c->d_queries_cap, static_cast<size_t>(e->dims), "query buffer");
        if (!rc) rc = ensure_pinned(&c->h_queries, &c->h_queries_cap, static_cast<size_t>(e->dims), "query staging");
        if (rc) return rc;
        memcpy(c->h_queries, host->h_query, e->dims * sizeof(float));
        CUDA_TRY(cudaMemcpyAsync(c->d_queries, c->h_queries, e->dims * sizeof(float), cudaMemcpyHostToDevice, stream));
        d_query = c->d_queries;
        return WAX_VS_OK;
    };
    if (e->n_rows == 0) {
        CUDA_TRY(cudaMemsetAsync(d_out, 0, static_cast<size_t>(k_eff) * sizeof(wax_vs_candidate), stream));
        return shard ? exchange_standalone() : WAX_VS_OK;
    }
    ScanParams p{};
    p.corpus = e->d_corpus; p.query = d_query;
    p.n_rows = static_cast<uint32_t>(e->n_rows); p.dims = e->dims; p.k = k_eff;
    p.block_keys = c->d_block_keys; p.ticket = c->d_ticket; p.out = d_out;
    p.frame_ids = d_ids; p.id_base = e->id_base; p.row_offset = row_offset;
    p.use_l2_hint = e->tune.l2_hint ? 1u : 0u;
    p.chunk_steps = e->tune.chunk_steps > 0 ? static_cast<uint32_t>(e->tune.chunk_steps) : 0u;   // auto: set below
    p.work_counter = c->d_ticket + 1;
    p.mask = d_mask;
    p.trace = e->debug_trace;

    const bool emit = k_eff > static_cast<uint32_t>(e->tune.fused_k_max);
    const int mode = emit ? 2 : (k_eff <= 32 ? 0 : 1);
    if (emit) {
        int32_t rc = ensure_dev(&c->d_dist_keys, &c->dist_keys_cap, static_cast<size_t>(e->n_rows), "distance keys");
        if (rc) return rc;
        if (!c->d_select) CUDA_TRY(cudaMalloc(&c->d_select, sizeof(SelectState)));
        if (!c->d_sel_keys) CUDA_TRY(cudaMalloc(&c->d_sel_keys, 16384 * sizeof(uint64_t)));
        p.dist_keys = c->d_dist_keys;
    }

    TmaConfig cfg{};
    bool use_tma = (e->tune.variant != 2) && pick_tma_config(e, &cfg, mode);
    if (e->tune.variant == 1 && !use_tma)
        return fail(WAX_VS_ERR_UNSUPPORTED, "TMA-staged kernel does not support dims=%u", e->dims);
    if (host && host->h_query) {
        const bool inline_ok = use_tma && !emit && e->tune.inline_query != 0 && e->dims <= static_cast<uint32_t>(kInlineQueryFloats);
        if (inline_ok) {
            memcpy(p.query_inline, host->h_query, e->dims * sizeof(float));
            p.query = nullptr;
        } else {
            int32_t rc = stage_query();
            if (rc) return rc;
            p.query = d_query;
        }
    }
    if (host && host->host_out && !emit && !shard) {
        p.host_out = host->host_out; p.host_flag = host->host_flag; p.host_seq = host->seq;
        const_cast<HostDelivery *>(host)->delivered = true;
    }
    if (use_tma && !emit && e->tune.tail_select) {   // selection tail: the idle ring is its staging area
        p.tail_select = 1u;
        p.tail_smem_bytes = static_cast<uint32_t>(static_cast<size_t>(cfg.warps) * cfg.stages * cfg.R * e->dims * sizeof(float));
    }
    bool fused_exchange = false;
    if (shard && !emit) {     // the merge keys (world * k uint32) live in the kernel's block-list shared memory
        const size_t list_bytes = p.tail_select ? p.tail_smem_bytes
                                                : static_cast<size_t>(use_tma ? cfg.warps : 8) * 32 * (mode == 0 ? 1 : 4) * sizeof(uint64_t);
        fused_exchange = e->tune.shard_fused != 0 && static_cast<size_t>(shard->world) * k_eff * sizeof(uint32_t) <= list_bytes;
        if (fused_exchange) { p.shard = *shard; p.shard.final_out = d_merged; }
    }
    int grid;
    const int grid_cap = static_cast<int>(c->block_keys_cap / 128);
    if (use_tma) {
        p.stages = static_cast<uint32_t>(cfg.stages);
        const uint64_t steps = (e->n_rows + cfg.R - 1) / cfg.R;
        const int max_grid = e->tune.grid > 0 ? e->tune.grid : e->sm_count;
        grid = static_cast<int>(std::min<uint64_t>(max_grid, (steps + cfg.warps - 1) / cfg.warps));
        grid = std::max(std::min(grid, grid_cap), 1);
        if (e->tune.chunk_steps < 0)     // auto: about two claims per warp at least (profiles/small_n_r02.jsonl), at most 8 steps
            p.chunk_steps = static_cast<uint32_t>(std::max<uint64_t>(1, std::min<uint64_t>(8, steps / (static_cast<uint64_t>(grid) * cfg.warps * 2))));
        CUDA_TRY(launch_tma(e, p, grid, cfg, e->similarity, mode, stream));
    } else {
        const int max_grid = e->tune.grid > 0 ? e->tune.grid : e->sm_count * e->tune.ldg_ctas_per_sm;
        grid = static_cast<int>(std::min<uint64_t>(max_grid, (e->n_rows + 7) / 8));
        grid = std::max(std::min(grid, grid_cap), 1);
        CUDA_TRY(launch_ldg(p, grid, e->similarity, mode, stream));
    }
    ++*launches;

    if (emit) {
        const uint32_t n = static_cast<uint32_t>(e->n_rows);
        const int sgrid = std::max(1, std::min<int>(e->sm_count * 4, static_cast<int>((n + 511) / 512)));
        select_init_kernel<<<1, 256, 0, stream>>>(c->d_select, k_eff);
        for (int pass = 0; pass < kSelectPasses; ++pass) {
            select_hist_kernel<<<sgrid, 512, 0, stream>>>(c->d_dist_keys, n, c->d_select, pass);
            select_scan_kernel<<<1, 1024, 0, stream>>>(c->d_select, pass);
        }
        select_compact_kernel<<<sgrid, 512, 0, stream>>>(c->d_dist_keys, n, c->d_select, c->d_sel_keys, 16384);
        uint32_t pow2 = 64;
        while (pow2 < k_eff) pow2 <<= 1;
        {   // opt-in shared-memory limits are per function AND per device: once per engine, not once per process
            std::lock_guard<std::mutex> ag(e->attr_mu);
            if (!e->sort_attr_set) {
                CUDA_TRY(cudaFuncSetAttribute(select_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
                e->sort_attr_set = true;
            }
        }
        select_sort_kernel<<<1, 1024, pow2 * sizeof(uint64_t), stream>>>(c->d_select, c->d_sel_keys, pow2, p);
        CUDA_TRY(cudaGetLastError());
        *launches += 3 + 2 * kSelectPasses;
    }
    if (shard && !fused_exchange) return exchange_standalone();
    return WAX_VS_OK;
}

// ---------------------------------------------------------------------------------------------------------
// batched path: tcgen05 TF32 nomination + exact re-score (waxvs_batch.cuh)
static PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    });
    return fn;
}

// Row-major [rows][dims] matrix (fp32, or bf16 for the shadow path), box = box_rows x 128 bytes (32 floats / 64 bf16),
// 128-byte swizzle, OOB -> zeros.
static int32_t make_tensor_map(CUtensorMap *map, const void *base, uint64_t rows, uint32_t dims, uint32_t box_rows,
                               bool bf16 = false) {
    auto enc = tensor_map_encoder();
    if (!enc) return fail(WAX_VS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const size_t esize = bf16 ? sizeof(__nv_bfloat16) : sizeof(float);
    const cuuint64_t gdim[2] = {dims, rows};
    const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(dims) * esize};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(bf16 ? kBatchKBlockBf16 : kBatchKBlock), box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                           const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(WAX_VS_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(r));
    return WAX_VS_OK;
}

static bool batch_bf16_wanted(const wax_vs_engine *e);
static bool batch_tensor_eligible(const wax_vs_engine *e, uint32_t n_queries, uint32_t k_eff) {
    const uint32_t min_batch = (e->tune.single_shadow && batch_bf16_wanted(e)) ? 1u : static_cast<uint32_t>(std::max(e->tune.batch_min, 1));
    return e->tune.batch_tensor && n_queries >= min_batch &&
           (e->similarity == WAX_VS_COSINE || e->similarity == WAX_VS_DOT) && e->dims % kBatchKBlock == 0 &&
           e->dims <= 8192 &&        // the proof's accumulation slack (dims * 2^-23) stays far below the operand bound
           k_eff >= 1 && e->n_rows >= 1 &&
           // 128 < k <= 1024 (the production candidate limit reaches 1 000, UnifiedSearch.swift:1195-1200): real batches
           // only.  Level 1 can rarely PROVE such a k (64 nominees per slice barely cover it) but its exactly re-scored
           // nominees give the filter level its threshold, and that level is complete by construction.
           (k_eff <= 128 || (k_eff <= static_cast<uint32_t>(kBatchRescoreMax) && e->tune.batch_large_k &&
                             n_queries >= static_cast<uint32_t>(std::max(e->tune.batch_min, 4)) &&
                             e->n_rows >= 64ull * k_eff));   // smaller corpora: too few row slices to nominate k rows
}

// 1/|v| per row + max |v|, cached per corpus version.  Appends only extend the cache (rows [norms_rows, n_rows) are
// computed, the running max only grows); anything that moves or overwrites rows resets norms_rows to 0.
static int32_t ensure_norms_locked(wax_vs_engine *e, cudaStream_t stream) {
    if (e->norms_rows == e->n_rows && e->d_inv_norm) return WAX_VS_OK;
    if (static_cast<size_t>(e->n_rows) > e->inv_norm_cap || !e->d_inv_norm) {
        e->norms_rows = 0;                                       // ensure_dev re-allocates: the cached prefix is gone
        const size_t want = static_cast<size_t>(std::max<uint64_t>(e->cap_rows, std::max<uint64_t>(e->n_rows, 1)));
        int32_t rc = ensure_dev(&e->d_inv_norm, &e->inv_norm_cap, want, "row norms");
        if (rc) return rc;
    }
    if (!e->d_max_norm) { CUDA_TRY(cudaMalloc(&e->d_max_norm, sizeof(uint32_t))); e->norms_rows = 0; }
    if (e->norms_rows > e->n_rows) e->norms_rows = 0;
    if (e->norms_rows == 0) CUDA_TRY(cudaMemsetAsync(e->d_max_norm, 0, sizeof(uint32_t), stream));
    const uint64_t first = e->norms_rows, count = e->n_rows - first;
    if (count) {
        const int grid = static_cast<int>(std::min<uint64_t>(static_cast<uint64_t>(e->sm_count) * 8, (count + 7) / 8));
        row_norms_kernel<<<std::max(grid, 1), 256, 0, stream>>>(e->d_corpus + first * e->dims, static_cast<uint32_t>(count), e->dims,
                                                                e->d_inv_norm + first, e->d_max_norm);
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaStreamSynchronize(stream));
    e->norms_rows = e->n_rows;
    return WAX_VS_OK;
}
static int32_t ensure_norms(wax_vs_engine *e, cudaStream_t stream) {
    std::lock_guard<std::mutex> g(e->norms_mu);
    return ensure_norms_locked(e, stream);
}

// The bf16 nominations want dims % 64 == 0 (whole 128-byte k-blocks of bf16).
static bool batch_bf16_wanted(const wax_vs_engine *e) {
    return e->tune.batch_bf16 != 0 && e->dims % kBatchKBlockBf16 == 0 && !e->shadow_unavailable;
}

// bf16 shadow of the corpus (cosine: rows pre-scaled by 1/|v|), cached per corpus version and extended incrementally
// by appends like the norms.  Returns WAX_VS_OK with e->shadow_valid == false when the extra dims*2 bytes per row do
// not fit in HBM (the caller then nominates in TF32 from the fp32 corpus; counter "shadow_unavailable").
static int32_t ensure_shadow(wax_vs_engine *e, cudaStream_t stream) {
    std::lock_guard<std::mutex> g(e->norms_mu);
    if ((e->shadow_valid && e->shadow_rows == e->n_rows) || e->shadow_unavailable) return WAX_VS_OK;
    int32_t rc = ensure_norms_locked(e, stream);
    if (rc) return rc;
    const size_t need = static_cast<size_t>(e->n_rows) * e->dims;                                  // must hold
    const size_t pref = static_cast<size_t>(std::max<uint64_t>(e->cap_rows, e->n_rows)) * e->dims;   // would like
    if (e->shadow_cap < need) {
        if (e->d_shadow) { cudaFree(e->d_shadow); e->d_shadow = nullptr; e->shadow_cap = 0; }
        e->shadow_rows = 0; e->shadow_valid = false;
        size_t free_b = 0, total_b = 0;
        size_t want = pref;
        size_t bytes = want * sizeof(__nv_bfloat16);
        // keep headroom for scratch and growth: the shadow must leave max(2 GiB, 10 % of the device) free.  Sized for
        // the corpus CAPACITY so that appends extend it in place; if only the live rows fit, take that.
        const bool info = cudaMemGetInfo(&free_b, &total_b) == cudaSuccess;
        const size_t headroom = std::max<size_t>(size_t(2) << 30, total_b / 10);
        if (info && free_b < bytes + headroom) { want = need; bytes = want * sizeof(__nv_bfloat16); }
        if (!info || free_b < bytes + headroom || cudaMalloc(&e->d_shadow, bytes) != cudaSuccess) {
            cudaGetLastError();
            e->shadow_unavailable = true;      // stays off for this engine: TF32 nominations need no extra memory
            return WAX_VS_OK;
        }
        e->shadow_cap = want;
    }
    if (e->shadow_rows > e->n_rows) e->shadow_rows = 0;
    const uint64_t first = e->shadow_rows, count = e->n_rows - first;
    if (count) {
        const int grid = static_cast<int>(std::min<uint64_t>(static_cast<uint64_t>(e->sm_count) * 16, (count * (e->dims / 4) + 255) / 256));
        shadow_bf16_kernel<<<std::max(grid, 1), 256, 0, stream>>>(e->d_corpus + first * e->dims,
                                                                  e->similarity == WAX_VS_COSINE ? e->d_inv_norm + first : nullptr,
                                                                  count, e->dims, e->d_shadow + first * e->dims);
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaStreamSynchronize(stream));
    e->shadow_rows = e->n_rows;
    e->shadow_valid = true;
    return WAX_VS_OK;
}

template <typename K>
static cudaError_t set_smem_attr(K kernel, uint32_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
}

template <typename K>
static cudaError_t launch_nominate(K kernel, uint32_t grid, uint32_t smem, bool pair, cudaStream_t stream,
                                   const CUtensorMap &map_q, const CUtensorMap &map_c, const BatchParams &bp) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kBatchThreads); cfg.stream = stream; cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    if (pair) {
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, kernel, map_q, map_c, bp);
}

// ring depth of the ARES shapes: what is left of the 227 KB after the resident queries
static int ares_stages(bool pair, int heap, uint32_t num_kb, int want) {
    int st = want;
    while (st > 2 && batch_ares_smem_bytes(st, heap, pair, static_cast<int>(num_kb)) > 227u * 1024u) --st;
    return batch_ares_smem_bytes(st, heap, pair, static_cast<int>(num_kb)) <= 227u * 1024u ? st : 0;
}

// P(X >= h) for X ~ Poisson(m): the chance that one row slice holds h or more of a query's "threatening" rows.
static double poisson_tail(double m, int h) {
    double term = std::exp(-m);                      // P(X = 0)
    for (int i = 1; i <= h; ++i) term *= m / i;      // P(X = h)
    double tail = 0.0, t = term;
    for (int i = h + 1; i < h + 200 && t > 1e-300; ++i) { tail += t; t *= m / i; }
    return std::min(1.0, tail + t);
}

// Enqueue the tensor-core nomination + exact finish for n_queries device-resident queries.  d_ok[i] = 1 when
// query i's result is proven exact; the caller sends the others to the filter level, then to enqueue_search.
// allow_bf16 = false forces TF32 nominations (adaptive level choice).  *used_bf16 reports what ran; d_tau_star
// (optional) receives each query's threshold for the filter level.
static int32_t enqueue_batch_tensor(wax_vs_engine *e, SearchCtx *c, const float *d_queries, uint32_t n_queries,
                                    uint32_t k_eff, uint64_t row_offset, wax_vs_candidate *d_out, uint32_t *d_ok,
                                    const uint64_t *d_ids, cudaStream_t stream, uint64_t *launches,
                                    bool allow_bf16 = true, bool *used_bf16 = nullptr, float *d_tau_star = nullptr,
                                    const uint32_t *d_mask = nullptr, uint32_t *used_heap = nullptr) {
    int32_t rc = ensure_norms(e, stream);
    if (rc) return rc;
    bool bf16 = allow_bf16 && batch_bf16_wanted(e);
    if (bf16) {
        if ((rc = ensure_shadow(e, stream))) return rc;
        bf16 = e->shadow_valid;
    }
    if (used_bf16) *used_bf16 = bf16;
    // k > 128: the union of the slices' 64-entry heaps must hold k nominees with some room (slices >= 1.15 k / 64), so
    // fewer query groups share the SMs and a large batch is split into several launches
    uint32_t max_groups = static_cast<uint32_t>(e->sm_count);
    if (k_eff > 128u) max_groups = std::max<uint32_t>(1u, max_groups / ((k_eff * 115u / 100u + 63u) / 64u));
    cudaError_t attr_err = cudaSuccess;
    {   // per function and per DEVICE: once per engine
        std::lock_guard<std::mutex> ag(e->attr_mu);
        if (!e->batch_attr_set) {
            auto chk = [&](cudaError_t r) { if (attr_err == cudaSuccess) attr_err = r; };
            chk(set_smem_attr(batch_nominate_kernel<4, 16, false>, batch_smem_bytes(4, 16)));
            chk(set_smem_attr(batch_nominate_kernel<3, 64, false>, batch_smem_bytes(3, 64)));
            chk(set_smem_attr(batch_nominate_kernel<6, 16, true>, batch_smem_bytes(6, 16, true)));
            chk(set_smem_attr(batch_nominate_kernel<4, 64, true>, batch_smem_bytes(4, 64, true)));
            chk(set_smem_attr(batch_nominate_kernel<4, 16, false, true>, batch_smem_bytes(4, 16)));
            chk(set_smem_attr(batch_nominate_kernel<3, 64, false, true>, batch_smem_bytes(3, 64)));
            chk(set_smem_attr(batch_nominate_kernel<4, 24, false, true>, batch_smem_bytes(4, 24) - 2048u));
            chk(set_smem_attr(batch_nominate_kernel<6, 16, true, true>, batch_smem_bytes(6, 16, true)));
            chk(set_smem_attr(batch_nominate_kernel<4, 64, true, true>, batch_smem_bytes(4, 64, true)));
            // ARES shapes: the ring depth is chosen at run time (<= the template's STAGES is what the kernel uses)
            chk(set_smem_attr(batch_nominate_kernel<3, 16, false, true, true>, 227u * 1024u));
            chk(set_smem_attr(batch_nominate_kernel<2, 16, false, true, true>, 227u * 1024u));
            chk(set_smem_attr(batch_nominate_kernel<2, 64, false, true, true>, 227u * 1024u));
            chk(set_smem_attr(batch_nominate_kernel<6, 16, true, true, true>, 227u * 1024u));
            chk(set_smem_attr(batch_nominate_kernel<4, 16, true, true, true>, 227u * 1024u));
            chk(set_smem_attr(batch_nominate_kernel<4, 64, true, true, true>, 227u * 1024u));
            chk(set_smem_attr(batch_nominate_kernel<5, 32, true, true, true>, 227u * 1024u));
            chk(set_smem_attr(batch_nominate_kernel<3, 24, false, true, true>, 227u * 1024u));
            chk(set_smem_attr(batch_nominate_kernel<5, 32, true, true>, batch_smem_bytes(5, 32, true)));
            chk(set_smem_attr(batch_tf32_ts_kernel<16>, batch_ts_smem_bytes(16)));
            chk(set_smem_attr(batch_tf32_ts_kernel<64>, batch_ts_smem_bytes(64)));
            chk(set_smem_attr(batch_nominate_kernel<4, 16, false, false, false, true>, batch_smem_bytes(4, 16)));
            chk(set_smem_attr(batch_nominate_kernel<4, 16, false, true, false, true>, batch_smem_bytes(4, 16)));
            chk(set_smem_attr(batch_nominate_kernel<3, 16, false, true, true, true>, 227u * 1024u));
            chk(set_smem_attr(filter_select_kernel, 16384 * 8));
            chk(set_smem_attr(batch_finish_kernel<kCosine>, (16384 + kBatchRescoreMax) * 8));
            chk(set_smem_attr(batch_finish_kernel<kDot>, (16384 + kBatchRescoreMax) * 8));
            e->batch_attr_set = attr_err == cudaSuccess;
        }
    }
    if (attr_err != cudaSuccess) return fail(WAX_VS_ERR_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err));

    // bf16 nominations: convert the queries once per call (n_queries x dims, tiny next to the corpus pass)
    if (bf16) {
        const size_t qn = static_cast<size_t>(n_queries) * e->dims;
        if ((rc = ensure_dev(&c->d_queries_bf16, &c->queries_bf16_cap, qn, "bf16 queries"))) return rc;
        const int g = static_cast<int>(std::min<size_t>((qn / 4 + 255) / 256, static_cast<size_t>(e->sm_count) * 8));
        shadow_bf16_kernel<<<std::max(g, 1), 256, 0, stream>>>(d_queries, nullptr, n_queries, e->dims, c->d_queries_bf16);
        CUDA_TRY(cudaGetLastError());
        ++*launches;
    }
    // how many nominees the finish kernel re-scores exactly: the (rescore+1)-th nominee bounds the rows it skips, and
    // the coarser bf16 bound needs more distance between it and the k-th result
    uint32_t rescore = static_cast<uint32_t>(kBatchRescore);
    if (e->tune.batch_rescore > 0) rescore = static_cast<uint32_t>(e->tune.batch_rescore);
    else if (bf16) rescore = k_eff <= 16 ? 256u : (k_eff <= 48 ? 512u : 1024u);
    if (k_eff > 128u) rescore = static_cast<uint32_t>(kBatchRescoreMax);   // the finish kernel writes k re-scored nominees
    rescore = rescore <= 256u ? 256u : (rescore <= 512u ? 512u : static_cast<uint32_t>(kBatchRescoreMax));

    for (uint32_t q0 = 0; q0 < n_queries; q0 += max_groups * kBatchM) {
        const uint32_t nq = std::min<uint32_t>(n_queries - q0, max_groups * kBatchM);
        uint32_t groups = (nq + kBatchM - 1) / kBatchM;
        // cta_group::2: CTA pairs (two query groups, one row slice) issue one 256-row MMA and each stages only half of
        // the corpus tile.  Needs at least two groups; an odd group count is padded with an all-out-of-range group.
        // TS shape: queries in TMEM + CTA pair (dims <= 384, dims % 128 == 0): shared memory carries only the corpus
        const bool ts = !bf16 && !d_mask && e->tune.batch_ts != 0 && groups >= 2 && e->dims <= 384 && e->dims % 128u == 0;
        bool pair = ts || (e->tune.batch_pair != 0 && groups >= 2);
        const uint32_t tile_rows = ts ? static_cast<uint32_t>(kTsN) : static_cast<uint32_t>(kBatchN);
        const uint32_t tiles_total = static_cast<uint32_t>((e->n_rows + tile_rows - 1) / tile_rows);
        auto slices_for = [&](bool pr, uint32_t g) {
            const uint32_t units = pr ? ((g + 1u) & ~1u) / 2u : g;                   // clusters (or CTAs) per slice
            const uint32_t unit_slots = static_cast<uint32_t>(e->sm_count) / (pr ? 2u : 1u);
            return std::max<uint32_t>(1, std::min<uint32_t>(unit_slots / units, tiles_total));
        };
        uint32_t slices = slices_for(pair, groups);
        // Nominee heap size per (slice, query) = kernel shape.  TF32 / TS shapes: 16 entries when 16 nominees per slice
        // comfortably cover k (16 * slices >= 8 k), else 64.  bf16 shapes (16 / 24 / 32 / 64): level 1 can prove a query
        // only if no slice holds `heap` rows scoring within the bf16 bound of the k-th result; with the corpus spread over
        // the slices those "threatening" rows (about 2.2 k of them for the bf16 bound on unit-scale embeddings) fall
        // ~Poisson(m = 2.2 k / slices) per slice.  ONE unproven query costs its whole batch a second pass (DESIGN 4.5.2:
        // k = 72 over 18 slices left 209 of 1024 queries unproven with 16 entries, none with 24 / 32), while 64-entry
        // heaps cost a stage of the ring (8.6 vs 6.1-6.9 ms): the heap that minimises the expected cost is picked below.
        const bool small_heap = e->tune.batch_heap == 16 || (e->tune.batch_heap == 0 && 16u * slices >= 8u * k_eff);
        uint32_t kprime = small_heap ? 16u : 64u;
        if (bf16 && !ts) {
            uint32_t want = static_cast<uint32_t>(std::max(e->tune.batch_heap, 0));
            if (k_eff > 128u && want == 0u) want = 64u;       // large k: level 1 only has to NOMINATE k rows (filter level decides)
            if (want != 16u && want != 24u && want != 32u && want != 64u) {
                // expected cost of a batch = the shape's relative time + P(some query of the batch is unproven) x one more
                // pass.  Threatening rows per query: ~2.2 k (cosine, unit rows) / ~2.8 k (dot: the bound scales with the
                // LARGEST row norm) -- calibrated on profiles/batch_k_sweep_r02p.jsonl and c5_proof_heap16_r02c.jsonl.
                static const uint32_t ladder[4] = {16u, 24u, 32u, 64u};
                static const double rel_time[4] = {1.00, 1.02, 1.10, 1.40};
                const double m = (e->similarity == WAX_VS_DOT ? 2.8 : 2.2) * k_eff / slices;
                uint32_t bump = 0;
                { std::lock_guard<std::mutex> pg(e->pool_mu); bump = e->heap_bump; }
                double best = 1e30;
                uint32_t pick = 3u;
                for (uint32_t i = 0; i < 4u; ++i) {
                    if (ladder[i] == 24u && pair) continue;                     // 24: single-CTA shapes only
                    if (ladder[i] == 32u && groups < 2u) continue;              // 32: cta_group::2 shapes only
                    const double p_fail = std::min(1.0, static_cast<double>(nq) * slices * poisson_tail(m, static_cast<int>(ladder[i])));
                    if (p_fail >= 1.0 && ladder[i] != 64u) continue;             // hopeless: every batch would pay a second pass
                    const double cost = rel_time[i] + 4.0 * p_fail;              // risk-averse: the model can be off
                    if (cost < best) { best = cost; pick = i; }
                }
                for (; bump > 0u && pick < 3u; --bump) {                         // the data overrules the model (see caller)
                    ++pick;
                    if (ladder[pick] == 24u && pair) ++pick;
                    if (pick < 3u && ladder[pick] == 32u && groups < 2u) ++pick;
                }
                want = ladder[std::min(pick, 3u)];
            }
            if (want == 24u && pair) want = 32u;                       // 24: single-CTA shapes; 32: cta_group::2 shapes
            if (want == 32u && groups < 2u) want = 64u;
            if (want == 32u && !pair) { pair = true; slices = slices_for(true, groups); }
            kprime = want;
            if (used_heap) *used_heap = std::max(*used_heap, kprime);
        }
        if (pair) groups = (groups + 1u) & ~1u;
        // resident queries (bf16) when they leave room for a useful ring: >= 3 corpus stages (pair: >= 4 half-tile stages)
        const uint32_t num_kb16 = e->dims / kBatchKBlockBf16;
        const int ares_want = pair ? (kprime == 32u ? 5 : (kprime == 64u ? 4 : 6)) : (kprime == 64u ? 2 : 3);
        const int ares_st = (bf16 && !ts && e->tune.batch_ares) ? ares_stages(pair, static_cast<int>(kprime), num_kb16, ares_want) : 0;
        const bool ares = bf16 && !ts && ares_st >= (pair ? 4 : 2) && !(kprime >= 24u && ares_st < ares_want);
        slices = std::max<uint32_t>(1, std::min<uint32_t>(slices, 16384u / kprime));   // union fits the finish sort
        const uint32_t grid = groups * slices;
        if ((rc = ensure_dev(&c->d_heaps, &c->heaps_cap, static_cast<size_t>(grid) * kBatchM * kprime, "nominee heaps"))) return rc;
        if ((rc = ensure_dev(&c->d_tau, &c->tau_cap, static_cast<size_t>(groups) * kBatchM, "shared thresholds"))) return rc;
        CUDA_TRY(cudaMemsetAsync(c->d_tau, 0, static_cast<size_t>(groups) * kBatchM * sizeof(uint32_t), stream));
        CUtensorMap map_q, map_c;
        const float *qbase = d_queries + static_cast<size_t>(q0) * e->dims;
        const uint32_t c_box = ts ? kTsN / 2 : (pair ? kBatchN / 2 : kBatchN);
        if (bf16) {
            if ((rc = make_tensor_map(&map_q, c->d_queries_bf16 + static_cast<size_t>(q0) * e->dims, nq, e->dims, kBatchM, true))) return rc;
            if ((rc = make_tensor_map(&map_c, e->d_shadow, e->n_rows, e->dims, c_box, true))) return rc;
        } else {
            if ((rc = make_tensor_map(&map_q, qbase, nq, e->dims, kBatchM))) return rc;
            if ((rc = make_tensor_map(&map_c, e->d_corpus, e->n_rows, e->dims, c_box))) return rc;
        }

        BatchParams bp{};
        bp.n_rows = static_cast<uint32_t>(e->n_rows); bp.dims = e->dims; bp.n_queries = nq; bp.groups = groups;
        bp.slices = slices; bp.tiles_total = tiles_total; bp.kprime = kprime; bp.metric = e->similarity;
        // the cosine shadow rows are pre-normalised: no epilogue scaling on the bf16 path
        bp.row_scale = (e->similarity == WAX_VS_COSINE && !bf16) ? e->d_inv_norm : nullptr;
        bp.heaps = c->d_heaps;
        bp.tau_global = c->d_tau;
        bp.no_insert = e->tune.batch_noinsert ? 1u : 0u;
        bp.allow_bits = d_mask;
        cudaError_t lerr = cudaSuccess;
        if (ts) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kBatchThreads); cfg.stream = stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            if (small_heap) { cfg.dynamicSmemBytes = batch_ts_smem_bytes(16); lerr = cudaLaunchKernelEx(&cfg, batch_tf32_ts_kernel<16>, map_c, qbase, bp); }
            else { cfg.dynamicSmemBytes = batch_ts_smem_bytes(64); lerr = cudaLaunchKernelEx(&cfg, batch_tf32_ts_kernel<64>, map_c, qbase, bp); }
        } else if (bf16) {
            const int kb = static_cast<int>(num_kb16);
            const bool h16 = kprime == 16u;
#define WAXVS_NOM(ST, HP, PR, AR, SMEM) lerr = launch_nominate(batch_nominate_kernel<ST, HP, PR, true, AR>, grid, (SMEM), PR, stream, map_q, map_c, bp)
            if (ares && pair) {
                if (kprime == 64u) WAXVS_NOM(4, 64, true, true, batch_ares_smem_bytes(4, 64, true, kb));
                else if (kprime == 32u) WAXVS_NOM(5, 32, true, true, batch_ares_smem_bytes(5, 32, true, kb));
                else if (ares_st >= 6) WAXVS_NOM(6, 16, true, true, batch_ares_smem_bytes(6, 16, true, kb));
                else WAXVS_NOM(4, 16, true, true, batch_ares_smem_bytes(4, 16, true, kb));
            } else if (ares) {
                if (kprime == 64u) WAXVS_NOM(2, 64, false, true, batch_ares_smem_bytes(2, 64, false, kb));
                else if (kprime == 24u) WAXVS_NOM(3, 24, false, true, batch_ares_smem_bytes(3, 24, false, kb));
                else if (ares_st >= 3) WAXVS_NOM(3, 16, false, true, batch_ares_smem_bytes(3, 16, false, kb));
                else WAXVS_NOM(2, 16, false, true, batch_ares_smem_bytes(2, 16, false, kb));
            } else if (pair) {
                if (h16) WAXVS_NOM(6, 16, true, false, batch_smem_bytes(6, 16, true));
                else if (kprime == 32u) WAXVS_NOM(5, 32, true, false, batch_smem_bytes(5, 32, true));
                else WAXVS_NOM(4, 64, true, false, batch_smem_bytes(4, 64, true));
            } else {
                if (kprime == 24u) WAXVS_NOM(4, 24, false, false, batch_smem_bytes(4, 24) - 2048u);
                else if (h16) WAXVS_NOM(4, 16, false, false, batch_smem_bytes(4, 16));
                else WAXVS_NOM(3, 64, false, false, batch_smem_bytes(3, 64));
            }
#undef WAXVS_NOM
        } else if (pair) {
            if (small_heap) lerr = launch_nominate(batch_nominate_kernel<6, 16, true>, grid, batch_smem_bytes(6, 16, true), true, stream, map_q, map_c, bp);
            else lerr = launch_nominate(batch_nominate_kernel<4, 64, true>, grid, batch_smem_bytes(4, 64, true), true, stream, map_q, map_c, bp);
        } else if (small_heap) {
            lerr = launch_nominate(batch_nominate_kernel<4, 16, false>, grid, batch_smem_bytes(4, 16), false, stream, map_q, map_c, bp);
        } else {
            lerr = launch_nominate(batch_nominate_kernel<3, 64, false>, grid, batch_smem_bytes(3, 64), false, stream, map_q, map_c, bp);
        }
        CUDA_TRY(lerr);
        CUDA_TRY(cudaGetLastError());

        FinishParams fp{};
        fp.corpus = e->d_corpus; fp.queries = qbase; fp.n_rows = bp.n_rows; fp.dims = e->dims; fp.n_queries = nq;
        fp.groups = groups; fp.slices = slices; fp.kprime = kprime; fp.k = k_eff; fp.metric = e->similarity;
        fp.heaps = c->d_heaps; fp.max_norm_bits = e->d_max_norm;
        fp.out = d_out + static_cast<size_t>(q0) * k_eff; fp.ok = d_ok + q0;
        fp.frame_ids = d_ids; fp.id_base = e->id_base; fp.row_offset = row_offset;
        uint32_t pow2 = 512;
        while (pow2 < slices * kprime) pow2 <<= 1;
        fp.pow2_all = pow2;
        fp.rescore = rescore;
        fp.eps_rel = bf16 ? kBf16Eps : kTf32Eps;
        fp.tau_star = d_tau_star ? d_tau_star + q0 : nullptr;
        fp.tau_stride = n_queries;
        const size_t fsmem = static_cast<size_t>(pow2 + rescore) * sizeof(uint64_t);
        if (e->similarity == WAX_VS_COSINE) batch_finish_kernel<kCosine><<<nq, 512, fsmem, stream>>>(fp);
        else batch_finish_kernel<kDot><<<nq, 512, fsmem, stream>>>(fp);
        CUDA_TRY(cudaGetLastError());
        *launches += 2;
    }
    return WAX_VS_OK;
}

static uint32_t clamp_topk(int64_t k) {  // MetalVectorEngine.swift:842-846
    if (k < 1) return 1;
    if (k > WAX_VS_MAX_RESULTS) return WAX_VS_MAX_RESULTS;
    return static_cast<uint32_t>(k);
}

// VectorMetric.score(fromDistance:) (VectorMetric.swift:32-43)
static float score_from_distance(uint8_t sim, float d) {
    if (!finite_f32(d)) return 0.0f;
    return sim == WAX_VS_COSINE ? 1.0f - d : -d;
}

// ---------------------------------------------------------------------------------------------------------
// corpus storage
static int32_t set_capacity(wax_vs_engine *e, uint64_t rows) {
    if (rows <= e->cap_rows) return WAX_VS_OK;
    float *n = nullptr;
    const size_t bytes = static_cast<size_t>(rows) * e->dims * sizeof(float);
    if (cudaMalloc(&n, bytes) != cudaSuccess) {
        // the bf16 shadow is derived data: give its HBM back before giving up (mutators hold the write lock)
        cudaGetLastError();
        if (e->d_shadow) {
            cudaFree(e->d_shadow); e->d_shadow = nullptr; e->shadow_cap = 0; e->shadow_valid = false; e->shadow_rows = 0;
            e->shadow_unavailable = true;
        }
        if (cudaMalloc(&n, bytes) != cudaSuccess)
            return fail(WAX_VS_ERR_CUDA, "Failed to resize vectors buffer (%zu bytes): %s", bytes,
                        cudaGetErrorString(cudaGetLastError()));
    }
    if (e->n_rows) {
        cudaError_t err = cudaMemcpy(n, e->d_corpus, static_cast<size_t>(e->n_rows) * e->dims * sizeof(float),
                                     cudaMemcpyDeviceToDevice);
        if (err != cudaSuccess) { cudaFree(n); return fail(WAX_VS_ERR_CUDA, "corpus copy failed: %s", cudaGetErrorString(err)); }
    }
    if (e->d_corpus) cudaFree(e->d_corpus);
    e->d_corpus = n;
    e->cap_rows = rows;
    return WAX_VS_OK;
}

// reserveIfNeeded (MetalVectorEngine.swift:857-871): doubling from 64.
static int32_t grow_for(wax_vs_engine *e, uint64_t required) {
    if (required > 0xFFFFFFFFull)
        return fail(WAX_VS_ERR_CAPACITY, "capacity exceeded: limit %llu, requested %llu", 0xFFFFFFFFull,
                    static_cast<unsigned long long>(required));
    if (required <= e->cap_rows) return WAX_VS_OK;
    uint64_t next = e->cap_rows ? e->cap_rows : 64;
    while (next < required) next = std::min<uint64_t>(next * 2, 0xFFFFFFFFull);
    return set_capacity(e, next);
}

static void materialize_ids(wax_vs_engine *e) {
    if (!e->ids_identity) return;
    e->ids.resize(e->n_rows);
    for (uint64_t r = 0; r < e->n_rows; ++r) e->ids[r] = e->id_base + r;
    e->ids_identity = false;
    e->map_valid = false;
    e->ids_sorted = true;            // id_base + row
    e->d_ids_dirty = true;
}
static void ensure_map(wax_vs_engine *e) {
    if (e->map_valid) return;
    e->map.reset(e->ids.size());
    for (size_t r = 0; r < e->ids.size(); ++r) e->map.put(e->ids[r], static_cast<uint32_t>(r));
    e->map_valid = true;
}
// frameId -> row (0xFFFFFFFF = absent) for explicit ids: binary search while the id array is sorted, else the hash table.
static uint32_t find_row(wax_vs_engine *e, uint64_t id) {
    if (e->ids_sorted) {
        const auto it = std::lower_bound(e->ids.begin(), e->ids.end(), id);
        return (it != e->ids.end() && *it == id) ? static_cast<uint32_t>(it - e->ids.begin()) : 0xFFFFFFFFu;
    }
    ensure_map(e);
    return e->map.find(id);
}
static int32_t sync_device_ids(wax_vs_engine *e, const uint64_t **out) {
    std::lock_guard<std::mutex> g(e->ids_mu);
    if (e->ids_identity) { *out = nullptr; return WAX_VS_OK; }
    if (e->d_ids_dirty) {
        int32_t rc = ensure_dev(&e->d_ids, &e->d_ids_cap, std::max<size_t>(e->ids.size(), 1), "frame ids");
        if (rc) return rc;
        if (!e->ids.empty())
            CUDA_TRY(cudaMemcpy(e->d_ids, e->ids.data(), e->ids.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
        e->d_ids_dirty = false;
    }
    *out = e->d_ids;
    return WAX_VS_OK;
}

// ---------------------------------------------------------------------------------------------------------
// C-ABI
extern "C" {

const char *wax_vs_last_error(void) { return g_last_error; }
const char *wax_vs_version(void) { return "waxvs_cuda 0.1 sm_100a (fused scan+top-k; TMA bulk staging)"; }

int32_t wax_vs_device_count(int32_t *out) {
    if (!out) return fail(WAX_VS_ERR_NULL, "out is NULL");
    int n = 0;
    cudaError_t err = cudaGetDeviceCount(&n);
    if (err != cudaSuccess) { *out = 0; cudaGetLastError(); return fail(WAX_VS_ERR_CUDA, "CUDA device not available: %s", cudaGetErrorString(err)); }
    *out = n;
    return WAX_VS_OK;
}

int32_t wax_vs_create(uint32_t dimensions, uint8_t similarity, const int32_t *devices, int32_t n_devices,
                      wax_vs_engine **out) {
    if (!out) return fail(WAX_VS_ERR_NULL, "out is NULL");
    *out = nullptr;
    if (dimensions == 0) return fail(WAX_VS_ERR_ARGUMENT, "dimensions must be > 0");  // :154-156
    if (dimensions > WAX_VS_MAX_DIMENSIONS)                                           // :157-162
        return fail(WAX_VS_ERR_CAPACITY, "capacity exceeded: limit %d, requested %u", WAX_VS_MAX_DIMENSIONS, dimensions);
    if (similarity > 2) return fail(WAX_VS_ERR_ARGUMENT, "vec similarity must be 0..2 (got %u)", similarity);
    if (n_devices > 1)
        return fail(WAX_VS_ERR_UNSUPPORTED, "one engine drives one device; shard with one engine per rank (wax_vs_search_device)");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return fail(WAX_VS_ERR_CUDA, "CUDA device not available");  // "Metal device not available" :167-169
    }
    int dev = 0;
    if (devices && n_devices == 1) dev = devices[0];
    else if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    if (dev < 0 || dev >= count) return fail(WAX_VS_ERR_ARGUMENT, "device ordinal %d out of range (0..%d)", dev, count - 1);

    wax_vs_engine *e = new (std::nothrow) wax_vs_engine();
    if (!e) return fail(WAX_VS_ERR_CUDA, "out of host memory");
    e->device = dev; e->dims = dimensions; e->similarity = similarity;
    DeviceGuard g(dev);
    int v = 0;
    if (!g.ok || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
        delete e;
        return fail(WAX_VS_ERR_CUDA, "failed to select CUDA device %d", dev);
    }
    e->sm_count = v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess) e->smem_optin = static_cast<size_t>(v);
    int32_t rc = set_capacity(e, 64);  // initialReserve (:19, :225-229)
    if (rc) { delete e; return rc; }
    *out = e;
    return WAX_VS_OK;
}

void wax_vs_destroy(wax_vs_engine *e) {
    if (!e) return;
    {
        std::unique_lock<std::shared_mutex> w(e->rw);
        DeviceGuard g(e->device);
        cudaDeviceSynchronize();
        shard_teardown(e, true);
        ingest_free(e);
        for (SearchCtx *c : e->pool) ctx_free(c);
        for (auto &kv : e->stream_ctx) ctx_free(kv.second);
        if (e->d_corpus) cudaFree(e->d_corpus);
        if (e->d_ids) cudaFree(e->d_ids);
        if (e->d_inv_norm) cudaFree(e->d_inv_norm);
        if (e->d_max_norm) cudaFree(e->d_max_norm);
        if (e->d_shadow) cudaFree(e->d_shadow);
    }
    delete e;
}

int32_t wax_vs_dimensions(const wax_vs_engine *e, uint32_t *out) {
    if (!e || !out) return fail(WAX_VS_ERR_NULL, "NULL argument");
    *out = e->dims;
    return WAX_VS_OK;
}
int32_t wax_vs_similarity(const wax_vs_engine *e, uint8_t *out) {
    if (!e || !out) return fail(WAX_VS_ERR_NULL, "NULL argument");
    *out = e->similarity;
    return WAX_VS_OK;
}
int32_t wax_vs_count(wax_vs_engine *e, uint64_t *out) {
    if (!e || !out) return fail(WAX_VS_ERR_NULL, "NULL argument");
    std::shared_lock<std::shared_mutex> r(e->rw);
    *out = e->n_rows;
    return WAX_VS_OK;
}

int32_t wax_vs_reserve(wax_vs_engine *e, uint64_t rows) {
    if (!e) return fail(WAX_VS_ERR_NULL, "engine is NULL");
    if (rows > 0xFFFFFFFFull)
        return fail(WAX_VS_ERR_CAPACITY, "capacity exceeded: limit %llu, requested %llu", 0xFFFFFFFFull,
                    static_cast<unsigned long long>(rows));
    std::unique_lock<std::shared_mutex> w(e->rw);
    DeviceGuard g(e->device);
    drain_device_path(e);
    int32_t rc = set_capacity(e, rows);
    if (rc == WAX_VS_OK && !e->ids_identity && rows > e->ids.capacity()) {
        e->ids.reserve(rows);                       // no id-array / hash-table regrowth during the appends that follow
        if (!e->ids_sorted && e->map_valid && e->map.keys.size() < rows * 2 + 16) {
            IdMap bigger;
            bigger.reset(rows);
            for (size_t r = 0; r < e->ids.size(); ++r) bigger.put(e->ids[r], static_cast<uint32_t>(r));
            e->map = std::move(bigger);
        }
    }
    return rc;
}

// Phase trace of the mutators for performance work: WAXVS_TRACE_INGEST=1 prints "<what>: <phase> <us>" lines to stderr.
struct IngestTrace {
    bool on;
    const char *what;
    std::chrono::steady_clock::time_point t;
    explicit IngestTrace(const char *w) : on(getenv("WAXVS_TRACE_INGEST") != nullptr), what(w), t(std::chrono::steady_clock::now()) {}
    void mark(const char *phase) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[waxvs] %s: %s %.1f us\n", what, phase, std::chrono::duration<double, std::micro>(now - t).count());
        t = now;
    }
};

// ---- bulk ingest / export plumbing ---------------------------------------------------------------
// Host <-> HBM at device speed from PAGEABLE caller memory: the bytes go through two pinned staging buffers; worker
// threads fill (or drain) one buffer while the DMA engine moves the other.  Caller memory that is already pinned
// (cudaHostAlloc / cudaHostRegister) is handed to the DMA engine directly.
static int32_t ingest_init(wax_vs_engine *e) {
    auto &g = e->ing;
    if (g.stream) return WAX_VS_OK;
    CUDA_TRY(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) CUDA_TRY(cudaEventCreateWithFlags(&g.ev[i], cudaEventDisableTiming));
    unsigned hw = std::thread::hardware_concurrency();
    long quota = 0, period = 0;                       // cgroup v2 CPU quota, when there is one
    if (FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r")) {
        char q[32] = "";
        if (fscanf(f, "%31s %ld", q, &period) == 2 && strcmp(q, "max") != 0) quota = atol(q);
        fclose(f);
    }
    if (quota > 0 && period > 0) hw = static_cast<unsigned>(std::max<long>(1, std::min<long>(hw ? hw : 1, (quota + period / 2) / period)));
    g.threads = static_cast<int>(std::max(1u, std::min(8u, hw ? hw : 1u)));
    return WAX_VS_OK;
}
// The two large pinned staging buffers are shared by every engine of the process (pinning 2 x 64 MB costs ~0.1 s, far
// more than most transfers): whoever holds g_staging_mu owns them for the length of one upload / download.  Small
// transfers use the engine's own 1 MB pair.
static std::mutex g_staging_mu;
static uint8_t *g_staging[2] = {nullptr, nullptr};
constexpr size_t kStagingBytes = size_t(64) << 20;
constexpr size_t kSmallStagingBytes = size_t(1) << 20;

static int32_t ingest_staging(wax_vs_engine *e, size_t want_bytes) {
    auto &g = e->ing;
    int32_t rc = ingest_init(e);
    if (rc) return rc;
    (void)want_bytes;
    if (g.pin_bytes >= kSmallStagingBytes) return WAX_VS_OK;
    for (int i = 0; i < 2; ++i) {
        if (cudaHostAlloc(reinterpret_cast<void **>(&g.pin[i]), kSmallStagingBytes, cudaHostAllocPortable) != cudaSuccess)
            return fail(WAX_VS_ERR_CUDA, "failed to allocate pinned ingest staging: %s", cudaGetErrorString(cudaGetLastError()));
    }
    g.pin_bytes = kSmallStagingBytes;
    return WAX_VS_OK;
}
// caller holds g_staging_mu
static int32_t shared_staging() {
    for (int i = 0; i < 2; ++i) {
        if (!g_staging[i] && cudaHostAlloc(reinterpret_cast<void **>(&g_staging[i]), kStagingBytes, cudaHostAllocPortable) != cudaSuccess)
            return fail(WAX_VS_ERR_CUDA, "failed to allocate the shared pinned staging (%zu bytes): %s", kStagingBytes, cudaGetErrorString(cudaGetLastError()));
    }
    return WAX_VS_OK;
}
static void ingest_free(wax_vs_engine *e) {
    auto &g = e->ing;
    for (int i = 0; i < 2; ++i) {
        if (g.pin[i]) cudaFreeHost(g.pin[i]);
        if (g.ev[i]) cudaEventDestroy(g.ev[i]);
        g.pin[i] = nullptr; g.ev[i] = nullptr;
    }
    if (g.d_stage) cudaFree(g.d_stage);
    if (g.d_index) cudaFree(g.d_index);
    if (g.stream) cudaStreamDestroy(g.stream);
    g = wax_vs_engine::Ingest();
}

static void parallel_memcpy(void *dst, const void *src, size_t bytes, int threads) {
    const size_t min_slice = size_t(2) << 20;
    int t = static_cast<int>(std::min<size_t>(static_cast<size_t>(std::max(threads, 1)), std::max<size_t>(1, bytes / min_slice)));
    if (t <= 1) { memcpy(dst, src, bytes); return; }
    std::vector<std::thread> pool;
    pool.reserve(static_cast<size_t>(t - 1));
    const size_t per = ((bytes + t - 1) / t + 63) & ~size_t(63);
    for (int i = 1; i < t; ++i) {
        const size_t off = std::min(bytes, per * i), len = std::min(bytes - off, per);
        if (len) pool.emplace_back([=] { memcpy(static_cast<char *>(dst) + off, static_cast<const char *>(src) + off, len); });
    }
    memcpy(dst, src, std::min(bytes, per));
    for (auto &th : pool) th.join();
}

static bool host_pointer_is_pinned(const void *p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// host (pageable or pinned) -> device, synchronous on return
static int32_t upload_bytes(wax_vs_engine *e, void *d_dst, const void *h_src, size_t bytes) {
    if (bytes == 0) return WAX_VS_OK;
    int32_t rc = ingest_staging(e, bytes);
    if (rc) return rc;
    auto &g = e->ing;
    e->ingest_h2d_bytes += bytes;
    if (host_pointer_is_pinned(h_src)) {
        CUDA_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, g.stream));
        CUDA_TRY(cudaStreamSynchronize(g.stream));
        return WAX_VS_OK;
    }
    const bool big = bytes > 2 * kSmallStagingBytes;
    std::unique_lock<std::mutex> pool(g_staging_mu, std::defer_lock);
    if (big) { pool.lock(); if ((rc = shared_staging())) return rc; }
    uint8_t *const *pin = big ? g_staging : g.pin;
    const size_t pin_bytes = big ? kStagingBytes : g.pin_bytes;
    size_t off = 0;
    for (int i = 0; off < bytes; ++i) {
        const int b = i & 1;
        const size_t len = std::min(pin_bytes, bytes - off);
        CUDA_TRY(cudaEventSynchronize(g.ev[b]));                 // the DMA that last read this buffer is done
        parallel_memcpy(pin[b], static_cast<const char *>(h_src) + off, len, g.threads);
        CUDA_TRY(cudaMemcpyAsync(static_cast<char *>(d_dst) + off, pin[b], len, cudaMemcpyHostToDevice, g.stream));
        CUDA_TRY(cudaEventRecord(g.ev[b], g.stream));
        off += len;
    }
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    return WAX_VS_OK;
}

// device -> host (pageable or pinned), synchronous on return
static int32_t download_bytes(wax_vs_engine *e, void *h_dst, const void *d_src, size_t bytes) {
    if (bytes == 0) return WAX_VS_OK;
    int32_t rc = ingest_staging(e, bytes);
    if (rc) return rc;
    auto &g = e->ing;
    e->ingest_d2h_bytes += bytes;
    if (host_pointer_is_pinned(h_dst)) {
        CUDA_TRY(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, g.stream));
        CUDA_TRY(cudaStreamSynchronize(g.stream));
        return WAX_VS_OK;
    }
    const bool big = bytes > 2 * kSmallStagingBytes;
    std::unique_lock<std::mutex> pool(g_staging_mu, std::defer_lock);
    if (big) { pool.lock(); if ((rc = shared_staging())) return rc; }
    uint8_t *const *pin = big ? g_staging : g.pin;
    const size_t pin_bytes = big ? kStagingBytes : g.pin_bytes;
    const size_t n_chunks = (bytes + pin_bytes - 1) / pin_bytes;
    auto issue = [&](size_t i) -> cudaError_t {
        const size_t off = i * pin_bytes, len = std::min(pin_bytes, bytes - off);
        cudaError_t err = cudaMemcpyAsync(pin[i & 1], static_cast<const char *>(d_src) + off, len, cudaMemcpyDeviceToHost, g.stream);
        if (err == cudaSuccess) err = cudaEventRecord(g.ev[i & 1], g.stream);
        return err;
    };
    CUDA_TRY(issue(0));
    for (size_t i = 0; i < n_chunks; ++i) {
        if (i + 1 < n_chunks) CUDA_TRY(issue(i + 1));            // its buffer was drained in iteration i-1
        CUDA_TRY(cudaEventSynchronize(g.ev[i & 1]));
        const size_t off = i * pin_bytes, len = std::min(pin_bytes, bytes - off);
        parallel_memcpy(static_cast<char *>(h_dst) + off, pin[i & 1], len, g.threads);
    }
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    return WAX_VS_OK;
}

int32_t wax_vs_add_batch(wax_vs_engine *e, const uint64_t *frame_ids, const float *rows, uint64_t n,
                         uint32_t vector_len) {
    if (!e) return fail(WAX_VS_ERR_NULL, "engine is NULL");
    if (n == 0) return WAX_VS_OK;  // guard !frameIds.isEmpty (:360)
    if (!frame_ids || !rows) return fail(WAX_VS_ERR_NULL, "NULL argument");
    if (vector_len != e->dims)     // :367-370
        return fail(WAX_VS_ERR_DIMENSION, "vector dimension mismatch: expected %u, got %u", e->dims, vector_len);
    IngestTrace tr("add_batch");
    std::unique_lock<std::shared_mutex> w(e->rw);
    DeviceGuard g(e->device);
    drain_device_path(e);
    tr.mark("lock+drain");
    int32_t rc = grow_for(e, e->n_rows + n);  // maxNewCount (:379-380)
    if (rc) return rc;
    tr.mark("grow");
    materialize_ids(e);
    tr.mark("ids");

    // Resolve the destination row of every batch item in order (the sequential loop at :384-398).
    std::vector<uint32_t> target(n);
    const uint64_t n0 = e->n_rows;
    bool pure_append = true;
    bool increasing = e->ids_sorted && (e->ids.empty() || frame_ids[0] > e->ids.back());
    for (uint64_t i = 1; i < n && increasing; ++i) increasing = frame_ids[i] > frame_ids[i - 1];
    if (increasing) {
        // the common bulk-ingest case: every id is new and larger than all stored ones -- no lookups, no hash table
        e->ids.insert(e->ids.end(), frame_ids, frame_ids + n);
        for (uint64_t i = 0; i < n; ++i) target[i] = static_cast<uint32_t>(n0 + i);
        e->n_rows += n;
        e->map_valid = false;
    } else {
        for (uint64_t i = 0; i < n; ++i) {
            if (!e->ids_sorted && i + 8 < n) e->map.prefetch(frame_ids[i + 8]);   // the table is far bigger than the caches
            uint32_t row = find_row(e, frame_ids[i]);
            if (row == 0xFFFFFFFFu) {
                row = static_cast<uint32_t>(e->n_rows);
                if (e->ids_sorted && !e->ids.empty() && frame_ids[i] < e->ids.back()) {
                    e->ids_sorted = false;                   // first out-of-order id: from now on the hash table answers
                    e->map_valid = false;
                    ensure_map(e);
                }
                e->ids.push_back(frame_ids[i]);
                if (!e->ids_sorted) e->map.put(frame_ids[i], row);
                else e->map_valid = false;
                ++e->n_rows;
            }
            target[i] = row;
            if (row != n0 + i) pure_append = false;
        }
    }
    e->d_ids_dirty = true;
    const size_t row_bytes = static_cast<size_t>(e->dims) * sizeof(float);
    tr.mark("resolve targets");
    if (pure_append) {
        invalidate_row_caches(e, n0);          // the cached norms / shadow of rows [0, n0) stay valid
        rc = upload_bytes(e, e->d_corpus + n0 * e->dims, rows, n * row_bytes);
        tr.mark("upload");
        return rc;
    }
    invalidate_row_caches(e, 0);
    // Overwrites present: a later item for the same row wins; earlier ones are dropped.
    {
        std::unordered_map<uint32_t, uint64_t> last;
        last.reserve(n * 2);
        for (uint64_t i = 0; i < n; ++i) last[target[i]] = i;
        for (uint64_t i = 0; i < n; ++i) if (last[target[i]] != i) target[i] = 0xFFFFFFFFu;
    }
    // Staged in HBM (persistent staging area, grown on demand) and scattered by one kernel: no per-call allocation.
    auto &ig = e->ing;
    if ((rc = ingest_init(e))) return rc;
    const uint64_t slab_rows = std::max<uint64_t>(1, std::min<uint64_t>(n, (size_t(256) << 20) / row_bytes));
    if ((rc = ensure_dev(&ig.d_stage, &ig.d_stage_cap, static_cast<size_t>(slab_rows) * e->dims, "upsert staging"))) return rc;
    if ((rc = ensure_dev(&ig.d_index, &ig.d_index_cap, static_cast<size_t>(slab_rows), "upsert targets"))) return rc;
    for (uint64_t done = 0; done < n; done += slab_rows) {
        const uint64_t m = std::min(slab_rows, n - done);
        if ((rc = upload_bytes(e, ig.d_stage, rows + done * e->dims, m * row_bytes))) return rc;
        CUDA_TRY(cudaMemcpyAsync(ig.d_index, target.data() + done, m * sizeof(uint32_t), cudaMemcpyHostToDevice, ig.stream));
        scatter_rows_kernel<<<static_cast<unsigned>(m), 128, 0, ig.stream>>>(e->d_corpus, ig.d_stage, ig.d_index, m, e->dims);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(ig.stream));
    }
    return WAX_VS_OK;
}

int32_t wax_vs_add(wax_vs_engine *e, uint64_t frame_id, const float *vector, uint32_t vector_len) {
    return wax_vs_add_batch(e, &frame_id, vector, 1, vector_len);
}

// remove(frameId:) for a whole set of frames (MetalVectorEngine.swift:423-444 applied n times, as ONE pass): unknown ids
// are ignored (:426), the surviving rows keep their relative order (:431-441).  The reference memmoves the tail once
// per id; here the survivors' source rows are computed once on the host, the matrix is compacted slab by slab through
// an HBM bounce buffer (gather kernel + copy back: destinations never overtake unread sources because rows only move
// down and slabs go in ascending order), the id array is compacted once and the id->row hash rebuilt once.
int32_t wax_vs_remove_batch(wax_vs_engine *e, const uint64_t *frame_ids, uint64_t n, uint64_t *out_removed) {
    if (!e) return fail(WAX_VS_ERR_NULL, "engine is NULL");
    if (out_removed) *out_removed = 0;
    if (n == 0) return WAX_VS_OK;
    if (!frame_ids) return fail(WAX_VS_ERR_NULL, "frame_ids is NULL");
    IngestTrace tr("remove_batch");
    std::unique_lock<std::shared_mutex> w(e->rw);
    if (e->n_rows == 0) return WAX_VS_OK;  // :425
    // which rows go
    std::vector<uint32_t> gone;
    gone.reserve(n);
    if (e->ids_identity) {
        for (uint64_t i = 0; i < n; ++i)
            if (frame_ids[i] >= e->id_base && frame_ids[i] - e->id_base < e->n_rows) gone.push_back(static_cast<uint32_t>(frame_ids[i] - e->id_base));
    } else {
        for (uint64_t i = 0; i < n; ++i) {
            const uint32_t r = find_row(e, frame_ids[i]);
            if (r != 0xFFFFFFFFu) gone.push_back(r);                // :426 unknown id = no-op
        }
    }
    if (gone.empty()) return WAX_VS_OK;
    std::sort(gone.begin(), gone.end());
    gone.erase(std::unique(gone.begin(), gone.end()), gone.end());
    DeviceGuard g(e->device);
    drain_device_path(e);
    materialize_ids(e);
    const uint64_t old_n = e->n_rows, first = gone.front(), new_n = old_n - gone.size();
    // source row of every destination row >= first (rows below the first removed row do not move)
    std::vector<uint32_t> src;
    src.reserve(static_cast<size_t>(new_n - first));
    {
        size_t gi = 0;
        for (uint64_t r = first; r < old_n; ++r) {
            if (gi < gone.size() && gone[gi] == r) { ++gi; continue; }
            src.push_back(static_cast<uint32_t>(r));
        }
    }
    tr.mark("resolve rows + sources");
    int32_t rc = ingest_init(e);
    if (rc) return rc;
    auto &ig = e->ing;
    const size_t row_bytes = static_cast<size_t>(e->dims) * sizeof(float);
    const uint64_t moving = src.size();
    if (moving) {
        const uint64_t slab_rows = std::max<uint64_t>(1, std::min<uint64_t>(moving, (size_t(256) << 20) / row_bytes));
        if ((rc = ensure_dev(&ig.d_stage, &ig.d_stage_cap, static_cast<size_t>(slab_rows) * e->dims, "compaction bounce buffer"))) return rc;
        if ((rc = ensure_dev(&ig.d_index, &ig.d_index_cap, static_cast<size_t>(slab_rows), "compaction sources"))) return rc;
        if ((rc = ingest_staging(e, slab_rows * sizeof(uint32_t)))) return rc;
        for (uint64_t done = 0; done < moving; done += slab_rows) {
            const uint64_t m = std::min(slab_rows, moving - done);
            // contiguous runs (nothing removed inside the slab) are a plain device copy; otherwise gather by index
            const bool contiguous = src[done + m - 1] - src[done] == m - 1;
            float *dst = e->d_corpus + (first + done) * e->dims;
            if (contiguous) {
                CUDA_TRY(cudaMemcpyAsync(ig.d_stage, e->d_corpus + static_cast<uint64_t>(src[done]) * e->dims, m * row_bytes,
                                         cudaMemcpyDeviceToDevice, ig.stream));
            } else {
                CUDA_TRY(cudaMemcpyAsync(ig.d_index, src.data() + done, m * sizeof(uint32_t), cudaMemcpyHostToDevice, ig.stream));
                const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(m, static_cast<uint64_t>(e->sm_count) * 32));
                gather_rows_kernel<<<grid, 128, 0, ig.stream>>>(ig.d_stage, e->d_corpus, ig.d_index, m, e->dims);
                CUDA_TRY(cudaGetLastError());
            }
            CUDA_TRY(cudaMemcpyAsync(dst, ig.d_stage, m * row_bytes, cudaMemcpyDeviceToDevice, ig.stream));
            CUDA_TRY(cudaStreamSynchronize(ig.stream));          // src / d_index are reused by the next slab
        }
    }
    tr.mark("compact matrix");
    // ids: one compaction, one hash rebuild (lazily, on the next lookup)
    for (uint64_t j = 0; j < moving; ++j) e->ids[first + j] = e->ids[src[j]];
    e->ids.resize(new_n);
    e->n_rows = new_n;
    e->map_valid = false;
    e->d_ids_dirty = true;
    invalidate_row_caches(e, first);           // rows below the first removed row did not move
    tr.mark("compact ids");
    if (out_removed) *out_removed = gone.size();
    return WAX_VS_OK;
}

int32_t wax_vs_remove(wax_vs_engine *e, uint64_t frame_id) {
    return wax_vs_remove_batch(e, &frame_id, 1, nullptr);
}

// Filter level (level 2 of the batched path): for queries level 1 could not prove.  One TF32 tensor-core pass in
// FILTER form appends EVERY row whose score' beats the query's fixed threshold tau* (= exact k-th score of level 1's
// re-scored nominees - eps_tf32, so no true top-k row can be missing) to the query's candidate list; every candidate
// is re-scored exactly, the k best are the answer.  Complete by construction -- d_ok[i] = 0 only if a list
// overflowed (more than filter_cap rows within 2 eps of the k-th score: near-duplicates en masse).
// bf16 = true: the pass reads the bf16 shadow (half the bytes of the fp32 corpus; d_tau must then be the thresholds built
// with the bf16 bound, which admit more candidates); false: TF32 from the fp32 corpus.
static int32_t enqueue_filter_level(wax_vs_engine *e, SearchCtx *c, const float *d_queries, const float *d_tau,
                                    uint32_t n_queries, uint32_t k_eff, uint64_t row_offset, wax_vs_candidate *d_out,
                                    uint32_t *d_ok, const uint64_t *d_ids, cudaStream_t stream, uint64_t *launches,
                                    bool bf16 = false, const uint32_t *d_mask = nullptr) {
    int32_t rc = ensure_norms(e, stream);
    if (rc) return rc;
    if (bf16) {
        const size_t qn = static_cast<size_t>(n_queries) * e->dims;
        if ((rc = ensure_dev(&c->d_queries_bf16, &c->queries_bf16_cap, qn, "bf16 queries"))) return rc;
        const int g = static_cast<int>(std::min<size_t>((qn / 4 + 255) / 256, static_cast<size_t>(e->sm_count) * 8));
        shadow_bf16_kernel<<<std::max(g, 1), 256, 0, stream>>>(d_queries, nullptr, n_queries, e->dims, c->d_queries_bf16);
        CUDA_TRY(cudaGetLastError());
        ++*launches;
    }
    uint32_t cap = 64;                      // a power of two in [64, 16384], at least k
    while ((cap < static_cast<uint32_t>(std::max(e->tune.filter_cap, 64)) || cap < k_eff) && cap < 16384u) cap <<= 1;
    if ((rc = ensure_dev(&c->d_cand_count, &c->cand_count_cap, static_cast<size_t>(n_queries), "filter counts"))) return rc;
    if ((rc = ensure_dev(&c->d_cand_rows, &c->cand_rows_cap, static_cast<size_t>(n_queries) * cap, "filter candidates"))) return rc;
    if ((rc = ensure_dev(&c->d_cand_keys, &c->cand_keys_cap, static_cast<size_t>(n_queries) * cap, "filter keys"))) return rc;
    CUDA_TRY(cudaMemsetAsync(c->d_cand_count, 0, static_cast<size_t>(n_queries) * sizeof(uint32_t), stream));
    const uint32_t max_groups = static_cast<uint32_t>(e->sm_count);
    const uint32_t tiles_total = static_cast<uint32_t>((e->n_rows + kBatchN - 1) / kBatchN);
    for (uint32_t q0 = 0; q0 < n_queries; q0 += max_groups * kBatchM) {
        const uint32_t nq = std::min<uint32_t>(n_queries - q0, max_groups * kBatchM);
        const uint32_t groups = (nq + kBatchM - 1) / kBatchM;
        const uint32_t slices = std::max<uint32_t>(1, std::min<uint32_t>(static_cast<uint32_t>(e->sm_count) / groups, tiles_total));
        CUtensorMap map_q, map_c;
        const float *qbase = d_queries + static_cast<size_t>(q0) * e->dims;
        if (bf16) {
            if ((rc = make_tensor_map(&map_q, c->d_queries_bf16 + static_cast<size_t>(q0) * e->dims, nq, e->dims, kBatchM, true))) return rc;
            if ((rc = make_tensor_map(&map_c, e->d_shadow, e->n_rows, e->dims, kBatchN, true))) return rc;
        } else {
            if ((rc = make_tensor_map(&map_q, qbase, nq, e->dims, kBatchM))) return rc;
            if ((rc = make_tensor_map(&map_c, e->d_corpus, e->n_rows, e->dims, kBatchN))) return rc;
        }
        BatchParams bp{};
        bp.n_rows = static_cast<uint32_t>(e->n_rows); bp.dims = e->dims; bp.n_queries = nq; bp.groups = groups;
        bp.slices = slices; bp.tiles_total = tiles_total; bp.kprime = 16; bp.metric = e->similarity;
        bp.row_scale = (e->similarity == WAX_VS_COSINE && !bf16) ? e->d_inv_norm : nullptr;   // shadow rows are pre-normalised
        bp.tau_fixed = d_tau + q0;
        bp.cand_count = c->d_cand_count + q0;
        bp.cand_rows = c->d_cand_rows + static_cast<size_t>(q0) * cap;
        bp.cand_cap = cap;
        bp.allow_bits = d_mask;
        if (bf16) {
            const uint32_t num_kb = e->dims / kBatchKBlockBf16;
            const int st = e->tune.batch_ares ? ares_stages(false, 16, num_kb, 3) : 0;
            if (st >= 3)
                CUDA_TRY(launch_nominate(batch_nominate_kernel<3, 16, false, true, true, true>, groups * slices,
                                         batch_ares_smem_bytes(3, 16, false, static_cast<int>(num_kb)), false, stream, map_q, map_c, bp));
            else
                CUDA_TRY(launch_nominate(batch_nominate_kernel<4, 16, false, true, false, true>, groups * slices, batch_smem_bytes(4, 16),
                                         false, stream, map_q, map_c, bp));
        } else {
            CUDA_TRY(launch_nominate(batch_nominate_kernel<4, 16, false, false, false, true>, groups * slices, batch_smem_bytes(4, 16),
                                     false, stream, map_q, map_c, bp));
        }
        const dim3 rgrid(32, nq);
        if (e->similarity == WAX_VS_COSINE)
            filter_rescore_kernel<kCosine><<<rgrid, 256, 0, stream>>>(e->d_corpus, qbase, e->dims, bp.cand_count, bp.cand_rows, cap,
                                                                      c->d_cand_keys + static_cast<size_t>(q0) * cap);
        else
            filter_rescore_kernel<kDot><<<rgrid, 256, 0, stream>>>(e->d_corpus, qbase, e->dims, bp.cand_count, bp.cand_rows, cap,
                                                                   c->d_cand_keys + static_cast<size_t>(q0) * cap);
        CUDA_TRY(cudaGetLastError());
        FilterSelectParams sp{};
        sp.cand_count = bp.cand_count; sp.keys = c->d_cand_keys + static_cast<size_t>(q0) * cap; sp.cand_cap = cap; sp.k = k_eff;
        sp.out = d_out + static_cast<size_t>(q0) * k_eff; sp.ok = d_ok + q0;
        sp.frame_ids = d_ids; sp.id_base = e->id_base; sp.row_offset = row_offset;
        filter_select_kernel<<<nq, 1024, static_cast<size_t>(cap) * sizeof(uint64_t), stream>>>(sp);
        CUDA_TRY(cudaGetLastError());
        *launches += 3;
    }
    return WAX_VS_OK;
}

// ---- search ---------------------------------------------------------------------------------------------------
// n_queries device-resident queries -> d_out[n_queries][k_eff] on c->stream: the tensor-core levels (bf16 shadow ->
// TF32 retry -> exact scan) when the batch is eligible, else one fused scan per query.  The tensor
// levels read their proof flags back, so they synchronise c->stream; the scan loop only enqueues.
static int32_t run_queries_on_device(wax_vs_engine *e, SearchCtx *c, const float *d_queries, uint32_t n_queries,
                                     uint32_t k_eff, uint64_t row_offset, wax_vs_candidate *d_out, const uint64_t *d_ids,
                                     uint64_t *launches, const uint32_t *d_mask = nullptr) {
    int32_t rc = WAX_VS_OK;
    bool tensor_path = batch_tensor_eligible(e, n_queries, k_eff), allow_bf16 = true;
    if (tensor_path) {
        std::lock_guard<std::mutex> pg(e->pool_mu);
        if (e->bf16_skip_batches > 0) { --e->bf16_skip_batches; allow_bf16 = false; }
    }
    // below batch_min the tensor path only pays off through the bf16 shadow (single_shadow): never TF32 for one query
    if (tensor_path && !allow_bf16 && n_queries < static_cast<uint32_t>(std::max(e->tune.batch_min, 1))) tensor_path = false;
    if (tensor_path) {
        // Batched: one tensor-core pass over the corpus nominates, the finish kernel re-scores exactly and
        // proves completeness; unproven queries (rare) are re-run on the exact single-query path below.
        if ((rc = ensure_dev(&c->d_ok, &c->ok_cap, static_cast<size_t>(n_queries), "proof flags"))) return rc;
        if ((rc = ensure_pinned(&c->h_ok, &c->h_ok_cap, static_cast<size_t>(n_queries), "proof flag staging"))) return rc;
        if ((rc = ensure_dev(&c->d_tau_star, &c->tau_star_cap, static_cast<size_t>(n_queries) * 2, "filter thresholds"))) return rc;
        if ((rc = ensure_pinned(&c->h_tau_star, &c->h_tau_star_cap, static_cast<size_t>(n_queries) * 2, "filter threshold staging"))) return rc;
        bool used_bf16 = false;
        uint32_t used_heap = 0;
        rc = enqueue_batch_tensor(e, c, d_queries, n_queries, k_eff, row_offset, d_out, c->d_ok, d_ids, c->stream, launches,
                                  allow_bf16, &used_bf16, c->d_tau_star, d_mask, &used_heap);
        if (rc) { cudaStreamSynchronize(c->stream); return rc; }
        CUDA_TRY(cudaMemcpyAsync(c->h_ok, c->d_ok, n_queries * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(c->h_tau_star, c->d_tau_star, 2 * n_queries * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        std::vector<uint32_t> unproven;
        for (uint32_t qi = 0; qi < n_queries; ++qi) if (!c->h_ok[qi]) unproven.push_back(qi);
        if (used_bf16 && k_eff <= 128u && unproven.size() * 4 > n_queries) {   // (large k is expected to need the filter level)
            std::lock_guard<std::mutex> pg(e->pool_mu);
            e->bf16_skip_batches = 16;
        }
        if (used_bf16 && k_eff <= 128u && used_heap != 0u && e->tune.batch_heap == 0) {
            // the data has the last word on the heap size: unproven queries -> one size up for `ttl` batches; when the
            // time-out ends one size down is probed again, and a failure during the probe doubles the time-out
            std::lock_guard<std::mutex> pg(e->pool_mu);
            e->last_heap = used_heap;
            if (!unproven.empty()) {
                if (used_heap < 64u) {
                    if (e->heap_probing) e->heap_backoff = std::min<uint32_t>(e->heap_backoff * 2u, 1u << 16);
                    e->heap_bump = std::min<uint32_t>(e->heap_bump + 1u, 3u);
                    e->heap_bump_ttl = e->heap_backoff;
                }
                e->heap_probing = false;
            } else {
                e->heap_probing = false;
                if (e->heap_bump > 0 && e->heap_bump_ttl > 0 && --e->heap_bump_ttl == 0) {
                    --e->heap_bump;
                    e->heap_probing = true;
                    e->heap_bump_ttl = e->heap_bump ? e->heap_backoff : 0;
                }
            }
        }
        uint64_t retried = 0, retried_bf16 = 0;
        // Level 2, the filter levels: the unproven queries that have a finite threshold, as one compacted sub-batch, get
        // ONE more tensor-core pass that keeps no heap -- it lists every row above the query's fixed threshold (complete
        // by construction).  First over the bf16 shadow when level 1 used it (half the bytes of an exact scan, so it pays
        // even for a single query; its wider bound admits more candidates), then -- for the lists that overflowed, and
        // only for sub-batches worth a tensor pass -- in TF32 over the fp32 corpus.  What is left takes the exact scan.
        auto filter_pass = [&](std::vector<uint32_t> &todo, bool bf16lvl) -> int32_t {
            const uint32_t nf = static_cast<uint32_t>(todo.size());
            int32_t frc;
            if ((frc = ensure_dev(&c->d_retry_q, &c->retry_q_cap, static_cast<size_t>(nf) * e->dims, "filter-level queries"))) return frc;
            if ((frc = ensure_dev(&c->d_retry_out, &c->retry_out_cap, static_cast<size_t>(nf) * k_eff, "filter-level results"))) return frc;
            if ((frc = ensure_dev(&c->d_retry_ok, &c->retry_ok_cap, static_cast<size_t>(nf), "filter-level flags"))) return frc;
            if ((frc = ensure_dev(&c->d_filter_tau, &c->filter_tau_cap, static_cast<size_t>(nf), "filter-level thresholds"))) return frc;
            if ((frc = ensure_pinned(&c->h_filter_tau, &c->h_filter_tau_cap, static_cast<size_t>(nf), "filter-level threshold staging"))) return frc;
            const float *taus = c->h_tau_star + (bf16lvl ? n_queries : 0u);      // [0]: TF32 bound, [1]: bf16 bound
            for (uint32_t i = 0; i < nf; ++i) c->h_filter_tau[i] = taus[todo[i]];
            CUDA_TRY(cudaMemcpyAsync(c->d_filter_tau, c->h_filter_tau, nf * sizeof(float), cudaMemcpyHostToDevice, c->stream));
            for (uint32_t i = 0; i < nf; ++i)
                CUDA_TRY(cudaMemcpyAsync(c->d_retry_q + static_cast<size_t>(i) * e->dims,
                                         d_queries + static_cast<size_t>(todo[i]) * e->dims, e->dims * sizeof(float),
                                         cudaMemcpyDeviceToDevice, c->stream));
            frc = enqueue_filter_level(e, c, c->d_retry_q, c->d_filter_tau, nf, k_eff, row_offset, c->d_retry_out,
                                       c->d_retry_ok, d_ids, c->stream, launches, bf16lvl, d_mask);
            if (frc) { cudaStreamSynchronize(c->stream); return frc; }
            CUDA_TRY(cudaMemcpyAsync(c->h_ok, c->d_retry_ok, nf * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            std::vector<uint32_t> overflowed;
            for (uint32_t i = 0; i < nf; ++i) {
                if (c->h_ok[i])
                    CUDA_TRY(cudaMemcpyAsync(d_out + static_cast<size_t>(todo[i]) * k_eff,
                                             c->d_retry_out + static_cast<size_t>(i) * k_eff, k_eff * sizeof(wax_vs_candidate),
                                             cudaMemcpyDeviceToDevice, c->stream));
                else
                    overflowed.push_back(todo[i]);
            }
            todo.swap(overflowed);
            return WAX_VS_OK;
        };
        if (e->tune.batch_retry && !unproven.empty()) {
            std::vector<uint32_t> todo, rest;
            for (uint32_t qi : unproven)
                ((std::isfinite(c->h_tau_star[qi]) && std::isfinite(c->h_tau_star[n_queries + qi])) ? todo : rest).push_back(qi);
            if (!todo.empty() && used_bf16 && e->tune.filter_bf16) {
                retried_bf16 = todo.size();
                if ((rc = filter_pass(todo, true))) return rc;
            }
            if (todo.size() >= static_cast<size_t>(std::max(e->tune.batch_min, 1))) {
                retried = todo.size();
                if ((rc = filter_pass(todo, false))) return rc;
            }
            rest.insert(rest.end(), todo.begin(), todo.end());
            std::sort(rest.begin(), rest.end());
            unproven.swap(rest);
        }
        for (uint32_t qi : unproven) {
            rc = enqueue_search(e, c, d_queries + static_cast<size_t>(qi) * e->dims, k_eff, row_offset,
                                d_out + static_cast<size_t>(qi) * k_eff, d_ids, c->stream, launches, d_mask);
            if (rc) { cudaStreamSynchronize(c->stream); return rc; }
        }
        {
            std::lock_guard<std::mutex> pg(e->pool_mu);
            e->batch_tensor_queries += n_queries - unproven.size();
            e->batch_fallback_queries += unproven.size();
            if (used_bf16) e->batch_bf16_queries += n_queries;
            else e->batch_tf32_queries += n_queries;
            e->batch_retry_queries += retried;
            e->batch_filter_bf16_queries += retried_bf16;
        }
    } else {
        for (uint32_t qi = 0; qi < n_queries; ++qi) {
            rc = enqueue_search(e, c, d_queries + static_cast<size_t>(qi) * e->dims, k_eff, row_offset,
                                d_out + static_cast<size_t>(qi) * k_eff, d_ids, c->stream, launches, d_mask);
            if (rc) { cudaStreamSynchronize(c->stream); return rc; }
        }
    }
    return WAX_VS_OK;
}

static int32_t search_host(wax_vs_engine *e, const float *queries, uint32_t n_queries, uint32_t query_len,
                           int64_t top_k, uint64_t *out_ids, float *out_scores, uint32_t out_stride,
                           uint32_t *out_n) {
    if (!e || !out_n) return fail(WAX_VS_ERR_NULL, "NULL argument");
    std::shared_lock<std::shared_mutex> r(e->rw);
    if (e->n_rows == 0) {  // guard vectorCount > 0 else { return [] } (:448) -- before validation, as the reference
        for (uint32_t i = 0; i < n_queries; ++i) out_n[i] = 0;
        return WAX_VS_OK;
    }
    if (n_queries == 0) return WAX_VS_OK;
    if (!queries) return fail(WAX_VS_ERR_NULL, "query is NULL");
    if (query_len != e->dims)  // validate (:449, :830-833)
        return fail(WAX_VS_ERR_DIMENSION, "vector dimension mismatch: expected %u, got %u", e->dims, query_len);
    const uint32_t limit = clamp_topk(top_k);
    const uint32_t k_eff = static_cast<uint32_t>(std::min<uint64_t>(limit, e->n_rows));  // topKCount (:451)
    if (!out_ids || !out_scores) return fail(WAX_VS_ERR_NULL, "output buffer is NULL");
    if (out_stride < k_eff)
        return fail(WAX_VS_ERR_BUFFER, "output buffers hold %u entries, need %u", out_stride, k_eff);

    DeviceGuard g(e->device);
    if (!g.ok) return fail(WAX_VS_ERR_CUDA, "failed to select CUDA device %d", e->device);
    SearchCtx *c = nullptr;
    int32_t rc = ctx_acquire(e, &c);
    if (rc) return rc;
    struct Rel { wax_vs_engine *e; SearchCtx *c; ~Rel() { ctx_release(e, c); } } rel{e, c};

    const size_t qfloats = static_cast<size_t>(n_queries) * e->dims;
    const size_t ncand = static_cast<size_t>(n_queries) * k_eff;
    // One query on the fused scan: the query rides in the kernel parameters and the kernel itself stores the result in
    // mapped host memory and raises a flag -- no H2D copy, no D2H copy, no stream synchronisation on the way.
    if (n_queries == 1 && e->tune.host_delivery && k_eff <= static_cast<uint32_t>(e->tune.fused_k_max) &&
        !batch_tensor_eligible(e, 1, k_eff)) {
        if ((rc = ensure_dev(&c->d_out, &c->d_out_cap, ncand, "result buffer"))) return rc;
        if ((rc = ensure_pinned(&c->h_out, &c->h_out_cap, ncand, "result staging"))) return rc;
        if (!c->h_flag) {
            CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&c->h_flag), sizeof(unsigned long long), cudaHostAllocMapped | cudaHostAllocPortable));
            *c->h_flag = 0; c->host_seq = 0;
        }
        HostDelivery hd{queries, c->h_out, c->h_flag, ++c->host_seq};
        uint64_t launches = 0;
        if ((rc = enqueue_search(e, c, nullptr, k_eff, 0, c->d_out, nullptr, c->stream, &launches, nullptr, nullptr, &hd))) {
            cudaStreamSynchronize(c->stream);
            return rc;
        }
        if (hd.delivered) {
            if ((rc = wait_host_flag(c->stream, c->h_flag, hd.seq, 30ull * 1000 * 1000 * 1000)) != WAX_VS_OK) {
                cudaStreamSynchronize(c->stream);
                return rc > 0 ? fail(WAX_VS_ERR_CUDA, "search reported a device-side error") : rc;
            }
        } else {
            CUDA_TRY(cudaMemcpyAsync(c->h_out, c->d_out, ncand * sizeof(wax_vs_candidate), cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
        }
        uint32_t m = 0;
        for (uint32_t i = 0; i < k_eff; ++i) {
            const wax_vs_candidate &cd = c->h_out[i];
            if (!cd.valid) continue;
            out_ids[m] = e->ids_identity ? e->id_base + cd.row : e->ids[cd.row];
            out_scores[m] = score_from_distance(e->similarity, cd.distance);
            ++m;
        }
        out_n[0] = m;
        return WAX_VS_OK;
    }
    if ((rc = ensure_dev(&c->d_queries, &c->d_queries_cap, qfloats, "query buffer"))) return rc;
    if ((rc = ensure_pinned(&c->h_queries, &c->h_queries_cap, qfloats, "query staging"))) return rc;
    if ((rc = ensure_dev(&c->d_out, &c->d_out_cap, ncand, "result buffer"))) return rc;
    if ((rc = ensure_pinned(&c->h_out, &c->h_out_cap, ncand, "result staging"))) return rc;

    memcpy(c->h_queries, queries, qfloats * sizeof(float));
    CUDA_TRY(cudaMemcpyAsync(c->d_queries, c->h_queries, qfloats * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    uint64_t launches = 0;
    if ((rc = run_queries_on_device(e, c, c->d_queries, n_queries, k_eff, 0, c->d_out, nullptr, &launches))) return rc;
    CUDA_TRY(cudaMemcpyAsync(c->h_out, c->d_out, ncand * sizeof(wax_vs_candidate), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));

    // row -> frameId and distance -> score on the host, as MetalVectorEngine.swift:595-603 does.
    for (uint32_t qi = 0; qi < n_queries; ++qi) {
        uint32_t m = 0;
        for (uint32_t i = 0; i < k_eff; ++i) {
            const wax_vs_candidate &cd = c->h_out[static_cast<size_t>(qi) * k_eff + i];
            if (!cd.valid) continue;
            const uint64_t row = cd.row;
            out_ids[static_cast<size_t>(qi) * out_stride + m] = e->ids_identity ? e->id_base + row : e->ids[row];
            out_scores[static_cast<size_t>(qi) * out_stride + m] = score_from_distance(e->similarity, cd.distance);
            ++m;
        }
        out_n[qi] = m;
    }
    return WAX_VS_OK;
}

int32_t wax_vs_search(wax_vs_engine *e, const float *query, uint32_t query_len, int64_t top_k,
                      uint64_t *out_ids, float *out_scores, uint32_t out_cap, uint32_t *out_n) {
    return search_host(e, query, 1, query_len, top_k, out_ids, out_scores, out_cap, out_n);
}

int32_t wax_vs_search_batch(wax_vs_engine *e, const float *queries, uint32_t n_queries, uint32_t query_len,
                            int64_t top_k, uint64_t *out_ids, float *out_scores, uint32_t out_stride,
                            uint32_t *out_n) {
    return search_host(e, queries, n_queries, query_len, top_k, out_ids, out_scores, out_stride, out_n);
}

int32_t wax_vs_search_device(wax_vs_engine *e, const float *d_queries, uint32_t n_queries, int64_t top_k,
                             uint64_t row_offset, wax_vs_candidate *d_candidates, void *cuda_stream) {
    if (!e || !d_queries || !d_candidates) return fail(WAX_VS_ERR_NULL, "NULL argument");
    std::shared_lock<std::shared_mutex> r(e->rw);
    DeviceGuard g(e->device);
    if (!g.ok) return fail(WAX_VS_ERR_CUDA, "failed to select CUDA device %d", e->device);
    const uint32_t k_eff = clamp_topk(top_k);
    SearchCtx *c = nullptr;
    int32_t rc = ctx_for_stream(e, cuda_stream, &c);
    if (rc) return rc;
    e->async_pending.store(true);
    const uint64_t *d_ids = nullptr;
    if ((rc = sync_device_ids(e, &d_ids))) return rc;
    uint64_t launches = 0;
    for (uint32_t qi = 0; qi < n_queries; ++qi) {
        rc = enqueue_search(e, c, d_queries + static_cast<size_t>(qi) * e->dims, k_eff, row_offset,
                            d_candidates + static_cast<size_t>(qi) * k_eff, d_ids,
                            static_cast<cudaStream_t>(cuda_stream), &launches);
        if (rc) return rc;
    }
    return WAX_VS_OK;
}

// Batched form of wax_vs_search_device: the tensor-core levels on the caller's stream for the rank's shard.  Unlike
// the single-query form it may SYNCHRONISE the stream (the proof flags are read back before the unproven queries
// are re-run), so on return d_candidates is complete on `cuda_stream` order and usually already materialised.
int32_t wax_vs_search_batch_device(wax_vs_engine *e, const float *d_queries, uint32_t n_queries, int64_t top_k,
                                   uint64_t row_offset, wax_vs_candidate *d_candidates, void *cuda_stream) {
    if (!e || !d_queries || !d_candidates) return fail(WAX_VS_ERR_NULL, "NULL argument");
    if (n_queries == 0) return WAX_VS_OK;
    std::shared_lock<std::shared_mutex> r(e->rw);
    DeviceGuard g(e->device);
    if (!g.ok) return fail(WAX_VS_ERR_CUDA, "failed to select CUDA device %d", e->device);
    const uint32_t k_eff = clamp_topk(top_k);
    SearchCtx *c = nullptr;
    int32_t rc = ctx_for_stream(e, cuda_stream, &c);
    if (rc) return rc;
    e->async_pending.store(true);
    const uint64_t *d_ids = nullptr;
    if ((rc = sync_device_ids(e, &d_ids))) return rc;
    uint64_t launches = 0;
    if (k_eff > e->n_rows) {   // a shard smaller than k: the scan pads with invalid candidates, the tensor path does not
        for (uint32_t qi = 0; qi < n_queries; ++qi) {
            rc = enqueue_search(e, c, d_queries + static_cast<size_t>(qi) * e->dims, k_eff, row_offset,
                                d_candidates + static_cast<size_t>(qi) * k_eff, d_ids, c->stream, &launches);
            if (rc) return rc;
        }
        return WAX_VS_OK;
    }
    return run_queries_on_device(e, c, d_queries, n_queries, k_eff, row_offset, d_candidates, d_ids, &launches);
}

// ---- row-sharded search: fused scan + NVLink exchange + merge (waxvs_shard.cuh) ---------------
// Handle blob exchanged between the ranks (WAX_VS_SHARD_HANDLE_BYTES): how a peer reaches this rank's mailbox.
struct ShardHandle {
    uint32_t magic, version;
    int32_t pid, device;
    uint64_t nonce;                  // per-process random: same (pid, nonce) = same process -> plain peer access
    uint64_t ptr;                    // mailbox device pointer (valid in the owning process)
    int32_t rank, world;
    cudaIpcMemHandle_t ipc;          // 64 bytes: for other processes on the node
    uint8_t pad[WAX_VS_SHARD_HANDLE_BYTES - 40 - sizeof(cudaIpcMemHandle_t)];
};
static_assert(sizeof(ShardHandle) == WAX_VS_SHARD_HANDLE_BYTES, "handle blob size");
static uint64_t process_nonce() {
    static const uint64_t n = [] { std::random_device rd; return (static_cast<uint64_t>(rd()) << 32) ^ rd() ^ 0x9E3779B97F4A7C15ull; }();
    return n;
}

// Caller holds the write lock, device selected.  Two steps because other PROCESSES may still have this rank's mailbox
// mapped: wax_vs_shard_close only unmaps the peers' mailboxes (free_own = false); the own mailbox is released when the
// engine is destroyed or re-opened, i.e. after the group has agreed (a barrier on the caller's side) that everyone
// has closed.
static void shard_teardown(wax_vs_engine *e, bool free_own) {
    auto &sh = e->shard;
    if (!sh.open) return;
    cudaDeviceSynchronize();
    for (int r = 0; r < sh.world; ++r) {
        if (r != sh.rank && sh.ipc[r] && sh.box[r]) cudaIpcCloseMemHandle(sh.box[r]);
        sh.ipc[r] = false;
        if (r != sh.rank) sh.box[r] = nullptr;
    }
    sh.connected = false;
    cudaGetLastError();
    if (!free_own) return;
    if (sh.box[sh.rank]) cudaFree(sh.box[sh.rank]);
    sh.box[sh.rank] = nullptr;
    if (sh.ctx) { ctx_free(sh.ctx); sh.ctx = nullptr; }
    if (sh.d_final) { cudaFree(sh.d_final); sh.d_final = nullptr; }
    if (sh.h_final) { cudaFreeHost(sh.h_final); sh.h_final = nullptr; }
    if (sh.h_flag) { cudaFreeHost(sh.h_flag); sh.h_flag = nullptr; }
    cudaGetLastError();
    sh.open = false;
    sh.seq = 0;
}

int32_t wax_vs_shard_open(wax_vs_engine *e, int32_t rank, int32_t world, uint64_t row_offset, uint8_t *out_handle) {
    if (!e || !out_handle) return fail(WAX_VS_ERR_NULL, "NULL argument");
    if (world < 1 || world > kShardMaxRanks || rank < 0 || rank >= world)
        return fail(WAX_VS_ERR_ARGUMENT, "rank %d of %d: world must be 1..%d", rank, world, kShardMaxRanks);
    std::unique_lock<std::shared_mutex> w(e->rw);
    DeviceGuard g(e->device);
    if (!g.ok) return fail(WAX_VS_ERR_CUDA, "failed to select CUDA device %d", e->device);
    drain_device_path(e);
    shard_teardown(e, true);
    auto &sh = e->shard;
    ShardMailbox *box = nullptr;
    CUDA_TRY(cudaMalloc(&box, sizeof(ShardMailbox)));
    cudaError_t err = cudaMemset(box, 0, sizeof(ShardMailbox));
    if (err == cudaSuccess) err = cudaMalloc(&sh.d_final, kShardKCap * sizeof(wax_vs_candidate));
    if (err == cudaSuccess) err = cudaHostAlloc(&sh.h_final, kShardKCap * sizeof(wax_vs_candidate), cudaHostAllocMapped | cudaHostAllocPortable);
    if (err == cudaSuccess) err = cudaHostAlloc(&sh.h_flag, sizeof(unsigned long long), cudaHostAllocMapped | cudaHostAllocPortable);
    if (err == cudaSuccess) err = cudaDeviceSynchronize();
    ShardHandle h{};
    if (err == cudaSuccess) err = cudaIpcGetMemHandle(&h.ipc, box);
    if (err != cudaSuccess) {
        cudaFree(box);
        if (sh.d_final) { cudaFree(sh.d_final); sh.d_final = nullptr; }
        if (sh.h_final) { cudaFreeHost(sh.h_final); sh.h_final = nullptr; }
        if (sh.h_flag) { cudaFreeHost(sh.h_flag); sh.h_flag = nullptr; }
        return fail(WAX_VS_ERR_CUDA, "failed to create the shard mailbox: %s", cudaGetErrorString(err));
    }
    *sh.h_flag = 0;
    int32_t rc = ctx_new(e, &sh.ctx, true);
    if (rc) { cudaFree(box); return rc; }
    sh.rank = rank; sh.world = world; sh.row_offset = row_offset;
    sh.box[rank] = box;
    sh.open = true; sh.connected = (world == 1);
    sh.seq = 0;
    h.magic = 0x48535857u; h.version = 1; h.pid = static_cast<int32_t>(getpid()); h.device = e->device;
    h.nonce = process_nonce(); h.ptr = reinterpret_cast<uint64_t>(box); h.rank = rank; h.world = world;
    memcpy(out_handle, &h, sizeof h);
    return WAX_VS_OK;
}

int32_t wax_vs_shard_connect(wax_vs_engine *e, const uint8_t *handles, int32_t n_handles) {
    if (!e || !handles) return fail(WAX_VS_ERR_NULL, "NULL argument");
    std::unique_lock<std::shared_mutex> w(e->rw);
    auto &sh = e->shard;
    if (!sh.open) return fail(WAX_VS_ERR_ARGUMENT, "wax_vs_shard_open has not been called");
    if (n_handles != sh.world) return fail(WAX_VS_ERR_ARGUMENT, "expected %d handles, got %d", sh.world, n_handles);
    DeviceGuard g(e->device);
    if (!g.ok) return fail(WAX_VS_ERR_CUDA, "failed to select CUDA device %d", e->device);
    for (int r = 0; r < sh.world; ++r) {
        ShardHandle h;
        memcpy(&h, handles + static_cast<size_t>(r) * sizeof h, sizeof h);
        if (h.magic != 0x48535857u || h.version != 1 || h.rank != r || h.world != sh.world)
            return fail(WAX_VS_ERR_ARGUMENT, "handle %d is not rank %d of %d", r, r, sh.world);
        if (r == sh.rank) continue;
        if (h.pid == static_cast<int32_t>(getpid()) && h.nonce == process_nonce()) {
            // same process (several engines in one host process): plain peer access to the other device
            if (h.device != e->device) {
                int can = 0;
                CUDA_TRY(cudaDeviceCanAccessPeer(&can, e->device, h.device));
                if (!can) return fail(WAX_VS_ERR_UNSUPPORTED, "device %d cannot access device %d (no P2P path)", e->device, h.device);
                cudaError_t pe = cudaDeviceEnablePeerAccess(h.device, 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled)
                    return fail(WAX_VS_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d) failed: %s", h.device, cudaGetErrorString(pe));
                cudaGetLastError();
            }
            sh.box[r] = reinterpret_cast<ShardMailbox *>(h.ptr);
            sh.ipc[r] = false;
        } else {
            void *ptr = nullptr;
            cudaError_t ie = cudaIpcOpenMemHandle(&ptr, h.ipc, cudaIpcMemLazyEnablePeerAccess);
            if (ie != cudaSuccess) {
                cudaGetLastError();
                return fail(WAX_VS_ERR_UNSUPPORTED, "cudaIpcOpenMemHandle for rank %d failed: %s (no P2P/IPC path between the ranks)",
                            r, cudaGetErrorString(ie));
            }
            sh.box[r] = static_cast<ShardMailbox *>(ptr);
            sh.ipc[r] = true;
        }
    }
    sh.connected = true;
    return WAX_VS_OK;
}

int32_t wax_vs_shard_close(wax_vs_engine *e) {
    if (!e) return fail(WAX_VS_ERR_NULL, "engine is NULL");
    std::unique_lock<std::shared_mutex> w(e->rw);
    DeviceGuard g(e->device);
    shard_teardown(e, false);
    return WAX_VS_OK;
}

// caller holds e->shard.mu: the next collective sequence number and the parameter block that goes with it
static ShardParams shard_params_next(wax_vs_engine *e) {
    auto &sh = e->shard;
    ShardParams sp{};
    for (int r = 0; r < sh.world; ++r) sp.box[r] = sh.box[r];
    sp.rank = static_cast<uint32_t>(sh.rank); sp.world = static_cast<uint32_t>(sh.world);
    sp.seq = ++sh.seq;
    sp.timeout_ns = sh.timeout_ns;
    return sp;
}

int32_t wax_vs_shard_search_device(wax_vs_engine *e, const float *d_query, int64_t top_k, wax_vs_candidate *d_candidates,
                                   void *cuda_stream) {
    if (!e || !d_query || !d_candidates) return fail(WAX_VS_ERR_NULL, "NULL argument");
    std::shared_lock<std::shared_mutex> r(e->rw);
    if (!e->shard.connected) return fail(WAX_VS_ERR_ARGUMENT, "the shard group is not connected (wax_vs_shard_open / _connect)");
    DeviceGuard g(e->device);
    if (!g.ok) return fail(WAX_VS_ERR_CUDA, "failed to select CUDA device %d", e->device);
    const uint32_t k_eff = clamp_topk(top_k);
    if (k_eff > static_cast<uint32_t>(kShardKCap))
        return fail(WAX_VS_ERR_UNSUPPORTED, "sharded search supports top_k <= %d (got %u)", kShardKCap, k_eff);
    SearchCtx *c = nullptr;
    int32_t rc = ctx_for_stream(e, cuda_stream, &c);
    if (rc) return rc;
    e->async_pending.store(true);
    const uint64_t *d_ids = nullptr;
    if ((rc = sync_device_ids(e, &d_ids))) return rc;
    std::lock_guard<std::mutex> sg(e->shard.mu);
    ShardParams sp = shard_params_next(e);
    uint64_t launches = 0;
    return enqueue_search(e, c, d_query, k_eff, e->shard.row_offset, d_candidates, d_ids,
                          static_cast<cudaStream_t>(cuda_stream), &launches, nullptr, &sp);
}

static int32_t shard_wait_host(wax_vs_engine *e, unsigned long long seq) {
    int32_t rc = wait_host_flag(e->shard.ctx->stream, e->shard.h_flag, seq, e->shard.timeout_ns);
    if (rc == 1)
        return fail(WAX_VS_ERR_CUDA, "shard exchange timed out: a peer rank did not deliver its candidates for query #%llu", seq);
    return rc;
}

int32_t wax_vs_shard_search(wax_vs_engine *e, const float *query, uint32_t query_len, int64_t top_k, uint64_t *out_ids,
                            float *out_scores, uint32_t out_cap, uint32_t *out_n) {
    if (!e || !out_n) return fail(WAX_VS_ERR_NULL, "NULL argument");
    std::shared_lock<std::shared_mutex> r(e->rw);
    *out_n = 0;
    if (!e->shard.connected) return fail(WAX_VS_ERR_ARGUMENT, "the shard group is not connected (wax_vs_shard_open / _connect)");
    if (!query) return fail(WAX_VS_ERR_NULL, "query is NULL");
    if (query_len != e->dims)
        return fail(WAX_VS_ERR_DIMENSION, "vector dimension mismatch: expected %u, got %u", e->dims, query_len);
    const uint32_t k_eff = clamp_topk(top_k);
    if (k_eff > static_cast<uint32_t>(kShardKCap))
        return fail(WAX_VS_ERR_UNSUPPORTED, "sharded search supports top_k <= %d (got %u)", kShardKCap, k_eff);
    if (!out_ids || !out_scores) return fail(WAX_VS_ERR_NULL, "output buffer is NULL");
    DeviceGuard g(e->device);
    if (!g.ok) return fail(WAX_VS_ERR_CUDA, "failed to select CUDA device %d", e->device);
    auto &sh = e->shard;
    std::lock_guard<std::mutex> sg(sh.mu);      // one host-path collective at a time: it owns sh.ctx and h_final
    SearchCtx *c = sh.ctx;
    int32_t rc;
    const uint64_t *d_ids = nullptr;
    if ((rc = sync_device_ids(e, &d_ids))) return rc;
    ShardParams sp = shard_params_next(e);
    sp.host_out = sh.h_final; sp.host_flag = sh.h_flag;      // mapped pinned: the kernel delivers the result itself
    HostDelivery hd{query, nullptr, nullptr, 0};             // the query rides in the kernel parameters when it fits
    uint64_t launches = 0;
    if ((rc = enqueue_search(e, c, nullptr, k_eff, sh.row_offset, sh.d_final, d_ids, c->stream, &launches, nullptr, &sp, &hd))) {
        cudaStreamSynchronize(c->stream);
        return rc;
    }
    if ((rc = shard_wait_host(e, sp.seq))) { cudaStreamSynchronize(c->stream); return rc; }
    uint32_t m = 0;
    for (uint32_t i = 0; i < k_eff; ++i) {
        const wax_vs_candidate &cd = sh.h_final[i];
        if (cd.valid != 1u) continue;
        if (m >= out_cap) return fail(WAX_VS_ERR_BUFFER, "output buffers hold %u entries, need more", out_cap);
        out_ids[m] = cd.frame_id;
        out_scores[m] = score_from_distance(e->similarity, cd.distance);
        ++m;
    }
    *out_n = m;
    return WAX_VS_OK;
}

// Device-timed sharded searches, strictly one query at a time on one stream (the same mode as wax_vs_debug_time_search):
// `n_queries` unit queries generated on device from generator stream `seed` (identical on every rank), warmup + iters
// collective searches back to back, CUDA events around the `iters`.  Every rank must make the same call.
int32_t wax_vs_debug_time_shard_search(wax_vs_engine *e, uint32_t n_queries, int64_t top_k, uint64_t seed, uint32_t warmup,
                                       uint32_t iters, float *out_ms_total, uint64_t *out_launches) {
    if (!e || !out_ms_total) return fail(WAX_VS_ERR_NULL, "NULL argument");
    if (n_queries == 0) n_queries = 1;
    std::shared_lock<std::shared_mutex> r(e->rw);
    if (!e->shard.connected) return fail(WAX_VS_ERR_ARGUMENT, "the shard group is not connected (wax_vs_shard_open / _connect)");
    DeviceGuard g(e->device);
    auto &sh = e->shard;
    std::lock_guard<std::mutex> sg(sh.mu);
    SearchCtx *c = sh.ctx;
    const uint32_t k_eff = clamp_topk(top_k);
    int32_t rc;
    if ((rc = ensure_dev(&c->d_queries, &c->d_queries_cap, static_cast<size_t>(n_queries) * e->dims, "query buffer"))) return rc;
    synth_fill_kernel<<<(n_queries + 255) / 256, 256, 0, c->stream>>>(c->d_queries, n_queries, e->dims, seed, 0, 1);
    CUDA_TRY(cudaGetLastError());
    const uint64_t *d_ids = nullptr;
    if ((rc = sync_device_ids(e, &d_ids))) return rc;
    uint64_t launches = 0;
    for (uint32_t it = 0; it < warmup + iters; ++it) {
        if (it == warmup) { launches = 0; CUDA_TRY(cudaEventRecord(c->ev0, c->stream)); }
        ShardParams sp = shard_params_next(e);
        rc = enqueue_search(e, c, c->d_queries + static_cast<size_t>(it % n_queries) * e->dims, k_eff, sh.row_offset,
                            sh.d_final, d_ids, c->stream, &launches, nullptr, &sp);
        if (rc) { cudaStreamSynchronize(c->stream); return rc; }
    }
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaEventElapsedTime(out_ms_total, c->ev0, c->ev1));
    if (out_launches) *out_launches = launches;
    return WAX_VS_OK;
}

// Device-side merge of all-gathered per-rank candidate lists (the sharded search_batch): stateless, enqueued on the
// caller's stream, no synchronisation.
int32_t wax_vs_merge_candidates_device(wax_vs_engine *e, const wax_vs_candidate *d_gathered, uint32_t world,
                                       uint32_t n_queries, uint32_t k, uint32_t k_out, wax_vs_candidate *d_out,
                                       void *cuda_stream) {
    if (!e || !d_gathered || !d_out) return fail(WAX_VS_ERR_NULL, "NULL argument");
    if (world == 0 || world > 1024u || k == 0 || k_out == 0 || k_out > k)
        return fail(WAX_VS_ERR_ARGUMENT, "merge: world %u, k %u, k_out %u", world, k, k_out);
    if (n_queries == 0) return WAX_VS_OK;
    DeviceGuard g(e->device);
    if (!g.ok) return fail(WAX_VS_ERR_CUDA, "failed to select CUDA device %d", e->device);
    merge_gathered_kernel<<<n_queries, 128, 0, static_cast<cudaStream_t>(cuda_stream)>>>(d_gathered, world, n_queries, k, k_out, d_out);
    CUDA_TRY(cudaGetLastError());
    return WAX_VS_OK;
}

// ---- filtered search -------------------------------------------------------------------
// The reference filters AFTER the engine call and over-fetches 3 x topK to compensate (UnifiedSearch.swift:58,
// 371-442, 1195-1200, 1241-1258).  Here the filter is pushed below the top-k: a row bitset consulted only for rows
// that would enter the list, or -- for small allow-lists -- a gather that scores only the listed rows.
// frameIds -> rows of this engine (unknown ids are ignored), as a bitset (bit set = the row may be returned) and as the
// list of rows the ids named.  Returns the number of rows that may be returned.
static uint64_t build_row_filter(wax_vs_engine *e, const uint64_t *frame_ids, uint64_t n_ids, int32_t mode,
                                 std::vector<uint32_t> &bits, std::vector<uint32_t> &listed) {
    const uint64_t n_rows = e->n_rows;
    bits.assign((n_rows + 31) / 32, mode == 0 ? 0u : 0xFFFFFFFFu);
    if (mode == 1 && (n_rows & 31u)) bits.back() = (1u << (n_rows & 31u)) - 1u;
    listed.clear();
    std::lock_guard<std::mutex> g(e->ids_mu);   // the lazily built id map is shared by concurrent readers
    // (find_row builds the lazily constructed hash table when it is needed: serialised by ids_mu)
    for (uint64_t i = 0; i < n_ids; ++i) {
        uint64_t row;
        if (e->ids_identity) {
            if (frame_ids[i] < e->id_base || frame_ids[i] - e->id_base >= n_rows) continue;
            row = frame_ids[i] - e->id_base;
        } else {
            const uint32_t f = find_row(e, frame_ids[i]);
            if (f == 0xFFFFFFFFu) continue;
            row = f;
        }
        const uint32_t w = static_cast<uint32_t>(row >> 5), b = 1u << (row & 31u);
        if (mode == 0) { if (!(bits[w] & b)) { bits[w] |= b; listed.push_back(static_cast<uint32_t>(row)); } }
        else if (bits[w] & b) { bits[w] &= ~b; listed.push_back(static_cast<uint32_t>(row)); }
    }
    return mode == 0 ? listed.size() : n_rows - listed.size();
}

// One filter, n_queries queries.  Small allow-lists: gather + exact score of the listed rows only (grid.y = query), one
// CTA per query sorts.  Otherwise the row bitset rides below the top-k: in the fused scan (one query, or a batch the
// tensor path cannot take) or in the tensor-core levels (nominations, filter level and the exact fall-back all consult
// the same bitset, so the completeness proof is a statement about the ALLOWED rows).
static int32_t search_filtered_host(wax_vs_engine *e, const float *queries, uint32_t n_queries, uint32_t query_len,
                                    int64_t top_k, const uint64_t *frame_ids, uint64_t n_ids, int32_t mode,
                                    uint64_t *out_ids, float *out_scores, uint32_t out_stride, uint32_t *out_n) {
    if (!e || !out_n) return fail(WAX_VS_ERR_NULL, "NULL argument");
    if (mode != 0 && mode != 1) return fail(WAX_VS_ERR_ARGUMENT, "filter mode must be 0 (allow-list) or 1 (deny-list)");
    if (n_ids && !frame_ids) return fail(WAX_VS_ERR_NULL, "frame_ids is NULL");
    std::shared_lock<std::shared_mutex> r(e->rw);
    for (uint32_t i = 0; i < n_queries; ++i) out_n[i] = 0;
    if (e->n_rows == 0 || n_queries == 0) return WAX_VS_OK;
    if (!queries) return fail(WAX_VS_ERR_NULL, "query is NULL");
    if (query_len != e->dims)
        return fail(WAX_VS_ERR_DIMENSION, "vector dimension mismatch: expected %u, got %u", e->dims, query_len);

    std::vector<uint32_t> bits, listed;
    const uint64_t allowed = build_row_filter(e, frame_ids, n_ids, mode, bits, listed);
    const uint32_t k_eff = static_cast<uint32_t>(std::min<uint64_t>(clamp_topk(top_k), allowed));
    if (k_eff == 0) return WAX_VS_OK;
    if (!out_ids || !out_scores) return fail(WAX_VS_ERR_NULL, "output buffer is NULL");
    if (out_stride < k_eff) return fail(WAX_VS_ERR_BUFFER, "output buffers hold %u entries, need %u", out_stride, k_eff);

    DeviceGuard g(e->device);
    if (!g.ok) return fail(WAX_VS_ERR_CUDA, "failed to select CUDA device %d", e->device);
    SearchCtx *c = nullptr;
    int32_t rc = ctx_acquire(e, &c);
    if (rc) return rc;
    struct Rel { wax_vs_engine *e; SearchCtx *c; ~Rel() { ctx_release(e, c); } } rel{e, c};
    const size_t qfloats = static_cast<size_t>(n_queries) * e->dims;
    const size_t ncand = static_cast<size_t>(n_queries) * k_eff;
    if ((rc = ensure_dev(&c->d_queries, &c->d_queries_cap, qfloats, "query buffer"))) return rc;
    if ((rc = ensure_pinned(&c->h_queries, &c->h_queries_cap, qfloats, "query staging"))) return rc;
    if ((rc = ensure_dev(&c->d_out, &c->d_out_cap, ncand, "result buffer"))) return rc;
    if ((rc = ensure_pinned(&c->h_out, &c->h_out_cap, ncand, "result staging"))) return rc;
    memcpy(c->h_queries, queries, qfloats * sizeof(float));
    CUDA_TRY(cudaMemcpyAsync(c->d_queries, c->h_queries, qfloats * sizeof(float), cudaMemcpyHostToDevice, c->stream));

    uint64_t launches = 0;
    if (mode == 0 && listed.size() <= 16384) {
        // small allow-list: gather + exact score of the listed rows only, then one-CTA sorts
        std::sort(listed.begin(), listed.end());
        const uint32_t n = static_cast<uint32_t>(listed.size());
        if ((rc = ensure_dev(&c->d_mask, &c->mask_cap, static_cast<size_t>(std::max<uint32_t>(n, 1)), "listed rows"))) return rc;
        if ((rc = ensure_dev(&c->d_gather_keys, &c->gather_cap, static_cast<size_t>(std::max<uint32_t>(n, 1)) * n_queries, "gather keys"))) return rc;
        CUDA_TRY(cudaMemcpyAsync(c->d_mask, listed.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        uint32_t pow2 = 64;
        while (pow2 < n) pow2 <<= 1;
        {
            std::lock_guard<std::mutex> ag(e->attr_mu);
            if (!e->gather_attr_set) {
                CUDA_TRY(cudaFuncSetAttribute(gather_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
                e->gather_attr_set = true;
            }
        }
        const uint32_t gx = std::max<uint32_t>(1, std::min<uint32_t>((n + 31) / 32, static_cast<uint32_t>(e->sm_count) * 8));
        for (uint32_t q0 = 0; q0 < n_queries; q0 += 32768u) {       // grid.y limit
            const uint32_t nq = std::min<uint32_t>(n_queries - q0, 32768u);
            const dim3 ggrid(gx, nq);
            const float *dq = c->d_queries + static_cast<size_t>(q0) * e->dims;
            uint64_t *keys = c->d_gather_keys + static_cast<size_t>(q0) * n;
            switch (e->similarity) {
                case WAX_VS_COSINE: gather_score_kernel<kCosine><<<ggrid, 256, 0, c->stream>>>(e->d_corpus, dq, e->dims, c->d_mask, n, keys); break;
                case WAX_VS_DOT: gather_score_kernel<kDot><<<ggrid, 256, 0, c->stream>>>(e->d_corpus, dq, e->dims, c->d_mask, n, keys); break;
                default: gather_score_kernel<kL2><<<ggrid, 256, 0, c->stream>>>(e->d_corpus, dq, e->dims, c->d_mask, n, keys); break;
            }
            CUDA_TRY(cudaGetLastError());
            ScanParams sp{};
            sp.k = k_eff; sp.out = c->d_out + static_cast<size_t>(q0) * k_eff; sp.id_base = e->id_base;
            gather_sort_kernel<<<nq, 1024, pow2 * sizeof(uint64_t), c->stream>>>(keys, n, pow2, sp);
            CUDA_TRY(cudaGetLastError());
            launches += 2;
        }
    } else {
        if ((rc = ensure_dev(&c->d_mask, &c->mask_cap, bits.size(), "row filter"))) return rc;
        CUDA_TRY(cudaMemcpyAsync(c->d_mask, bits.data(), bits.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        if (n_queries == 1) {
            rc = enqueue_search(e, c, c->d_queries, k_eff, 0, c->d_out, nullptr, c->stream, &launches, c->d_mask);
            if (rc) { cudaStreamSynchronize(c->stream); return rc; }
        } else if ((rc = run_queries_on_device(e, c, c->d_queries, n_queries, k_eff, 0, c->d_out, nullptr, &launches, c->d_mask))) {
            return rc;
        }
    }
    CUDA_TRY(cudaMemcpyAsync(c->h_out, c->d_out, ncand * sizeof(wax_vs_candidate), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));   // also keeps `bits` / `listed` alive until the copies are done
    for (uint32_t qi = 0; qi < n_queries; ++qi) {
        uint32_t m = 0;
        for (uint32_t i = 0; i < k_eff; ++i) {
            const wax_vs_candidate &cd = c->h_out[static_cast<size_t>(qi) * k_eff + i];
            if (!cd.valid) continue;
            out_ids[static_cast<size_t>(qi) * out_stride + m] = e->ids_identity ? e->id_base + cd.row : e->ids[cd.row];
            out_scores[static_cast<size_t>(qi) * out_stride + m] = score_from_distance(e->similarity, cd.distance);
            ++m;
        }
        out_n[qi] = m;
    }
    return WAX_VS_OK;
}

int32_t wax_vs_search_filtered(wax_vs_engine *e, const float *query, uint32_t query_len, int64_t top_k,
                               const uint64_t *frame_ids, uint64_t n_ids, int32_t mode, uint64_t *out_ids,
                               float *out_scores, uint32_t out_cap, uint32_t *out_n) {
    return search_filtered_host(e, query, 1, query_len, top_k, frame_ids, n_ids, mode, out_ids, out_scores, out_cap, out_n);
}

int32_t wax_vs_search_batch_filtered(wax_vs_engine *e, const float *queries, uint32_t n_queries, uint32_t query_len,
                                     int64_t top_k, const uint64_t *frame_ids, uint64_t n_ids, int32_t mode,
                                     uint64_t *out_ids, float *out_scores, uint32_t out_stride, uint32_t *out_n) {
    return search_filtered_host(e, queries, n_queries, query_len, top_k, frame_ids, n_ids, mode, out_ids, out_scores,
                                out_stride, out_n);
}

// The row-sharded form: every rank passes the SAME ids; a rank resolves the ones its shard holds (the others are
// unknown to it and ignored), its fused scan consults the bitset below the top-k, and the usual in-kernel exchange
// merges the ranks' lists -- the answer is the filtered top-k of the whole corpus, identical on every rank.
int32_t wax_vs_shard_search_filtered(wax_vs_engine *e, const float *query, uint32_t query_len, int64_t top_k,
                                     const uint64_t *frame_ids, uint64_t n_ids, int32_t mode, uint64_t *out_ids,
                                     float *out_scores, uint32_t out_cap, uint32_t *out_n) {
    if (!e || !out_n) return fail(WAX_VS_ERR_NULL, "NULL argument");
    if (mode != 0 && mode != 1) return fail(WAX_VS_ERR_ARGUMENT, "filter mode must be 0 (allow-list) or 1 (deny-list)");
    if (n_ids && !frame_ids) return fail(WAX_VS_ERR_NULL, "frame_ids is NULL");
    std::shared_lock<std::shared_mutex> r(e->rw);
    *out_n = 0;
    if (!e->shard.connected) return fail(WAX_VS_ERR_ARGUMENT, "the shard group is not connected (wax_vs_shard_open / _connect)");
    if (!query) return fail(WAX_VS_ERR_NULL, "query is NULL");
    if (query_len != e->dims)
        return fail(WAX_VS_ERR_DIMENSION, "vector dimension mismatch: expected %u, got %u", e->dims, query_len);
    const uint32_t k_eff = clamp_topk(top_k);
    if (k_eff > static_cast<uint32_t>(kShardKCap))
        return fail(WAX_VS_ERR_UNSUPPORTED, "sharded search supports top_k <= %d (got %u)", kShardKCap, k_eff);
    if (!out_ids || !out_scores) return fail(WAX_VS_ERR_NULL, "output buffer is NULL");
    DeviceGuard g(e->device);
    if (!g.ok) return fail(WAX_VS_ERR_CUDA, "failed to select CUDA device %d", e->device);
    std::vector<uint32_t> bits, listed;
    if (e->n_rows) build_row_filter(e, frame_ids, n_ids, mode, bits, listed);
    auto &sh = e->shard;
    std::lock_guard<std::mutex> sg(sh.mu);      // one host-path collective at a time: it owns sh.ctx and h_final
    SearchCtx *c = sh.ctx;
    int32_t rc;
    const uint64_t *d_ids = nullptr;
    if ((rc = sync_device_ids(e, &d_ids))) return rc;
    if (!bits.empty()) {
        if ((rc = ensure_dev(&c->d_mask, &c->mask_cap, bits.size(), "row filter"))) return rc;
        CUDA_TRY(cudaMemcpyAsync(c->d_mask, bits.data(), bits.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    }
    ShardParams sp = shard_params_next(e);
    sp.host_out = sh.h_final; sp.host_flag = sh.h_flag;
    HostDelivery hd{query, nullptr, nullptr, 0};
    uint64_t launches = 0;
    if ((rc = enqueue_search(e, c, nullptr, k_eff, sh.row_offset, sh.d_final, d_ids, c->stream, &launches,
                             bits.empty() ? nullptr : c->d_mask, &sp, &hd))) {
        cudaStreamSynchronize(c->stream);
        return rc;
    }
    if ((rc = shard_wait_host(e, sp.seq))) { cudaStreamSynchronize(c->stream); return rc; }
    CUDA_TRY(cudaStreamSynchronize(c->stream));   // `bits` must outlive its upload
    uint32_t m = 0;
    for (uint32_t i = 0; i < k_eff; ++i) {
        const wax_vs_candidate &cd = sh.h_final[i];
        if (cd.valid != 1u) continue;
        if (m >= out_cap) return fail(WAX_VS_ERR_BUFFER, "output buffers hold %u entries, need more", out_cap);
        out_ids[m] = cd.frame_id;
        out_scores[m] = score_from_distance(e->similarity, cd.distance);
        ++m;
    }
    *out_n = m;
    return WAX_VS_OK;
}

// ---- persistence ---------------------------------------------------------------------------------------------
static uint64_t mv2v_length(const wax_vs_engine *e) {
    return 36ull + e->n_rows * e->dims * 4ull + 8ull + e->n_rows * 8ull;
}

int32_t wax_vs_serialized_length(wax_vs_engine *e, uint64_t *out) {
    if (!e || !out) return fail(WAX_VS_ERR_NULL, "NULL argument");
    std::shared_lock<std::shared_mutex> r(e->rw);
    *out = mv2v_length(e);
    return WAX_VS_OK;
}

int32_t wax_vs_serialize(wax_vs_engine *e, uint8_t *dst, uint64_t cap, uint64_t *out_len) {
    if (!e || !dst) return fail(WAX_VS_ERR_NULL, "NULL argument");
    std::shared_lock<std::shared_mutex> r(e->rw);
    const uint64_t need = mv2v_length(e);
    if (out_len) *out_len = need;
    if (cap < need) return fail(WAX_VS_ERR_BUFFER, "serialize needs %llu bytes, buffer has %llu",
                                static_cast<unsigned long long>(need), static_cast<unsigned long long>(cap));
    DeviceGuard g(e->device);
    uint8_t *p = dst;
    const uint8_t magic[4] = {0x4D, 0x56, 0x32, 0x56};  // "MV2V" (:686)
    memcpy(p, magic, 4); p += 4;
    const uint16_t version = 1; memcpy(p, &version, 2); p += 2;  // :687-688
    *p++ = 2;                                                     // encoding (:689)
    *p++ = e->similarity;                                         // :690
    memcpy(p, &e->dims, 4); p += 4;                               // :691-692
    memcpy(p, &e->n_rows, 8); p += 8;                             // :693-694
    const uint64_t vbytes = e->n_rows * e->dims * 4ull;
    memcpy(p, &vbytes, 8); p += 8;                                // :697-699
    memset(p, 0, 8); p += 8;                                      // reserved (:700)
    if (vbytes) {                                                 // :703-705, through the pinned double-buffered D2H pipeline
        std::lock_guard<std::mutex> ig(e->ingest_mu);             // serialize holds only the READ lock: one exporter at a time
        int32_t rc = download_bytes(e, p, e->d_corpus, vbytes);
        if (rc) return rc;
    }
    p += vbytes;
    const uint64_t ibytes = e->n_rows * 8ull;
    memcpy(p, &ibytes, 8); p += 8;                                // :707-709
    if (e->ids_identity) {
        for (uint64_t i = 0; i < e->n_rows; ++i) { const uint64_t id = e->id_base + i; memcpy(p + i * 8, &id, 8); }
    } else if (ibytes) {
        memcpy(p, e->ids.data(), ibytes);                         // :710
    }
    return WAX_VS_OK;
}

int32_t wax_vs_deserialize(wax_vs_engine *e, const uint8_t *src, uint64_t len) {
    if (!e || !src) return fail(WAX_VS_ERR_NULL, "NULL argument");
    std::unique_lock<std::shared_mutex> w(e->rw);
    // Reason strings follow MetalVectorEngine.deserialize (:716-815) / VectorSerializer.decodeVecSegment (:84-157).
    if (len < 36) return fail(WAX_VS_ERR_FORMAT, "Metal segment too small: %llu bytes", static_cast<unsigned long long>(len));
    const uint8_t magic[4] = {0x4D, 0x56, 0x32, 0x56};
    if (memcmp(src, magic, 4) != 0) return fail(WAX_VS_ERR_FORMAT, "Metal segment magic mismatch");
    uint16_t version; memcpy(&version, src + 4, 2);
    if (version != 1) return fail(WAX_VS_ERR_FORMAT, "Unsupported Metal segment version %u", version);
    if (src[6] != 2) return fail(WAX_VS_ERR_FORMAT, "Unsupported Metal segment encoding %u", src[6]);
    if (src[7] > 2 || src[7] != e->similarity)
        return fail(WAX_VS_ERR_FORMAT, "Metric mismatch: expected %u, got %u", e->similarity, src[7]);
    uint32_t dims; memcpy(&dims, src + 8, 4);
    if (dims != e->dims) return fail(WAX_VS_ERR_FORMAT, "Dimension mismatch: expected %u, got %u", e->dims, dims);
    uint64_t count, vbytes; memcpy(&count, src + 12, 8); memcpy(&vbytes, src + 20, 8);
    for (int i = 0; i < 8; ++i)
        if (src[28 + i] != 0) return fail(WAX_VS_ERR_FORMAT, "Metal segment reserved bytes must be zero");
    if (count > 0xFFFFFFFFull) return fail(WAX_VS_ERR_CAPACITY, "capacity exceeded: limit %llu, requested %llu", 0xFFFFFFFFull, static_cast<unsigned long long>(count));
    if (vbytes != count * static_cast<uint64_t>(dims) * 4ull) return fail(WAX_VS_ERR_FORMAT, "Vector data length mismatch");
    if (len < 36 + vbytes + 8) return fail(WAX_VS_ERR_FORMAT, "Metal segment missing frameId length");
    uint64_t ibytes; memcpy(&ibytes, src + 36 + vbytes, 8);
    if (ibytes != count * 8ull) return fail(WAX_VS_ERR_FORMAT, "FrameId data length mismatch");
    if (len != 36 + vbytes + 8 + ibytes)
        return fail(WAX_VS_ERR_FORMAT, "vec segment length mismatch: expected %llu, got %llu",
                    static_cast<unsigned long long>(36 + vbytes + 8 + ibytes), static_cast<unsigned long long>(len));
    DeviceGuard g(e->device);
    drain_device_path(e);
    int32_t rc = set_capacity(e, std::max<uint64_t>(count, 64));  // reservedCapacity = max(...) (:791-792)
    if (rc) return rc;
    if (vbytes && (rc = upload_bytes(e, e->d_corpus, src + 36, vbytes))) return rc;  // :794-799, pinned double-buffered H2D
    e->n_rows = count;
    e->ids.resize(count);
    if (count) memcpy(e->ids.data(), src + 36 + vbytes + 8, ibytes);  // :809-811
    e->ids_identity = false;
    e->map_valid = false;
    e->ids_sorted = true;
    for (uint64_t i = 1; i < count && e->ids_sorted; ++i) e->ids_sorted = e->ids[i] > e->ids[i - 1];
    e->d_ids_dirty = true;
    invalidate_row_caches(e, 0);
    return WAX_VS_OK;
}

// ---- instrumentation ------------------------------------------------------------------------------------------
int32_t wax_vs_debug_pool_stats(wax_vs_engine *e, uint64_t *allocations, uint64_t *reuses) {
    if (!e) return fail(WAX_VS_ERR_NULL, "engine is NULL");
    std::lock_guard<std::mutex> g(e->pool_mu);
    if (allocations) *allocations = e->pool_allocs;
    if (reuses) *reuses = e->pool_reuses;
    return WAX_VS_OK;
}

int32_t wax_vs_debug_fill_synthetic(wax_vs_engine *e, uint64_t seed, uint64_t first_row, uint64_t rows,
                                    uint64_t id_base, int32_t normalize) {
    if (!e) return fail(WAX_VS_ERR_NULL, "engine is NULL");
    if (rows > 0xFFFFFFFFull) return fail(WAX_VS_ERR_CAPACITY, "capacity exceeded: limit %llu, requested %llu", 0xFFFFFFFFull, static_cast<unsigned long long>(rows));
    std::unique_lock<std::shared_mutex> w(e->rw);
    DeviceGuard g(e->device);
    drain_device_path(e);
    e->n_rows = 0;
    int32_t rc = set_capacity(e, std::max<uint64_t>(rows, 64));
    if (rc) return rc;
    if (rows) {
        const unsigned blocks = static_cast<unsigned>((rows + 255) / 256);
        synth_fill_kernel<<<blocks, 256>>>(e->d_corpus, rows, e->dims, seed, first_row, normalize);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaDeviceSynchronize());
    }
    e->n_rows = rows;
    e->ids.clear(); e->ids.shrink_to_fit();
    e->ids_identity = true; e->id_base = id_base;
    e->map = IdMap(); e->map_valid = true; e->ids_sorted = true;
    e->d_ids_dirty = true;
    invalidate_row_caches(e, 0);
    return WAX_VS_OK;
}

int32_t wax_vs_debug_read_rows(wax_vs_engine *e, uint64_t first, uint64_t n, float *dst) {
    if (!e || !dst) return fail(WAX_VS_ERR_NULL, "NULL argument");
    std::shared_lock<std::shared_mutex> r(e->rw);
    if (first + n > e->n_rows) return fail(WAX_VS_ERR_ARGUMENT, "row range out of bounds");
    DeviceGuard g(e->device);
    if (n) CUDA_TRY(cudaMemcpy(dst, e->d_corpus + first * e->dims, n * e->dims * sizeof(float), cudaMemcpyDeviceToHost));
    return WAX_VS_OK;
}

int32_t wax_vs_debug_time_search(wax_vs_engine *e, uint32_t n_queries, int64_t top_k, uint64_t seed,
                                 uint32_t warmup, uint32_t iters, float *out_ms_total, uint64_t *out_launches) {
    if (!e || !out_ms_total) return fail(WAX_VS_ERR_NULL, "NULL argument");
    if (n_queries == 0) n_queries = 1;
    std::shared_lock<std::shared_mutex> r(e->rw);
    DeviceGuard g(e->device);
    SearchCtx *c = nullptr;
    int32_t rc = ctx_acquire(e, &c);
    if (rc) return rc;
    struct Rel { wax_vs_engine *e; SearchCtx *c; ~Rel() { ctx_release(e, c); } } rel{e, c};
    const uint32_t k_eff = clamp_topk(top_k);
    const size_t qfloats = static_cast<size_t>(n_queries) * e->dims;
    if ((rc = ensure_dev(&c->d_queries, &c->d_queries_cap, qfloats, "query buffer"))) return rc;
    if ((rc = ensure_dev(&c->d_out, &c->d_out_cap, static_cast<size_t>(n_queries) * k_eff, "result buffer"))) return rc;
    // n_queries distinct unit queries (generator stream `seed`); step i searches query i mod n_queries.
    synth_fill_kernel<<<(n_queries + 255) / 256, 256, 0, c->stream>>>(c->d_queries, n_queries, e->dims, seed, 0, 1);
    CUDA_TRY(cudaGetLastError());
    uint64_t launches = 0;
    // time_overlap: consecutive (independent) queries alternate over two streams so that one scan's tail overlaps
    // the next one's prologue -- what the sharded engine does with search_many_async.
    SearchCtx *c2 = nullptr;
    if (e->tune.time_overlap) {
        if ((rc = ctx_acquire(e, &c2))) return rc;
        if ((rc = ensure_dev(&c2->d_out, &c2->d_out_cap, static_cast<size_t>(n_queries) * k_eff, "result buffer"))) { ctx_release(e, c2); return rc; }
    }
    struct Rel2 { wax_vs_engine *e; SearchCtx *c; ~Rel2() { if (c) ctx_release(e, c); } } rel2{e, c2};
    for (uint32_t it = 0; it < warmup + iters; ++it) {
        if (it == warmup) {
            launches = 0;
            if (c2) { CUDA_TRY(cudaEventRecord(c2->ev0, c2->stream)); CUDA_TRY(cudaStreamWaitEvent(c->stream, c2->ev0, 0)); }
            CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
            if (c2) CUDA_TRY(cudaStreamWaitEvent(c2->stream, c->ev0, 0));
        }
        const uint32_t qi = it % n_queries;
        SearchCtx *cc = (c2 && (it & 1u)) ? c2 : c;
        if (c2 && it == 0) { CUDA_TRY(cudaEventRecord(c->ev1, c->stream)); CUDA_TRY(cudaStreamWaitEvent(c2->stream, c->ev1, 0)); }  // queries ready
        rc = enqueue_search(e, cc, c->d_queries + static_cast<size_t>(qi) * e->dims, k_eff, 0,
                            cc->d_out + static_cast<size_t>(qi) * k_eff, nullptr, cc->stream, &launches);
        if (rc) { cudaStreamSynchronize(c->stream); if (c2) cudaStreamSynchronize(c2->stream); return rc; }
    }
    if (c2) { CUDA_TRY(cudaEventRecord(c2->ev1, c2->stream)); CUDA_TRY(cudaStreamWaitEvent(c->stream, c2->ev1, 0)); }
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaEventElapsedTime(out_ms_total, c->ev0, c->ev1));
    if (out_launches) *out_launches = launches;
    return WAX_VS_OK;
}

int32_t wax_vs_debug_stream_read(wax_vs_engine *e, uint32_t iters, float *out_best_ms, uint64_t *out_bytes) {
    if (!e || !out_best_ms || !out_bytes) return fail(WAX_VS_ERR_NULL, "NULL argument");
    std::shared_lock<std::shared_mutex> r(e->rw);
    DeviceGuard g(e->device);
    SearchCtx *c = nullptr;
    int32_t rc = ctx_acquire(e, &c);
    if (rc) return rc;
    struct Rel { wax_vs_engine *e; SearchCtx *c; ~Rel() { ctx_release(e, c); } } rel{e, c};
    const uint64_t bytes = e->n_rows * e->dims * sizeof(float) / 16 * 16;
    *out_bytes = bytes;
    *out_best_ms = 0.0f;
    if (bytes == 0) return WAX_VS_OK;
    float best = 1e30f;
    for (uint32_t it = 0; it < iters + 2; ++it) {
        CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
        stream_read_kernel<<<e->sm_count * 4, 512, 0, c->stream>>>(reinterpret_cast<const uint4 *>(e->d_corpus),
                                                                     bytes / 16, c->d_ticket + 2);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (it >= 2 && ms < best) best = ms;
    }
    CUDA_TRY(cudaMemsetAsync(c->d_ticket, 0, 4 * sizeof(uint32_t), c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *out_best_ms = best;
    return WAX_VS_OK;
}

// Host <-> device transfer rates on this box, in GB/s, for `bytes` of pageable host memory (profiles/ingest_*.json):
//   [0] one-thread memcpy pageable -> pinned     [1] the staging copy with the engine's worker threads
//   [2] DMA pinned -> HBM                         [3] DMA HBM -> pinned
//   [4] upload pipeline pageable -> HBM           [5] download pipeline HBM -> pageable       [6] worker threads
int32_t wax_vs_debug_transfer_probe(wax_vs_engine *e, uint64_t bytes, float *out7) {
    if (!e || !out7) return fail(WAX_VS_ERR_NULL, "NULL argument");
    std::unique_lock<std::shared_mutex> w(e->rw);
    DeviceGuard g(e->device);
    bytes = std::max<uint64_t>(bytes, 1u << 20);
    int32_t rc = ingest_staging(e, bytes);
    if (rc) return rc;
    auto &ig = e->ing;
    if ((rc = ensure_dev(&ig.d_stage, &ig.d_stage_cap, static_cast<size_t>((bytes + 3) / 4), "probe buffer"))) return rc;
    std::vector<uint8_t> host(bytes, 1);
    auto secs = [](auto t0) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    uint8_t *probe_pin = nullptr;
    {
        std::lock_guard<std::mutex> pool(g_staging_mu);
        if ((rc = shared_staging())) return rc;
        probe_pin = g_staging[0];
    }
    const size_t chunk = std::min<size_t>(kStagingBytes, bytes);
    auto best = [&](auto fn) { double b = 1e30; for (int i = 0; i < 3; ++i) { auto t0 = std::chrono::steady_clock::now(); fn(); b = std::min(b, secs(t0)); } return b; };
    out7[0] = static_cast<float>(chunk / 1e9 / best([&] { memcpy(probe_pin, host.data(), chunk); }));
    out7[1] = static_cast<float>(chunk / 1e9 / best([&] { parallel_memcpy(probe_pin, host.data(), chunk, ig.threads); }));
    out7[2] = static_cast<float>(chunk / 1e9 / best([&] { cudaMemcpyAsync(ig.d_stage, probe_pin, chunk, cudaMemcpyHostToDevice, ig.stream); cudaStreamSynchronize(ig.stream); }));
    out7[3] = static_cast<float>(chunk / 1e9 / best([&] { cudaMemcpyAsync(probe_pin, ig.d_stage, chunk, cudaMemcpyDeviceToHost, ig.stream); cudaStreamSynchronize(ig.stream); }));
    out7[4] = static_cast<float>(bytes / 1e9 / best([&] { upload_bytes(e, ig.d_stage, host.data(), bytes); }));
    out7[5] = static_cast<float>(bytes / 1e9 / best([&] { download_bytes(e, host.data(), ig.d_stage, bytes); }));
    out7[6] = static_cast<float>(ig.threads);
    CUDA_TRY(cudaGetLastError());
    return WAX_VS_OK;
}

// Where a single fused search spends its time (TMA-staged kernels with the selection tail): averages over `iters`
// searches, microseconds: [0] kernel start -> last warp leaves the scan loop, [1] -> last CTA has selected its k,
// [2] -> the last CTA starts the grid stage, [3] -> result written (kernel end), [4] event-timed duration of the
// launch on the stream (launch overhead = [4] - [3]).
int32_t wax_vs_debug_phase_trace(wax_vs_engine *e, int64_t top_k, uint32_t iters, float *out5) {
    if (!e || !out5) return fail(WAX_VS_ERR_NULL, "NULL argument");
    std::unique_lock<std::shared_mutex> w(e->rw);
    DeviceGuard g(e->device);
    SearchCtx *c = nullptr;
    int32_t rc = ctx_acquire(e, &c);
    if (rc) return rc;
    struct Rel { wax_vs_engine *e; SearchCtx *c; ~Rel() { ctx_release(e, c); e->debug_trace = nullptr; } } rel{e, c};
    const uint32_t k_eff = clamp_topk(top_k);
    if ((rc = ensure_dev(&c->d_queries, &c->d_queries_cap, static_cast<size_t>(e->dims), "query buffer"))) return rc;
    if ((rc = ensure_dev(&c->d_out, &c->d_out_cap, static_cast<size_t>(k_eff), "result buffer"))) return rc;
    synth_fill_kernel<<<1, 256, 0, c->stream>>>(c->d_queries, 1, e->dims, 99, 0, 1);
    unsigned long long *d_trace = nullptr;
    CUDA_TRY(cudaMalloc(&d_trace, 8 * sizeof(unsigned long long)));
    double acc[5] = {0, 0, 0, 0, 0};
    uint64_t launches = 0;
    for (uint32_t it = 0; it < iters + 3; ++it) {
        const unsigned long long init[8] = {~0ull, 0, 0, 0, 0, 0, 0, 0};
        cudaMemcpyAsync(d_trace, init, sizeof init, cudaMemcpyHostToDevice, c->stream);
        cudaStreamSynchronize(c->stream);
        e->debug_trace = d_trace;
        cudaEventRecord(c->ev0, c->stream);
        rc = enqueue_search(e, c, c->d_queries, k_eff, 0, c->d_out, nullptr, c->stream, &launches);
        cudaEventRecord(c->ev1, c->stream);
        e->debug_trace = nullptr;
        if (rc) { cudaStreamSynchronize(c->stream); cudaFree(d_trace); return rc; }
        unsigned long long t[8];
        cudaMemcpyAsync(t, d_trace, sizeof t, cudaMemcpyDeviceToHost, c->stream);
        cudaStreamSynchronize(c->stream);
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        if (it < 3) continue;
        for (int i = 0; i < 4; ++i) acc[i] += (t[i + 1] > t[0] && t[0] != ~0ull) ? (t[i + 1] - t[0]) * 1e-3 : 0.0;
        acc[4] += ms * 1e3;
    }
    cudaFree(d_trace);
    for (int i = 0; i < 5; ++i) out5[i] = static_cast<float>(acc[i] / std::max(iters, 1u));
    CUDA_TRY(cudaGetLastError());
    return WAX_VS_OK;
}

int32_t wax_vs_debug_batch_stats(wax_vs_engine *e, uint64_t *tensor_queries, uint64_t *fallback_queries) {
    if (!e) return fail(WAX_VS_ERR_NULL, "engine is NULL");
    std::lock_guard<std::mutex> g(e->pool_mu);
    if (tensor_queries) *tensor_queries = e->batch_tensor_queries;
    if (fallback_queries) *fallback_queries = e->batch_fallback_queries;
    return WAX_VS_OK;
}

int32_t wax_vs_debug_counter(wax_vs_engine *e, const char *name, uint64_t *out) {
    if (!e || !name || !out) return fail(WAX_VS_ERR_NULL, "NULL argument");
    std::lock_guard<std::mutex> g(e->pool_mu);
    if (!strcmp(name, "batch_tensor_queries")) *out = e->batch_tensor_queries;
    else if (!strcmp(name, "batch_fallback_queries")) *out = e->batch_fallback_queries;
    else if (!strcmp(name, "batch_bf16_queries")) *out = e->batch_bf16_queries;
    else if (!strcmp(name, "batch_retry_queries")) *out = e->batch_retry_queries;               // TF32 filter level
    else if (!strcmp(name, "batch_filter_bf16_queries")) *out = e->batch_filter_bf16_queries;   // bf16-shadow filter level
    else if (!strcmp(name, "shadow_bytes")) *out = e->shadow_valid ? e->shadow_rows * e->dims * sizeof(__nv_bfloat16) : 0;   // live rows
    else if (!strcmp(name, "shadow_capacity_bytes")) *out = e->shadow_cap * sizeof(__nv_bfloat16);                            // HBM held
    else if (!strcmp(name, "shadow_unavailable")) *out = e->shadow_unavailable ? 1 : 0;   // bf16 shadow did not fit: TF32 level runs
    else if (!strcmp(name, "batch_tf32_queries")) *out = e->batch_tf32_queries;
    else if (!strcmp(name, "ingest_h2d_bytes")) *out = e->ingest_h2d_bytes;
    else if (!strcmp(name, "ingest_d2h_bytes")) *out = e->ingest_d2h_bytes;
    else if (!strcmp(name, "norms_rows")) *out = e->norms_rows;       // rows whose cached 1/|v| is valid
    else if (!strcmp(name, "shadow_rows")) *out = e->shadow_rows;     // rows whose bf16 shadow is valid
    else if (!strcmp(name, "batch_heap_bump")) *out = e->heap_bump;          // sizes above the model's nominee-heap choice (adaptive)
    else if (!strcmp(name, "batch_last_heap")) *out = e->last_heap;          // nominee heap entries of the last bf16 level-1 launch
    else if (!strcmp(name, "pool_allocs")) *out = e->pool_allocs;
    else if (!strcmp(name, "pool_reuses")) *out = e->pool_reuses;
    else return fail(WAX_VS_ERR_ARGUMENT, "unknown counter '%s'", name);
    return WAX_VS_OK;
}

int32_t wax_vs_debug_time_search_batch(wax_vs_engine *e, uint32_t n_queries, int64_t top_k, uint64_t seed,
                                       uint32_t warmup, uint32_t iters, float *out_ms_total,
                                       uint64_t *out_launches, uint32_t *out_unproven) {
    if (!e || !out_ms_total) return fail(WAX_VS_ERR_NULL, "NULL argument");
    if (n_queries == 0) n_queries = 1;
    std::shared_lock<std::shared_mutex> r(e->rw);
    DeviceGuard g(e->device);
    const uint32_t k_eff = static_cast<uint32_t>(std::min<uint64_t>(clamp_topk(top_k), std::max<uint64_t>(e->n_rows, 1)));
    if (!batch_tensor_eligible(e, n_queries, k_eff))
        return fail(WAX_VS_ERR_UNSUPPORTED, "batch of %u queries, k=%u, dims=%u is not eligible for the tensor path", n_queries, k_eff, e->dims);
    SearchCtx *c = nullptr;
    int32_t rc = ctx_acquire(e, &c);
    if (rc) return rc;
    struct Rel { wax_vs_engine *e; SearchCtx *c; ~Rel() { ctx_release(e, c); } } rel{e, c};
    const size_t qfloats = static_cast<size_t>(n_queries) * e->dims;
    if ((rc = ensure_dev(&c->d_queries, &c->d_queries_cap, qfloats, "query buffer"))) return rc;
    if ((rc = ensure_dev(&c->d_out, &c->d_out_cap, static_cast<size_t>(n_queries) * k_eff, "result buffer"))) return rc;
    if ((rc = ensure_dev(&c->d_ok, &c->ok_cap, static_cast<size_t>(n_queries), "proof flags"))) return rc;
    if ((rc = ensure_pinned(&c->h_ok, &c->h_ok_cap, static_cast<size_t>(n_queries), "proof flag staging"))) return rc;
    synth_fill_kernel<<<(n_queries + 255) / 256, 256, 0, c->stream>>>(c->d_queries, n_queries, e->dims, seed, 0, 1);
    CUDA_TRY(cudaGetLastError());
    if ((rc = ensure_norms(e, c->stream))) return rc;   // cached per corpus version: outside the timed region
    if (batch_bf16_wanted(e) && (rc = ensure_shadow(e, c->stream))) return rc;   // likewise
    uint64_t launches = 0;
    for (uint32_t it = 0; it < warmup + iters; ++it) {
        if (it == warmup) { launches = 0; CUDA_TRY(cudaEventRecord(c->ev0, c->stream)); }
        uint32_t used_heap = 0;
        rc = enqueue_batch_tensor(e, c, c->d_queries, n_queries, k_eff, 0, c->d_out, c->d_ok, nullptr, c->stream, &launches,
                                  true, nullptr, nullptr, nullptr, &used_heap);
        if (rc) { cudaStreamSynchronize(c->stream); return rc; }
        if (used_heap) { std::lock_guard<std::mutex> pg(e->pool_mu); e->last_heap = used_heap; }
    }
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->h_ok, c->d_ok, n_queries * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaEventElapsedTime(out_ms_total, c->ev0, c->ev1));
    if (out_launches) *out_launches = launches;
    if (out_unproven) {
        uint32_t bad = 0;
        for (uint32_t i = 0; i < n_queries; ++i) bad += c->h_ok[i] ? 0u : 1u;
        *out_unproven = bad;
    }
    return WAX_VS_OK;
}

int32_t wax_vs_debug_set_option(wax_vs_engine *e, const char *key, int64_t value) {
    if (!e || !key) return fail(WAX_VS_ERR_NULL, "NULL argument");
    std::unique_lock<std::shared_mutex> w(e->rw);
    const int v = static_cast<int>(value);
    if (!strcmp(key, "variant")) e->tune.variant = v;
    else if (!strcmp(key, "rows_per_step")) e->tune.rows_per_step = v;
    else if (!strcmp(key, "stages")) e->tune.stages = v;
    else if (!strcmp(key, "warps")) e->tune.warps = v;
    else if (!strcmp(key, "grid")) e->tune.grid = v;
    else if (!strcmp(key, "l2_hint")) e->tune.l2_hint = v;
    else if (!strcmp(key, "chunk_steps")) e->tune.chunk_steps = v;
    else if (!strcmp(key, "fused_k_max")) e->tune.fused_k_max = std::max(0, std::min(128, v));
    else if (!strcmp(key, "batch_tensor")) e->tune.batch_tensor = v;
    else if (!strcmp(key, "batch_min")) e->tune.batch_min = v;
    else if (!strcmp(key, "batch_noinsert")) e->tune.batch_noinsert = v;
    else if (!strcmp(key, "batch_heap")) e->tune.batch_heap = v;
    else if (!strcmp(key, "batch_large_k")) e->tune.batch_large_k = v;
    else if (!strcmp(key, "tma_max_dims")) e->tune.tma_max_dims = static_cast<uint32_t>(std::max(v, 0));
    else if (!strcmp(key, "batch_pair")) e->tune.batch_pair = v;
    else if (!strcmp(key, "batch_ts")) e->tune.batch_ts = v;
    else if (!strcmp(key, "batch_bf16")) { e->tune.batch_bf16 = v; e->shadow_unavailable = false; e->bf16_skip_batches = 0; }
    else if (!strcmp(key, "batch_ares")) e->tune.batch_ares = v;
    else if (!strcmp(key, "batch_rescore")) e->tune.batch_rescore = v;
    else if (!strcmp(key, "batch_retry")) e->tune.batch_retry = v;
    else if (!strcmp(key, "filter_cap")) e->tune.filter_cap = v;
    else if (!strcmp(key, "filter_bf16")) e->tune.filter_bf16 = v;
    else if (!strcmp(key, "single_shadow")) e->tune.single_shadow = v;
    else if (!strcmp(key, "shard_fused")) e->tune.shard_fused = v;
    else if (!strcmp(key, "tail_select")) e->tune.tail_select = v;
    else if (!strcmp(key, "inline_query")) e->tune.inline_query = v;
    else if (!strcmp(key, "host_delivery")) e->tune.host_delivery = v;
    else if (!strcmp(key, "shard_timeout_ms")) e->shard.timeout_ns = static_cast<unsigned long long>(std::max<int64_t>(value, 1)) * 1000000ull;
    else if (!strcmp(key, "time_overlap")) e->tune.time_overlap = v;
    else if (!strcmp(key, "ldg_ctas_per_sm")) e->tune.ldg_ctas_per_sm = std::max(1, v);
    else return fail(WAX_VS_ERR_ARGUMENT, "unknown option '%s'", key);
    return WAX_VS_OK;
}

}  // extern "C"
