/*
 * wax_vs_cuda.h -- C-ABI of libwaxvs_cuda.so: the B200 (sm_100a) brute-force vector scan + top-k
 * that replaces WaxVectorSearch's Metal compute pipeline and CPU fallback behind the
 * `VectorSearchEngine` Swift protocol.
 *
 * Shape of the boundary.  Wax has no C-ABI for vector search; the seam is the Swift protocol
 *   Sources/WaxVectorSearch/VectorSearchEngine.swift:10-18
 * plus the concrete-type extras callers use (MetalVectorEngine.swift: init :153, isAvailable :144,
 * serialize :682, deserialize :716, addBatchStreaming :404).  Every entry point below is what a
 * `CUDAVectorEngine` Swift actor binds for one of those members; the cited line is the reference
 * member it replaces.  Conventions follow the repo's only FFI precedent, WaxCoreCompressionC
 * (Sources/WaxCoreCompressionC/include/wax_compression_shims.h:7-34): int32_t return code (0 = ok,
 * negative = distinct failure), caller-owned plain pointers + sizes, every pointer NULL-checked,
 * no ownership transfer.  INTEGRATION.md shows the Swift binding.
 *
 * Threading (mirrors the actor's AsyncReadWriteLock, MetalVectorEngine.swift:56-80): any number of
 * concurrent wax_vs_search* calls XOR one mutator (add/remove/reserve/deserialize/destroy).  The
 * library also enforces this internally with a reader/writer lock, so misuse blocks instead of racing.
 * search calls block the calling thread until results are in the caller's buffers.
 *
 * All arithmetic is IEEE fp32; ids are uint64; the corpus lives row-major in HBM.
 */
#ifndef WAX_VS_CUDA_H
#define WAX_VS_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wax_vs_engine wax_vs_engine; /* opaque: owns HBM, streams, scratch pool */

/* Return codes.  The Swift side maps them onto the WaxError cases the Metal engine throws. */
enum {
    WAX_VS_OK = 0,
    WAX_VS_ERR_NULL = -1,          /* NULL argument                         -> WaxError.invalidToc       */
    WAX_VS_ERR_DIMENSION = -2,     /* vector length != dimensions           -> WaxError.encodingError
                                      (MetalVectorEngine.swift:830-833, :360-370)                       */
    WAX_VS_ERR_CAPACITY = -3,      /* dims > 1 000 000 / rows > UInt32.max  -> WaxError.capacityExceeded
                                      (MetalVectorEngine.swift:157-162, :858-860)                       */
    WAX_VS_ERR_CUDA = -4,          /* device / allocation / launch failure  -> WaxError.invalidToc(reason:) */
    WAX_VS_ERR_FORMAT = -5,        /* malformed MV2V blob                   -> WaxError.invalidToc(reason:)
                                      (MetalVectorEngine.swift:718-808)                                 */
    WAX_VS_ERR_ARGUMENT = -6,      /* dims == 0, unknown similarity, ...    -> WaxError.invalidToc       */
    WAX_VS_ERR_BUFFER = -7,        /* caller buffer too small                                           */
    WAX_VS_ERR_UNSUPPORTED = -8    /* feature not available in this build                                */
};

/* VecSimilarity raw values (Sources/WaxCore/FileFormat/MV2SEnums.swift:34-38). */
enum { WAX_VS_COSINE = 0, WAX_VS_DOT = 1, WAX_VS_L2 = 2 };

#define WAX_VS_MAX_RESULTS 10000      /* MetalVectorEngine.swift:18 */
#define WAX_VS_MAX_DIMENSIONS 1000000 /* Sources/WaxCore/Constants.swift:51 */

/* One candidate of the device-side result list (sharded search: what a rank contributes to the
   NCCL all-gather).  24 bytes, naturally aligned. */
typedef struct wax_vs_candidate {
    float distance;     /* USearch-convention distance (ascending = better)                        */
    uint32_t valid;     /* 1 = real candidate, 0 = padding (fewer than k finite candidates)        */
    uint64_t row;       /* global row = row_offset + local row: the cross-shard tie-break key      */
    uint64_t frame_id;
} wax_vs_candidate;

/* ---- availability / lifetime ---------------------------------------------------------------- */

/* MetalVectorEngine.isAvailable (MetalVectorEngine.swift:144-146): available <=> rc == 0 && *out > 0. */
int32_t wax_vs_device_count(int32_t *out);

/* MetalVectorEngine.init(metric:dimensions:) (MetalVectorEngine.swift:153-274).  Unlike the Metal
   engine (cosine only, :163-165) all three metrics are supported, with USearchVectorEngine's
   semantics (USearchVectorEngine.swift:44-67; VectorMetric.swift:21-30).  `devices`/`n_devices`: the
   CUDA ordinals to use; NULL/0 = current device.  This build places one engine on one device
   (n_devices must be <= 1); row-sharding across GPUs is one engine per rank + wax_vs_search_device +
   one all-gather (wax_b200/sharded.py). */
int32_t wax_vs_create(uint32_t dimensions, uint8_t similarity, const int32_t *devices, int32_t n_devices,
                      wax_vs_engine **out);
void wax_vs_destroy(wax_vs_engine *engine);

/* `dimensions` property of the protocol (VectorSearchEngine.swift:11). */
int32_t wax_vs_dimensions(const wax_vs_engine *engine, uint32_t *out);
int32_t wax_vs_similarity(const wax_vs_engine *engine, uint8_t *out);
/* vectorCount (MetalVectorEngine.swift:50). */
int32_t wax_vs_count(wax_vs_engine *engine, uint64_t *out);

/* ---- corpus mutation -------------------------------------------------------------------------- */

/* reserveIfNeeded (MetalVectorEngine.swift:857-871): make room for `rows` rows in HBM. */
int32_t wax_vs_reserve(wax_vs_engine *engine, uint64_t rows);

/* add(frameId:vector:) (MetalVectorEngine.swift:330-357): upsert one row. `vector_len` must equal
   dimensions (validate, :830-833). */
int32_t wax_vs_add(wax_vs_engine *engine, uint64_t frame_id, const float *vector, uint32_t vector_len);

/* Bulk transfers (add_batch, deserialize, serialize) move caller memory through two pinned staging buffers with the
   host-side copy of one chunk overlapping the DMA of the other, so PAGEABLE caller buffers still reach PCIe speed;
   caller memory that is already pinned is handed to the DMA engine directly. */
/* addBatch(frameIds:vectors:) (MetalVectorEngine.swift:359-402): upsert n rows, `rows` is n x dims
   row-major (the Swift side flattens [[Float]]).  An id already present is overwritten in place
   (:385-389), a new id is appended at row N (:390-397); later duplicates inside one batch overwrite
   earlier ones, as the reference's sequential loop does.  n == 0 is a no-op (:360). */
int32_t wax_vs_add_batch(wax_vs_engine *engine, const uint64_t *frame_ids, const float *rows, uint64_t n,
                         uint32_t vector_len);

/* remove(frameId:) (MetalVectorEngine.swift:423-444): unknown id / empty engine = no-op rc 0 (:425-426);
   known id: row deleted, later rows keep their relative order (:431-441). */
int32_t wax_vs_remove(wax_vs_engine *engine, uint64_t frame_id);

/* remove(frameId:) for n frames in ONE pass (the reference memmoves the whole tail once per id,
   MetalVectorEngine.swift:431-441): same result as n calls of wax_vs_remove in any order -- unknown / repeated ids
   are ignored, surviving rows keep their relative order -- with one compaction of the matrix in HBM, one compaction
   of the id array and one id->row hash rebuild.  *out_removed (optional) = rows actually deleted. */
int32_t wax_vs_remove_batch(wax_vs_engine *engine, const uint64_t *frame_ids, uint64_t n, uint64_t *out_removed);

/* ---- search ------------------------------------------------------------------------------------- */

/* search(vector:topK:) (VectorSearchEngine.swift:13; MetalVectorEngine.swift:446-627;
   USearchVectorEngine.swift:201-216).
     - empty engine: *out_n = 0, rc 0 (:448)
     - query_len != dimensions: WAX_VS_ERR_DIMENSION before any work (:449)
     - top_k clamped to [1, 10000] (:450, :842-846); returns min(k, N) rows minus non-finite (:597)
     - best first: ascending distance, ties by ascending row (the reference leaves ties unspecified)
     - out_scores[i] = VectorMetric.score(fromDistance:) (VectorMetric.swift:32-43):
         cosine 1 - d with d = 1 - q.v/(|q||v|), ALWAYS divided by the in-kernel |q|;
         dot -(1 - q.v);  l2 -sum (q-v)^2
   out_ids / out_scores need room for out_cap entries, out_cap >= min(clamp(top_k), N) else
   WAX_VS_ERR_BUFFER. */
int32_t wax_vs_search(wax_vs_engine *engine, const float *query, uint32_t query_len, int64_t top_k,
                      uint64_t *out_ids, float *out_scores, uint32_t out_cap, uint32_t *out_n);

/* n_queries independent searches over one pass of the corpus (batched form of the above; results of
   query i start at out_ids[i*out_stride], count out_n[i]).  out_stride >= min(clamp(top_k), N). */
int32_t wax_vs_search_batch(wax_vs_engine *engine, const float *queries, uint32_t n_queries,
                            uint32_t query_len, int64_t top_k, uint64_t *out_ids, float *out_scores,
                            uint32_t out_stride, uint32_t *out_n);

/* Filtered search -- SURVEY.md section 8(f) rank 4, an API EXTENSION over the reference: Wax filters frames
   after the engine call and over-fetches 3 x topK to compensate (UnifiedSearch.swift:58, :371-442, :1195-1200,
   :1241-1258).  Here the filter is applied below the top-k, so exactly min(clamp(top_k), #allowed) best
   allowed rows come back.  mode 0: only rows whose frameId is in frame_ids[] may be returned (allow-list);
   mode 1: rows whose frameId is in frame_ids[] are excluded (deny-list, e.g. deleted / superseded frames).
   Unknown ids are ignored.  Same ordering, scoring and error behaviour as wax_vs_search. */
int32_t wax_vs_search_filtered(wax_vs_engine *engine, const float *query, uint32_t query_len, int64_t top_k,
                               const uint64_t *frame_ids, uint64_t n_ids, int32_t mode, uint64_t *out_ids,
                               float *out_scores, uint32_t out_cap, uint32_t *out_n);

/* The batched form: ONE filter, n_queries queries, one pass over the corpus (results of query i start at
   out_ids[i*out_stride], count out_n[i]; out_stride >= min(clamp(top_k), #allowed)).  Allow-lists of <= 16 384 rows
   score only the listed rows; otherwise the row filter rides below the top-k of the tensor-core levels (nominations,
   filter level and the exact fall-back consult the same bitset) or of the fused scan.  Results are identical to
   n_queries calls of wax_vs_search_filtered. */
int32_t wax_vs_search_batch_filtered(wax_vs_engine *engine, const float *queries, uint32_t n_queries,
                                     uint32_t query_len, int64_t top_k, const uint64_t *frame_ids, uint64_t n_ids,
                                     int32_t mode, uint64_t *out_ids, float *out_scores, uint32_t out_stride,
                                     uint32_t *out_n);

/* Device-resident form used by the row-sharded engine: `d_queries` (n_queries x dims) and
   `d_candidates` (n_queries x k_eff entries, k_eff = min(clamp(top_k), 10000) -- NOT clipped to N, padding
   has valid = 0) are DEVICE pointers on the engine's device; the work is enqueued on `cuda_stream`
   (a cudaStream_t; NULL = legacy default stream) and the call returns without synchronising.
   candidate.row = row_offset + local row.
   Ordering against mutators: the library remembers that device-path work was enqueued and every mutator
   (add / remove / reserve / deserialize / fill) drains the DEVICE (cudaDeviceSynchronize) under its write lock
   before it touches the corpus, so an in-flight scan never reads rows that are being moved. */
int32_t wax_vs_search_device(wax_vs_engine *engine, const float *d_queries, uint32_t n_queries,
                             int64_t top_k, uint64_t row_offset, wax_vs_candidate *d_candidates,
                             void *cuda_stream);

/* Batched device-resident form (the row-sharded engine's search_batch): same pointers and candidate layout as
   wax_vs_search_device, but the queries go through the batched tensor-core levels (bf16-shadow nominations ->
   TF32 retry -> exact scan; results identical to n_queries single-query calls) whenever the batch is eligible
   (wax_vs_search_batch's rules).  Those levels read their proof flags back, so this call MAY synchronise
   `cuda_stream` before it returns; d_candidates is complete in `cuda_stream` order. */
int32_t wax_vs_search_batch_device(wax_vs_engine *engine, const float *d_queries, uint32_t n_queries,
                                   int64_t top_k, uint64_t row_offset, wax_vs_candidate *d_candidates,
                                   void *cuda_stream);

/* ---- row-sharded search across the GPUs of one node (no reference counterpart) -------------
   One engine per GPU holds a contiguous row range of the corpus; the query is replicated; every rank scans its shard
   and the per-shard top-k lists are exchanged and merged under the total order (distance, GLOBAL row).  The exchange is
   fused into the scan launch: the kernel's last CTA writes its k candidates straight into every rank's mailbox over
   NVLink / NVSwitch peer memory, raises a flag, waits for the others' flags and merges world x k candidates -- no
   collective launch, no D2H copy, no host merge (wax_b200/csrc/waxvs_shard.cuh).  Every rank gets the same result.

   Setup: each rank calls wax_vs_shard_open (allocates its mailbox, returns a WAX_VS_SHARD_HANDLE_BYTES blob), the
   blobs are exchanged by any out-of-band means (the Python mirror uses one torch.distributed all-gather; a single
   process driving several GPUs just passes them along), then every rank calls wax_vs_shard_connect with all `world`
   blobs in rank order (CUDA IPC between processes, peer access inside one process).
   Searches are COLLECTIVE: every rank must issue the same searches in the same order (same query, same top_k).
   top_k is clamped to [1, 10000] as usual but must not exceed WAX_VS_SHARD_MAX_K (the fused top-k range);
   larger k -> WAX_VS_ERR_UNSUPPORTED (gather wax_vs_search_device candidates instead).  A rank whose peers never
   arrive gets WAX_VS_ERR_CUDA after the exchange timeout (20 s; option "shard_timeout_ms") instead of hanging. */
#define WAX_VS_SHARD_HANDLE_BYTES 128
#define WAX_VS_SHARD_MAX_RANKS 16
#define WAX_VS_SHARD_MAX_K 128
int32_t wax_vs_shard_open(wax_vs_engine *engine, int32_t rank, int32_t world, uint64_t row_offset,
                          uint8_t *out_handle /* WAX_VS_SHARD_HANDLE_BYTES */);
int32_t wax_vs_shard_connect(wax_vs_engine *engine, const uint8_t *handles /* n_handles x HANDLE_BYTES, rank order */,
                             int32_t n_handles);
/* Leave the group: unmaps the peers' mailboxes.  This rank's own mailbox stays allocated until wax_vs_destroy or the
   next wax_vs_shard_open, because other processes may still have it mapped -- close on every rank, synchronise the
   ranks (a barrier), then destroy. */
int32_t wax_vs_shard_close(wax_vs_engine *engine);
/* search(vector:topK:) over the whole sharded corpus; host query in, host ids / scores out (best first, as
   wax_vs_search); blocks until the merged result is in the caller's buffers.  out_cap >= clamp(top_k). */
int32_t wax_vs_shard_search(wax_vs_engine *engine, const float *query, uint32_t query_len, int64_t top_k,
                            uint64_t *out_ids, float *out_scores, uint32_t out_cap, uint32_t *out_n);
/* Device-side merge for the sharded search_batch: d_gathered = [world][n_queries][k] candidates exactly as an
   all-gather of the ranks' wax_vs_search_batch_device outputs leaves them (rank-major; every per-query list sorted, padding
   valid = 0 last); d_out = [n_queries][k_out] (k_out <= k), the k_out best of each query under (distance, GLOBAL row)
   -- ranks must be ordered by ascending row ranges.  Enqueued on cuda_stream, no synchronisation. */
int32_t wax_vs_merge_candidates_device(wax_vs_engine *engine, const wax_vs_candidate *d_gathered, uint32_t world,
                                       uint32_t n_queries, uint32_t k, uint32_t k_out, wax_vs_candidate *d_out,
                                       void *cuda_stream);
/* wax_vs_search_filtered over the whole sharded corpus: every rank passes the SAME frame_ids / mode; a rank resolves
   the ids its own shard holds (the rest are unknown to it and ignored), its fused scan applies the row filter below
   the top-k and the in-kernel exchange merges the ranks' lists.  Still one launch per query per rank. */
int32_t wax_vs_shard_search_filtered(wax_vs_engine *engine, const float *query, uint32_t query_len, int64_t top_k,
                                     const uint64_t *frame_ids, uint64_t n_ids, int32_t mode, uint64_t *out_ids,
                                     float *out_scores, uint32_t out_cap, uint32_t *out_n);
/* Device-resident form: d_query (dims floats) and d_candidates (clamp(top_k) merged entries, padding valid = 0) are
   device pointers; enqueued on `cuda_stream`, returns without synchronising. */
int32_t wax_vs_shard_search_device(wax_vs_engine *engine, const float *d_query, int64_t top_k,
                                   wax_vs_candidate *d_candidates, void *cuda_stream);

/* ---- persistence: "MV2V" v1 encoding = 2, byte-identical to MetalVectorEngine.serialize ---------- */

/* serialize() (MetalVectorEngine.swift:682-714). */
int32_t wax_vs_serialized_length(wax_vs_engine *engine, uint64_t *out);
int32_t wax_vs_serialize(wax_vs_engine *engine, uint8_t *dst, uint64_t cap, uint64_t *out_len);
/* deserialize(_:) (MetalVectorEngine.swift:716-815; VectorSerializer.swift:84-157): replaces the
   engine's contents.  Any violation -> WAX_VS_ERR_FORMAT with the reference's reason string in
   wax_vs_last_error(). */
int32_t wax_vs_deserialize(wax_vs_engine *engine, const uint8_t *src, uint64_t len);

/* Thread-local, NUL-terminated reason for the last non-zero return on this thread (the `reason:` of
   the WaxError the Swift side throws). Never NULL. */
const char *wax_vs_last_error(void);

/* ---- instrumentation (not part of the reference surface) -------------------------------------------- */

/* debugBufferPoolStats (MetalVectorEngine.swift:119-121; MetalVectorEnginePoolTests.swift:7-20):
   how many per-search scratch contexts were ever allocated / reused. */
int32_t wax_vs_debug_pool_stats(wax_vs_engine *engine, uint64_t *allocations, uint64_t *reuses);

/* Replace the contents with `rows` synthetic rows generated ON DEVICE (100 M x 384 does not fit a host):
   row r holds generator row (first_row + r) of stream `seed` (bit-identical to
   oracle wax_oracle_synth_row), frameId = id_base + r. */
int32_t wax_vs_debug_fill_synthetic(wax_vs_engine *engine, uint64_t seed, uint64_t first_row,
                                    uint64_t rows, uint64_t id_base, int32_t normalize);
/* Copy rows [first, first+n) back to the host (tests). */
int32_t wax_vs_debug_read_rows(wax_vs_engine *engine, uint64_t first, uint64_t n, float *dst);

/* Kernel-only timing with everything resident in HBM: generates `n_queries` distinct unit queries on device
   (generator stream `seed`), runs `warmup` + `iters` single-query searches back to back on one stream (step i
   uses query i mod n_queries), brackets the `iters` with CUDA events on that stream and returns the total
   milliseconds plus the number of kernel launches inside the bracket. */
int32_t wax_vs_debug_time_search(wax_vs_engine *engine, uint32_t n_queries, int64_t top_k, uint64_t seed,
                                 uint32_t warmup, uint32_t iters, float *out_ms_total,
                                 uint64_t *out_launches);

/* wax_vs_debug_time_search for the sharded path: warmup + iters collective searches strictly one at a time on one
   stream, CUDA events around the `iters` (every rank makes the same call; queries are generated on device from
   `seed`, identical on every rank). */
int32_t wax_vs_debug_time_shard_search(wax_vs_engine *engine, uint32_t n_queries, int64_t top_k, uint64_t seed,
                                       uint32_t warmup, uint32_t iters, float *out_ms_total, uint64_t *out_launches);

/* Host <-> device transfer rates on this box (GB/s) for `bytes` of pageable host memory: out7 = {one-thread staging
   memcpy, staging memcpy with the worker threads, DMA pinned->HBM, DMA HBM->pinned, upload pipeline pageable->HBM,
   download pipeline HBM->pageable, worker threads}.  Explains what bounds wax_vs_add_batch / wax_vs_serialize. */
int32_t wax_vs_debug_transfer_probe(wax_vs_engine *engine, uint64_t bytes, float *out7);

/* Where one fused search spends its time, from %globaltimer stamps inside the kernel (averages over `iters` searches,
   microseconds): out5 = {kernel start -> last warp leaves the scan loop, -> last CTA has selected its k, -> the last CTA
   starts the grid stage, -> result written, CUDA-event duration of the launch on its stream}. */
int32_t wax_vs_debug_phase_trace(wax_vs_engine *engine, int64_t top_k, uint32_t iters, float *out5);

/* Batched-path instrumentation: how many queries were answered by the tensor-core nomination path with a
   completed exactness proof, and how many had to be re-run on the exact single-query path. */
int32_t wax_vs_debug_batch_stats(wax_vs_engine *engine, uint64_t *tensor_queries, uint64_t *fallback_queries);

/* Named instrumentation counters: "batch_tensor_queries", "batch_fallback_queries", "batch_bf16_queries" (queries
   nominated from the bf16 shadow), "batch_retry_queries" (bf16-unproven queries retried on the TF32 nominations),
   "shadow_bytes" (HBM held by the bf16 shadow), "shadow_unavailable" (1 = the shadow did not fit in HBM, batches
   nominate in TF32 at about half the rate), "batch_tf32_queries", "pool_allocs", "pool_reuses". */
int32_t wax_vs_debug_counter(wax_vs_engine *engine, const char *name, uint64_t *out);

/* Device-only timing of the batched path (n_queries synthetic unit queries per step, everything resident):
   total milliseconds of `iters` steps (CUDA events on the launching stream), kernel launches in the bracket and
   the number of queries of the last step whose proof did not complete (they would be re-run exactly). */
int32_t wax_vs_debug_time_search_batch(wax_vs_engine *engine, uint32_t n_queries, int64_t top_k, uint64_t seed,
                                       uint32_t warmup, uint32_t iters, float *out_ms_total,
                                       uint64_t *out_launches, uint32_t *out_unproven);

/* Streaming-read ceiling on the same box: a plain coalesced LDG.128 read of the live corpus bytes, best of
   `iters` (milliseconds, and the bytes read).  Context for the roofline fraction. */
int32_t wax_vs_debug_stream_read(wax_vs_engine *engine, uint32_t iters, float *out_best_ms, uint64_t *out_bytes);

/* Tuning knobs for experiments ("variant", "ctas_per_sm", ...).  Unknown key -> WAX_VS_ERR_ARGUMENT. */
int32_t wax_vs_debug_set_option(wax_vs_engine *engine, const char *key, int64_t value);

/* Library build info: "waxvs_cuda <version> sm_100a ...". */
const char *wax_vs_version(void);

#ifdef __cplusplus
}
#endif
#endif /* WAX_VS_CUDA_H */
