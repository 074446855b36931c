// waxvs_scan.cuh -- the fused single-pass scan: query L2-norm + distance + top-k in ONE launch.
//
// Replaces, for the CUDA engine, the reference's two-dispatch pipeline
//   cosineDistanceKernelSIMD8/SIMD4  (Shaders/CosineDistance.metal:233-328, :152-229)  -> N*4 B distances
//   topKReduceDistances/Entries loop (Shaders/TopKReduction.metal:103-167; MetalVectorEngine.swift:511-575)
// and the host-side query normalisation (VectorSearchSession.swift:70-76).  The corpus is the only HBM
// stream: no distance array, no second pass.
//
// Data layout: corpus row-major [n_rows][dims] fp32 in HBM (as MetalVectorEngine.swift:134,345-349).
//
// Work split (HBM-bound, ~1 flop/byte -- deliberately NOT a GEMM):
//   * persistent grid, one CTA per SM; every warp owns a private ring of `stages` shared-memory tiles of
//     R rows and streams "steps" (R consecutive rows = R*dims*4 contiguous bytes) with 1-D TMA bulk copies
//     (cp.async.bulk -> SASS UBLKCP) completing on the warp's own mbarriers -- no block-wide barrier in
//     the main loop, (stages-1)*R*dims*4 bytes in flight per warp;
//   * a lane reads its 16-byte chunks (lane + 32c) of a row with conflict-free LDS.128, FMAs them against
//     the query chunks it keeps in registers into 4 accumulators per quantity (element i -> accumulator
//     i mod 128: the order oracle ACC_F32_TREE mirrors, so results are bit-exact against it);
//   * R rows are reduced together with a shuffle reduce-scatter (warp_reduce_scatter) and finished by the
//     lane that owns the row (IEEE sqrt/div, USearch zero-norm rules);
//   * each warp keeps a sorted k<=32 list in registers; a row is compared against the list's k-th key
//     (one 64-bit compare) and inserted by shuffles only when it wins (rare after warm-up);
//   * per-CTA merge in shared memory, per-grid merge by the last CTA to finish (atomic ticket): still the
//     same launch.
#pragma once
#include "waxvs_common.cuh"
#include "waxvs_shard.cuh"
#include "../../include/wax_vs_cuda.h"

namespace waxvs {

constexpr int kInlineQueryFloats = 512;

struct ScanParams {
    const float *corpus;        // [n_rows][dims]
    const float *query;         // [dims]
    uint32_t n_rows;
    uint32_t dims;
    uint32_t k;                 // entries to produce (<= 128 for the fused list kernels)
    uint32_t stages;            // ring depth per warp (TMA kernels)
    uint64_t *block_keys;       // [grid][32*E] scratch
    uint32_t *ticket;           // zero on entry, zero again on exit
    wax_vs_candidate *out;      // [k] results, best first
    uint32_t *dist_keys;        // emit mode: [n_rows] orderable distance keys (WAXVS_UKEY_NONE = dropped)
    const uint64_t *frame_ids;  // device ids or nullptr (then id = id_base + row)
    uint64_t id_base;
    uint64_t row_offset;        // added to the reported row (shard offset)
    uint32_t use_l2_hint;       // 1: evict-first policy on the corpus stream
    uint32_t chunk_steps;       // > 0: dynamic scheduling, warps claim chunks of this many steps from work_counter
    uint32_t *work_counter;     // zero on entry, zero again on exit (reset by the last CTA)
    const uint32_t *mask;       // optional row filter: bit r of mask[r / 32] set = row r may be returned (nullptr = all)
    // Host delivery (the synchronous single-query entry point): the last CTA also stores the result into mapped pinned
    // host memory and then raises a host-visible flag, so the caller needs neither a D2H copy nor a stream
    // synchronisation; and a short query travels in the kernel parameters instead of through an H2D copy.
    wax_vs_candidate *host_out;         // [k] mapped pinned, or nullptr
    unsigned long long *host_flag;      // mapped pinned: set to host_seq once host_out is complete
    unsigned long long host_seq;
    alignas(16) float query_inline[kInlineQueryFloats];   // used when query == nullptr (TMA-staged kernels, dims <= kInlineQueryFloats)
    // Tail of the launch: 0 = pairwise bitonic merges of sorted lists (warp lists -> block list -> last CTA merges the
    // grid's block lists), 1 = exact radix SELECTION (finish_topk_select): block-wide k-th-smallest over the keys, the
    // last CTA selects over grid x k keys staged in `tail_smem_bytes` of the (by then idle) ring.  Same result bits.
    uint32_t tail_select;
    uint32_t tail_smem_bytes;
    unsigned long long *trace;  // instrumentation (wax_vs_debug_phase_trace) or nullptr: [0] min kernel start, [1] max end of a
                                // warp's scan loop, [2] max end of a CTA's selection, [3] start and [4] end of the last CTA's
                                // grid stage -- %globaltimer nanoseconds
    ShardParams shard;          // shard.world > 0: `out` is this rank's local list and the last CTA goes on to exchange it
                                // with the other ranks over NVLink and to merge (waxvs_shard.cuh): still the same launch
};

__device__ __forceinline__ bool row_allowed(const ScanParams &p, uint32_t row) {
    return p.mask == nullptr || ((__ldg(p.mask + (row >> 5)) >> (row & 31u)) & 1u) != 0u;
}

__device__ __forceinline__ void write_candidate(const ScanParams &p, int slot, uint64_t key) {
    wax_vs_candidate c;
    if (key == WAXVS_KEY_NONE) {
        c.distance = 0.0f; c.valid = 0; c.row = 0; c.frame_id = 0;
    } else {
        const uint32_t row = static_cast<uint32_t>(key);
        c.distance = from_orderable_u32(static_cast<uint32_t>(key >> 32));
        c.valid = 1;
        c.row = p.row_offset + row;
        c.frame_id = p.frame_ids ? p.frame_ids[row] : p.id_base + row;
    }
    p.out[slot] = c;
    if (p.host_out) p.host_out[slot] = c;
}

// CTA merge + grid merge + output.  Called by every thread of the CTA after the scan loop.
// lists: shared [warps][E*32] u64;  block_keys: global [grid][E*32].
template <int E>
__device__ __forceinline__ void finish_topk(const ScanParams &p, WarpTopK<E> &tk, uint64_t *lists, int warp,
                                            int lane, int warps) {
    const int k = static_cast<int>(p.k);
    constexpr int W = E * 32;
    __shared__ uint32_t s_last;
    auto load_list = [&](const uint64_t *src, uint64_t (&dst)[E]) {
#pragma unroll
        for (int j = 0; j < E; ++j) dst[j] = src[j * 32 + lane];
    };
#pragma unroll
    for (int j = 0; j < E; ++j) lists[warp * W + j * 32 + lane] = tk.key[j];
    __syncthreads();
    if (warp == 0) {
#pragma unroll 1
        for (int w = 1; w < warps; ++w) {
            uint64_t other[E];
            load_list(lists + w * W, other);
            tk.merge_sorted(other, lane, k);
        }
#pragma unroll
        for (int j = 0; j < E; ++j) p.block_keys[static_cast<size_t>(blockIdx.x) * W + j * 32 + lane] = tk.key[j];
        __threadfence();
        __syncwarp();
        if (lane == 0) s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    tk.init();
    // Block lists come from L2 (~1 us each if loaded one by one): fetch four at a time, then merge.
    // (the merge loops are kept rolled: every inlined merge is ~150 instructions and this code runs once, so unrolled
    // copies only buy instruction-cache misses -- the E = 4 tail was 80 us of them at 10 K rows)
    constexpr int U = (E == 1) ? 4 : 2;
#pragma unroll 1
    for (uint32_t b0 = warp; b0 < gridDim.x; b0 += warps * U) {
        uint64_t other[U][E];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t b = b0 + u * warps;
#pragma unroll
            for (int j = 0; j < E; ++j)
                other[u][j] = (b < gridDim.x) ? ld_cg_u64(p.block_keys + static_cast<size_t>(b) * W + j * 32 + lane)
                                              : WAXVS_KEY_NONE;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) tk.merge_sorted(other[u], lane, k);
    }
    __syncthreads();  // everyone is done reading lists from the CTA merge
#pragma unroll
    for (int j = 0; j < E; ++j) lists[warp * W + j * 32 + lane] = tk.key[j];
    __syncthreads();
    if (warp == 0) {
#pragma unroll 1
        for (int w = 1; w < warps; ++w) {
            uint64_t other[E];
            load_list(lists + w * W, other);
            tk.merge_sorted(other, lane, k);
        }
#pragma unroll
        for (int j = 0; j < E; ++j)
            if (j * 32 + lane < k) write_candidate(p, j * 32 + lane, tk.key[j]);
        if (lane == 0) { *p.ticket = 0u; if (p.work_counter) *p.work_counter = 0u; }
        if (p.host_flag && !p.shard.world) {          // result complete in host memory: tell the waiting caller
            __threadfence_system();
            __syncwarp();
            if (lane == 0) st_release_sys_u64(p.host_flag, p.host_seq);
        }
    }
    // Row-sharded search: push the local list to every rank, wait for theirs, merge -- the block lists are done with,
    // their shared memory holds the distance keys of the merge (the host checks that world * k * 4 bytes fit).
    if (p.shard.world) shard_exchange_cta(p.shard, p.out, p.k, reinterpret_cast<uint32_t *>(lists));
}

// ------------------------------------------------------------------------------------------------------------
// Selection tail (round 2).  The merge tail above costs one ~150-instruction bitonic merge per pair of lists: 7 per
// CTA plus ~150 for the grid in the last CTA -- ~20 us at k = 10 and ~65 us at k = 72, which is most of a search over a
// real (<= 174 K-row) Wax index.  Selection does not care about order: a block-wide MSB-first radix select (8 bits a
// pass, it stops as soon as the bin holding the k-th key holds one key) finds the k-th smallest key exactly, the keys
// at or below it are the answer; only the final k are ranked (k x k compares) to come out sorted.
struct SelectScratch {
    uint32_t hist[256];
    uint64_t sel[128];          // the selected keys of the final stage
    uint32_t digit, k_rem, in_bin, n_sel;
    unsigned long long found;
};

// All threads of the CTA call.  for_each(f): f(key) for every key the calling thread owns (WAXVS_KEY_NONE = absent; it
// is the largest key, so it only matters when fewer than k real keys exist).  Returns the k-th smallest key of the CTA's
// keys (WAXVS_KEY_NONE when there are fewer than k real ones): the selection is {key <= result, key != NONE}.
template <typename ForEach>
__device__ __forceinline__ uint64_t block_select_kth(ForEach for_each, uint32_t k, SelectScratch *ss) {
    const uint32_t tid = threadIdx.x;
    uint64_t prefix = 0;
    uint32_t k_rem = k;
#pragma unroll 1
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (uint32_t i = tid; i < 256u; i += blockDim.x) ss->hist[i] = 0;
        __syncthreads();
        const uint64_t want = (shift == 56) ? 0ull : (prefix >> (shift + 8));
        for_each([&](uint64_t key) {
            if (shift == 56 || (key >> (shift + 8)) == want) atomicAdd(&ss->hist[(key >> shift) & 255u], 1u);
        });
        __syncthreads();
        if (tid < 32) {             // the digit whose cumulative count reaches k_rem: 8 bins per lane + a warp scan
            uint32_t c[8], sum = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { c[i] = ss->hist[tid * 8 + i]; sum += c[i]; }
            uint32_t incl = sum;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t up = __shfl_up_sync(WAXVS_FULL_MASK, incl, off);
                if (static_cast<int>(tid) >= off) incl += up;
            }
            const uint32_t excl = incl - sum;
            if (excl < k_rem && k_rem <= incl) {      // exactly one lane
                uint32_t run = excl;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (run < k_rem && k_rem <= run + c[i]) { ss->digit = tid * 8 + i; ss->k_rem = k_rem - run; ss->in_bin = c[i]; }
                    run += c[i];
                }
            }
        }
        __syncthreads();
        prefix |= static_cast<uint64_t>(ss->digit) << shift;
        k_rem = ss->k_rem;
        if (ss->in_bin == 1u && shift > 0) {          // one key carries this prefix: it is the k-th smallest -- fetch it
            const uint64_t top = prefix >> shift;
            for_each([&](uint64_t key) { if ((key >> shift) == top) ss->found = key; });
            __syncthreads();
            const uint64_t x = ss->found;
            __syncthreads();
            return x;
        }
    }
    return prefix;
}

// The selection tail: called by every thread of the CTA after the scan loop (the warps' register lists need not be
// merged, or even sorted, for this).  scratch = the dynamic shared memory (the idle ring), p.tail_smem_bytes of it.
template <int E>
__device__ __forceinline__ void finish_topk_select(const ScanParams &p, WarpTopK<E> &tk, unsigned char *scratch) {
    __shared__ SelectScratch ss;
    __shared__ uint32_t s_last2;
    const uint32_t tid = threadIdx.x, nthr = blockDim.x, k = p.k;
    // ---- stage A: this CTA's k smallest keys -> block_keys[blockIdx][0..k) (unordered, NONE-padded)
    auto own_keys = [&](auto f) {
#pragma unroll
        for (int j = 0; j < E; ++j) f(tk.key[j]);
    };
    const uint64_t xa = block_select_kth(own_keys, k, &ss);
    if (tid == 0) ss.n_sel = 0;
    __syncthreads();
    uint64_t *mine = p.block_keys + static_cast<size_t>(blockIdx.x) * k;
#pragma unroll
    for (int j = 0; j < E; ++j)
        if (tk.key[j] != WAXVS_KEY_NONE && tk.key[j] <= xa) mine[atomicAdd(&ss.n_sel, 1u)] = tk.key[j];
    __syncthreads();
    for (uint32_t i = ss.n_sel + tid; i < k; i += nthr) mine[i] = WAXVS_KEY_NONE;
    __threadfence();
    __syncthreads();
    if (p.trace && tid == 0) atomicMax(p.trace + 2, global_timer_ns());
    if (tid == 0) s_last2 = (atomicAdd(p.ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last2) return;
    __threadfence();
    if (p.trace && tid == 0) p.trace[3] = global_timer_ns();
    // ---- stage B (last CTA): the k smallest of the grid's gridDim.x * k keys, ranked, written out
    const uint32_t total = gridDim.x * k;
    const bool staged = static_cast<size_t>(total) * sizeof(uint64_t) <= p.tail_smem_bytes;
    uint64_t *sk = reinterpret_cast<uint64_t *>(scratch);
    if (staged) {
        for (uint32_t i = tid; i < total; i += nthr) sk[i] = ld_cg_u64(p.block_keys + i);
        __syncthreads();
    }
    auto grid_keys = [&](auto f) {
        if (staged) { for (uint32_t i = tid; i < total; i += nthr) f(sk[i]); }
        else { for (uint32_t i = tid; i < total; i += nthr) f(ld_cg_u64(p.block_keys + i)); }
    };
    const uint64_t xb = block_select_kth(grid_keys, k, &ss);
    if (tid == 0) ss.n_sel = 0;
    __syncthreads();
    grid_keys([&](uint64_t key) { if (key != WAXVS_KEY_NONE && key <= xb) ss.sel[atomicAdd(&ss.n_sel, 1u)] = key; });
    __syncthreads();
    const uint32_t n_sel = ss.n_sel;                  // == k, or every real key when there are fewer than k
    for (uint32_t i = tid; i < k; i += nthr) {
        if (i < n_sel) {
            const uint64_t key = ss.sel[i];
            uint32_t rank = 0;
            for (uint32_t j = 0; j < n_sel; ++j) rank += ss.sel[j] < key ? 1u : 0u;
            write_candidate(p, static_cast<int>(rank), key);
        } else {
            write_candidate(p, static_cast<int>(i), WAXVS_KEY_NONE);      // padding after the n_sel ranked entries
        }
    }
    if (p.host_flag && !p.shard.world) __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        *p.ticket = 0u;
        if (p.work_counter) *p.work_counter = 0u;
        if (p.host_flag && !p.shard.world) st_release_sys_u64(p.host_flag, p.host_seq);
        if (p.trace) p.trace[4] = global_timer_ns();
    }
    if (p.shard.world) shard_exchange_cta(p.shard, p.out, p.k, reinterpret_cast<uint32_t *>(scratch));
}

// ------------------------------------------------------------------------------------------------------------
// TMA-staged kernel.  C > 0: dims == 128*C, query chunks in registers, loops fully unrolled (the hot shapes
// 128..1024).  C == 0: any dims % 4 == 0 whose rows fit a stage (1536, 3072, 1000, ...): same algorithm with the
// chunk count at run time and the query chunks read from a shared-memory copy.
//   R      rows per step (power of two)
//   E      register-list slots per lane: fused top-k for k <= 32*E (E = 1: k <= 32, E = 4: k <= 128 -- the production
//          candidateLimit of 72, UnifiedSearch.swift:1195-1200, stays in the single launch)
//   EMIT   false: fused top-k;  true: write orderable distance keys for the large-k select path
template <int C, int R, int METRIC, int E, bool EMIT>
__global__ void __launch_bounds__(512, 1) scan_tma_kernel(const __grid_constant__ ScanParams p) {
    const int D4 = C > 0 ? 32 * C : static_cast<int>(p.dims / 4u);          // float4 per row
    const int CN = C > 0 ? C : (D4 + 31) / 32;                               // chunks per lane
    const uint32_t ROW_BYTES = C > 0 ? 512u * C : p.dims * 4u;
    const uint32_t STAGE_BYTES = ROW_BYTES * R;
    constexpr int LANES_PER_ROW = 32 / R;

    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    const uint32_t stages = p.stages;
    unsigned char *ring = smem + static_cast<size_t>(warp) * stages * STAGE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + static_cast<size_t>(warps) * stages * STAGE_BYTES) +
                     warp * stages;
    uint64_t *lists = reinterpret_cast<uint64_t *>(smem + static_cast<size_t>(warps) * stages * STAGE_BYTES) +
                      warps * stages;
    uint32_t *stage_step = reinterpret_cast<uint32_t *>(lists + warps * 32 * E) + warp * stages;  // step held by each stage
    float4 *qs = reinterpret_cast<float4 *>(reinterpret_cast<uint32_t *>(lists + warps * 32 * E) + warps * stages + 4);  // C == 0

    // ---- query chunks in registers (C > 0) or shared memory (C == 0) + fused |q|^2 ----
    float4 q[C > 0 ? C : 1];
    const float4 *q4 = reinterpret_cast<const float4 *>(p.query);
    const float4 *qp = reinterpret_cast<const float4 *>(p.query_inline);     // kernel-parameter copy (query == nullptr)
    auto load_q = [&](int i) -> float4 { return q4 ? __ldg(q4 + i) : qp[i]; };
    if (C > 0) {
#pragma unroll
        for (int c = 0; c < (C > 0 ? C : 1); ++c) q[c] = load_q(lane + 32 * c);
    } else {
        qs = reinterpret_cast<float4 *>((reinterpret_cast<uintptr_t>(qs) + 15) & ~uintptr_t(15));
        for (int i = threadIdx.x; i < CN * 32; i += blockDim.x) qs[i] = (i < D4) ? load_q(i) : make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
    }
    auto qchunk = [&](int c) -> float4 { return C > 0 ? q[C > 0 ? c : 0] : qs[lane + 32 * c]; };
    float a2 = 0.0f, sqrt_a2 = 0.0f;
    if (METRIC == kCosine) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int c = 0; c < CN; ++c) {
            const float4 qc = qchunk(c);
            s0 = __fmaf_rn(qc.x, qc.x, s0); s1 = __fmaf_rn(qc.y, qc.y, s1);
            s2 = __fmaf_rn(qc.z, qc.z, s2); s3 = __fmaf_rn(qc.w, qc.w, s3);
        }
        a2 = warp_butterfly_sum(__fadd_rn(__fadd_rn(s0, s1), __fadd_rn(s2, s3)));
        sqrt_a2 = __fsqrt_rn(a2);
    }

    if (p.trace && threadIdx.x == 0) atomicMin(p.trace + 0, global_timer_ns());
    const uint32_t total_warps = gridDim.x * warps;
    const uint32_t gwarp = blockIdx.x * warps + warp;
    const uint32_t n_steps = (p.n_rows + R - 1) / R;
    uint64_t policy = 0;
    if (p.use_l2_hint) policy = l2_policy_evict_first();

    auto issue = [&](uint32_t step, uint32_t s) {   // lane 0 only
        const uint32_t row0 = step * R;
        const uint32_t rows = min(static_cast<uint32_t>(R), p.n_rows - row0);
        const uint32_t bytes = rows * ROW_BYTES;
        stage_step[s] = step;                        // released to the warp by the mbarrier arrive below
        mbar_arrive_expect_tx(&bars[s], bytes);
        const float *src = p.corpus + static_cast<size_t>(row0) * p.dims;
        if (p.use_l2_hint) bulk_copy_g2s_hint(ring + s * STAGE_BYTES, src, bytes, &bars[s], policy);
        else bulk_copy_g2s(ring + s * STAGE_BYTES, src, bytes, &bars[s]);
    };

    // Step sequence of this warp.  Static: gwarp, gwarp + total_warps, ...  Dynamic (chunk_steps > 0, fused
    // top-k only): chunks of `chunk_steps` consecutive steps claimed from a global counter, the next claim
    // always in flight, so SMs that stream faster simply take more chunks (no tail imbalance).
    const bool dynamic = !EMIT && p.chunk_steps > 0 && p.work_counter != nullptr;
    const uint32_t chunk = p.chunk_steps;
    uint32_t cur = gwarp, cur_end = 0, claim_l0 = 0;
    if (dynamic) {
        if (lane == 0) { cur = atomicAdd(p.work_counter, chunk); claim_l0 = atomicAdd(p.work_counter, chunk); }
        cur = __shfl_sync(WAXVS_FULL_MASK, cur, 0);
        cur_end = min(cur + chunk, n_steps);
    }
    auto next_step = [&]() -> uint32_t {             // warp-uniform; >= n_steps when the warp is out of work
        if (!dynamic) { const uint32_t st = cur; cur = (cur < n_steps) ? cur + total_warps : cur; return st; }
        if (cur < cur_end) return cur++;
        if (cur >= n_steps) return n_steps;
        const uint32_t start = __shfl_sync(WAXVS_FULL_MASK, claim_l0, 0);
        if (lane == 0 && start < n_steps) claim_l0 = atomicAdd(p.work_counter, chunk);
        cur = start;
        cur_end = min(start + chunk, n_steps);
        if (cur >= n_steps) return n_steps;
        return cur++;
    };

    if (lane == 0) {
        for (uint32_t s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncwarp();
    uint32_t issued = 0, consumed = 0;
    for (uint32_t s = 0; s < stages; ++s) {
        const uint32_t step = next_step();
        if (step >= n_steps) break;
        if (lane == 0) issue(step, s);
        ++issued;
    }

    WarpTopK<E> tk;
    tk.init();
    const int k = static_cast<int>(p.k);

    uint32_t s = 0, parity = 0;
    while (consumed < issued) {
        mbar_wait_parity(&bars[s], parity);
        const uint32_t step = stage_step[s];
        const float4 *tile = reinterpret_cast<const float4 *>(ring + s * STAGE_BYTES);

        float sum0[R], sum1[R];
        if (C == 0) {
            // Generic rows (dims < 128 or not one of the unrolled multiples of 128), several rows per step: chunk-outer /
            // row-inner, so a query chunk is read from shared memory once for the R rows (the row-outer form read it per row:
            // twice the shared-memory traffic, 5.6 instead of 7.3 TB/s at 2560 dims).  Each row still sees its chunks in ascending
            // order on its own accumulators: the same operations in the same order as the row-outer loop below.
            float a[R][4], b[R][4];
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int j = 0; j < 4; ++j) { a[r][j] = 0.f; b[r][j] = 0.f; }
            // (a few chunks in flight per lane: with one or two rows per step the loads of consecutive chunks must overlap)
#pragma unroll (R >= 8 ? 1 : 8 / R)
            for (int c = 0; c < CN; ++c) {
                if (lane + 32 * c >= D4) break;
                const float4 qc = qchunk(c);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float4 v = tile[r * D4 + lane + 32 * c];
                    if (METRIC == kL2) {
                        const float dx = __fsub_rn(qc.x, v.x), dy = __fsub_rn(qc.y, v.y);
                        const float dz = __fsub_rn(qc.z, v.z), dw = __fsub_rn(qc.w, v.w);
                        a[r][0] = __fmaf_rn(dx, dx, a[r][0]); a[r][1] = __fmaf_rn(dy, dy, a[r][1]);
                        a[r][2] = __fmaf_rn(dz, dz, a[r][2]); a[r][3] = __fmaf_rn(dw, dw, a[r][3]);
                    } else {
                        a[r][0] = __fmaf_rn(qc.x, v.x, a[r][0]); a[r][1] = __fmaf_rn(qc.y, v.y, a[r][1]);
                        a[r][2] = __fmaf_rn(qc.z, v.z, a[r][2]); a[r][3] = __fmaf_rn(qc.w, v.w, a[r][3]);
                        if (METRIC == kCosine) {
                            b[r][0] = __fmaf_rn(v.x, v.x, b[r][0]); b[r][1] = __fmaf_rn(v.y, v.y, b[r][1]);
                            b[r][2] = __fmaf_rn(v.z, v.z, b[r][2]); b[r][3] = __fmaf_rn(v.w, v.w, b[r][3]);
                        }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                sum0[r] = __fadd_rn(__fadd_rn(a[r][0], a[r][1]), __fadd_rn(a[r][2], a[r][3]));
                sum1[r] = __fadd_rn(__fadd_rn(b[r][0], b[r][1]), __fadd_rn(b[r][2], b[r][3]));
            }
        } else
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float a0 = 0.f, a1 = 0.f, a2_ = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
#pragma unroll
            for (int c = 0; c < CN; ++c) {
                if (C == 0 && lane + 32 * c >= D4) break;   // ragged last chunk (dims % 128 != 0): this lane has no element
                const float4 v = tile[r * D4 + lane + 32 * c];
                const float4 qc = qchunk(c);
                if (METRIC == kL2) {
                    const float dx = __fsub_rn(qc.x, v.x), dy = __fsub_rn(qc.y, v.y);
                    const float dz = __fsub_rn(qc.z, v.z), dw = __fsub_rn(qc.w, v.w);
                    a0 = __fmaf_rn(dx, dx, a0); a1 = __fmaf_rn(dy, dy, a1);
                    a2_ = __fmaf_rn(dz, dz, a2_); a3 = __fmaf_rn(dw, dw, a3);
                } else {
                    a0 = __fmaf_rn(qc.x, v.x, a0); a1 = __fmaf_rn(qc.y, v.y, a1);
                    a2_ = __fmaf_rn(qc.z, v.z, a2_); a3 = __fmaf_rn(qc.w, v.w, a3);
                    if (METRIC == kCosine) {
                        b0 = __fmaf_rn(v.x, v.x, b0); b1 = __fmaf_rn(v.y, v.y, b1);
                        b2 = __fmaf_rn(v.z, v.z, b2); b3 = __fmaf_rn(v.w, v.w, b3);
                    }
                }
            }
            sum0[r] = __fadd_rn(__fadd_rn(a0, a1), __fadd_rn(a2_, a3));
            sum1[r] = __fadd_rn(__fadd_rn(b0, b1), __fadd_rn(b2, b3));
        }
        __syncwarp();  // every lane has consumed this stage (and read stage_step): safe to refill it
        {
            const uint32_t next = next_step();
            if (next < n_steps) {
                if (lane == 0) issue(next, s);
                ++issued;
            }
        }
        ++consumed;
        if (++s == stages) { s = 0; parity ^= 1u; }

        warp_reduce_scatter<R>(sum0, lane);
        if (METRIC == kCosine) warp_reduce_scatter<R>(sum1, lane);

        const uint32_t my_row = step * R + (lane / LANES_PER_ROW);
        float d;
        if (METRIC == kCosine) d = finish_cos(sum0[0], a2, sqrt_a2, sum1[0]);
        else if (METRIC == kDot) d = finish_dot(sum0[0]);
        else d = finish_l2(sum0[0]);
        const bool leader = (lane % LANES_PER_ROW) == 0;
        const bool ok = (my_row < p.n_rows) && finite_f32(d);

        if (EMIT) {
            if (leader && my_row < p.n_rows)
                p.dist_keys[my_row] = (ok && row_allowed(p, my_row)) ? orderable_u32(d) : WAXVS_UKEY_NONE;
        } else {
            const uint64_t key = ok ? make_key(d, my_row) : WAXVS_KEY_NONE;
            // the filter is consulted only for rows that would enter the list (rare once the list is warm)
            const bool cand = leader && key < tk.thresh && row_allowed(p, my_row);
            uint32_t m = __ballot_sync(WAXVS_FULL_MASK, cand);
            if (E == 1) {
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const uint64_t x = shfl_u64(key, src);
                    if (x < tk.thresh) tk.insert(x, lane, k);
                }
            } else if (m) {                             // batched insertion (WarpTopK::flush)
                if (tk.npend + __popc(m) > 32) tk.flush(lane, k);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    tk.park(shfl_u64(key, src), lane);
                }
            }
        }
    }

    if (p.trace && lane == 0) atomicMax(p.trace + 1, global_timer_ns());
    if (!EMIT) {
        if (E > 1) tk.flush(lane, k);
        if (p.tail_select) finish_topk_select<E>(p, tk, smem);
        else finish_topk<E>(p, tk, lists, warp, lane, warps);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Generic kernel: any dims (including dims % 4 != 0 and rows too large for a shared-memory tile).
// One warp per row, coalesced direct global loads (LDG.128 when dims % 4 == 0), same accumulation order.
template <int METRIC, int E, bool EMIT>
__global__ void __launch_bounds__(256, 4) scan_ldg_kernel(const __grid_constant__ ScanParams p) {
    __shared__ uint64_t lists[8 * 32 * E];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    const uint32_t dims = p.dims;
    const bool vec4 = (dims % 4u) == 0u;
    const uint32_t d4 = dims / 4u;

    float a2 = 0.0f, sqrt_a2 = 0.0f;
    if (METRIC == kCosine) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        for (uint32_t base = 4u * lane; base < dims; base += 128u) {
            const float x = __ldg(p.query + base);
            const float y = (base + 1 < dims) ? __ldg(p.query + base + 1) : 0.0f;
            const float z = (base + 2 < dims) ? __ldg(p.query + base + 2) : 0.0f;
            const float w = (base + 3 < dims) ? __ldg(p.query + base + 3) : 0.0f;
            s0 = __fmaf_rn(x, x, s0);
            if (base + 1 < dims) s1 = __fmaf_rn(y, y, s1);
            if (base + 2 < dims) s2 = __fmaf_rn(z, z, s2);
            if (base + 3 < dims) s3 = __fmaf_rn(w, w, s3);
        }
        a2 = warp_butterfly_sum(__fadd_rn(__fadd_rn(s0, s1), __fadd_rn(s2, s3)));
        sqrt_a2 = __fsqrt_rn(a2);
    }

    WarpTopK<E> tk;
    tk.init();
    const int k = static_cast<int>(p.k);
    const uint32_t total_warps = gridDim.x * warps;

    for (uint32_t row = blockIdx.x * warps + warp; row < p.n_rows; row += total_warps) {
        const float *v = p.corpus + static_cast<size_t>(row) * dims;
        float a0 = 0.f, a1 = 0.f, a2_ = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
        auto acc = [&](float qx, float vx, float &a, float &b) {
            if (METRIC == kL2) { const float dd = __fsub_rn(qx, vx); a = __fmaf_rn(dd, dd, a); }
            else {
                a = __fmaf_rn(qx, vx, a);
                if (METRIC == kCosine) b = __fmaf_rn(vx, vx, b);
            }
        };
        if (vec4) {
            const float4 *v4 = reinterpret_cast<const float4 *>(v);
            const float4 *q4 = reinterpret_cast<const float4 *>(p.query);
            // the row is read once: streaming loads (evict-first, no L1 residency) leave L1 to the query, which every row
            // re-reads; eight 16-byte loads in flight per lane
#pragma unroll 8
            for (uint32_t c = lane; c < d4; c += 32u) {
                const float4 x = __ldcs(v4 + c);
                const float4 y = __ldg(q4 + c);
                acc(y.x, x.x, a0, b0); acc(y.y, x.y, a1, b1); acc(y.z, x.z, a2_, b2); acc(y.w, x.w, a3, b3);
            }
        } else {
            for (uint32_t base = 4u * lane; base < dims; base += 128u) {
                acc(__ldg(p.query + base), __ldg(v + base), a0, b0);
                if (base + 1 < dims) acc(__ldg(p.query + base + 1), __ldg(v + base + 1), a1, b1);
                if (base + 2 < dims) acc(__ldg(p.query + base + 2), __ldg(v + base + 2), a2_, b2);
                if (base + 3 < dims) acc(__ldg(p.query + base + 3), __ldg(v + base + 3), a3, b3);
            }
        }
        const float s0 = warp_butterfly_sum(__fadd_rn(__fadd_rn(a0, a1), __fadd_rn(a2_, a3)));
        float d;
        if (METRIC == kCosine) {
            const float s1 = warp_butterfly_sum(__fadd_rn(__fadd_rn(b0, b1), __fadd_rn(b2, b3)));
            d = finish_cos(s0, a2, sqrt_a2, s1);
        } else if (METRIC == kDot) d = finish_dot(s0);
        else d = finish_l2(s0);
        const bool ok = finite_f32(d);
        if (EMIT) {
            if (lane == 0) p.dist_keys[row] = (ok && row_allowed(p, row)) ? orderable_u32(d) : WAXVS_UKEY_NONE;
        } else if (ok) {
            const uint64_t key = make_key(d, row);
            if (key < tk.thresh && row_allowed(p, row)) {
                if (E == 1) tk.insert(key, lane, k);
                else {
                    if (tk.npend == 32) tk.flush(lane, k);
                    tk.park(key, lane);
                }
            }
        }
    }
    if (!EMIT) {
        if (E > 1) tk.flush(lane, k);
        finish_topk<E>(p, tk, lists, warp, lane, warps);
    }
}

}  // namespace waxvs
