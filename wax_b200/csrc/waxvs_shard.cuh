// waxvs_shard.cuh -- the row-sharded search's ONE exchange, fused into the scan launch.
//
// the corpus shards row-wise across the GPUs of a node, every rank scans its shard for the
// replicated query and the per-shard top-k lists (k x 24 B per rank) are exchanged and merged under the same total
// order (distance ascending, GLOBAL row ascending).  The reference has no distributed code; round 1 did this with
// torch.distributed.all_gather_into_tensor + D2H + a numpy merge, which cost ~0.13 ms per query at 8 GPUs -- half a
// 1.25 M-row scan.  Here the exchange is part of the scan kernel itself:
//
//   * every rank owns a MAILBOX in its HBM (ShardMailbox); peers map it (CUDA IPC between the one-process-per-GPU
//     ranks, plain peer access inside one process) and WRITE into it over NVLink / NVSwitch -- nobody ever reads
//     remote memory (a remote read is a round trip, a remote write is posted);
//   * the last CTA of the scan (the one that already merges the grid's block lists) pushes its k candidates into
//     slot [seq % depth][rank] of EVERY rank's mailbox, fences at system scope, raises flag[slot][rank] = seq on every
//     rank, then spins on its OWN mailbox until all `world` flags carry seq;
//   * it merges the world x k candidates (every list is sorted: a candidate's final position is its index plus, for
//     every other rank, a binary search of its distance in that rank's list -- upper bound for lower ranks, lower
//     bound for higher ranks, which is exactly the (distance, global row) order because shards are contiguous and
//     ascending by rank), writes the k best to the result buffer and, for the host entry point, straight into mapped
//     pinned host memory followed by a host-visible flag: no collective launch, no D2H copy, no host merge;
//   * slot reuse is guarded by acknowledgements (acks[r] = last seq rank r has finished reading from its own mailbox),
//     so any number of queries may be in flight on any streams.
//
// Every rank computes the same merge from the same bytes, so all ranks return identical results (what an all-gather
// followed by a merge on every rank gives).  Spins carry a timeout (a rank that never arrives must not hang the GPU).
#pragma once
#include "waxvs_common.cuh"
#include "../../include/wax_vs_cuda.h"

namespace waxvs {

constexpr int kShardMaxRanks = 16;
constexpr int kShardDepth = 8;       // queries whose candidates a mailbox can hold at once
constexpr int kShardKCap = 128;      // = the fused top-k range

struct ShardMailbox {
    unsigned long long flags[kShardDepth][kShardMaxRanks];   // flags[s][r] = seq: rank r's candidates for seq are in cands[s][r]
    unsigned long long acks[kShardMaxRanks];                  // acks[r] = last seq rank r finished reading from ITS OWN mailbox
    wax_vs_candidate cands[kShardDepth][kShardMaxRanks][kShardKCap];
};

struct ShardParams {
    ShardMailbox *box[kShardMaxRanks];   // [rank] = own mailbox, others = peers' (device-accessible)
    uint32_t rank, world;                // world == 0: not sharded
    unsigned long long seq;              // 1, 2, 3, ... identical on every rank for the same query
    unsigned long long timeout_ns;
    wax_vs_candidate *final_out;         // [k] merged result (device)
    wax_vs_candidate *host_out;          // [k] merged result in mapped pinned host memory, or nullptr
    unsigned long long *host_flag;       // mapped pinned: seq when host_out is complete (bit 63 = exchange timed out)
};

constexpr unsigned long long kShardErrorBit = 1ull << 63;

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Called by EVERY thread of one CTA.  `local` = this rank's sorted candidate list ([k], padding valid = 0 last), written
// to global memory by this CTA before the call.  `skeys` = shared-memory scratch of world * k uint32.
__device__ __noinline__ void shard_exchange_cta(const ShardParams &sh, const wax_vs_candidate *local, uint32_t k,
                                                uint32_t *skeys) {
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    const uint32_t slot = static_cast<uint32_t>(sh.seq % kShardDepth);
    ShardMailbox *mine = sh.box[sh.rank];
    __shared__ uint32_t s_err;
    if (tid == 0) s_err = 0;
    __syncthreads();   // also: the caller's writes of local[] are visible to the whole CTA

    // 1. the slot is free once every rank has finished reading what it held `depth` queries ago
    if (tid < sh.world && sh.seq > kShardDepth) {
        const unsigned long long need = sh.seq - kShardDepth, t0 = global_timer_ns();
        while (ld_acquire_sys_u64(&mine->acks[tid]) < need)
            if (global_timer_ns() - t0 > sh.timeout_ns) { s_err = 1; break; }
    }
    __syncthreads();

    // 2. push the local list into every rank's mailbox (posted NVLink writes; own mailbox included)
    for (uint32_t i = tid; i < sh.world * k; i += nthr) {
        const uint32_t r = i / k, j = i % k;
        const unsigned long long *src = reinterpret_cast<const unsigned long long *>(local + j);
        unsigned long long *dst = reinterpret_cast<unsigned long long *>(&sh.box[r]->cands[slot][sh.rank][j]);
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < sh.world) st_release_sys_u64(&sh.box[tid]->flags[slot][sh.rank], sh.seq);

    // 3. wait for every rank's list
    if (tid < sh.world) {
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys_u64(&mine->flags[slot][tid]) < sh.seq)
            if (global_timer_ns() - t0 > sh.timeout_ns) { s_err = 1; break; }
    }
    __syncthreads();
    const bool failed = s_err != 0;

    // 4. merge: distance keys of all lists in shared memory, final position by binary searches
    const uint32_t total = sh.world * k;
    for (uint32_t i = tid; i < total; i += nthr) {
        const unsigned long long w0 = ld_cg_u64(reinterpret_cast<const uint64_t *>(&mine->cands[slot][i / k][i % k]));
        const uint32_t valid = static_cast<uint32_t>(w0 >> 32);
        skeys[i] = (valid == 1u && !failed) ? orderable_u32(__uint_as_float(static_cast<uint32_t>(w0))) : WAXVS_UKEY_NONE;
    }
    __syncthreads();
    for (uint32_t i = tid; i < total; i += nthr) {
        const uint32_t r = i / k, j = i % k, key = skeys[i];
        uint32_t pos = j;
        for (uint32_t r2 = 0; r2 < sh.world && pos < k; ++r2) {
            if (r2 == r) continue;
            const uint32_t *lst = skeys + r2 * k;
            uint32_t lo = 0, hi = k;
            while (lo < hi) {                       // lower ranks hold lower global rows: their equal distances come first
                const uint32_t mid = (lo + hi) >> 1, v = lst[mid];
                const bool before = (r2 < r) ? (v <= key) : (v < key);
                if (before) lo = mid + 1; else hi = mid;
            }
            pos += lo;
        }
        if (pos < k) {
            const uint64_t *src = reinterpret_cast<const uint64_t *>(&mine->cands[slot][r][j]);
            unsigned long long w0 = ld_cg_u64(src), w1 = ld_cg_u64(src + 1), w2 = ld_cg_u64(src + 2);
            if (key == WAXVS_UKEY_NONE) { w0 = 0; w1 = 0; w2 = 0; }        // padding: valid = 0
            unsigned long long *dst = reinterpret_cast<unsigned long long *>(sh.final_out + pos);
            dst[0] = w0; dst[1] = w1; dst[2] = w2;
            if (sh.host_out) {
                unsigned long long *hd = reinterpret_cast<unsigned long long *>(sh.host_out + pos);
                hd[0] = w0; hd[1] = w1; hd[2] = w2;
            }
        }
    }
    __threadfence_system();
    __syncthreads();

    // 5. tell every rank this mailbox slot has been read; tell the host the result is complete
    if (tid < sh.world) st_release_sys_u64(&sh.box[tid]->acks[sh.rank], sh.seq);
    if (tid == 0 && sh.host_flag) st_release_sys_u64(sh.host_flag, failed ? (sh.seq | kShardErrorBit) : sh.seq);
}

// Stand-alone form (one CTA): the local list was produced by earlier work on the stream (an empty shard, a scan
// configuration whose shared-memory lists are too small for the fused form).
__global__ void __launch_bounds__(256) shard_exchange_kernel(const ShardParams sh, const wax_vs_candidate *local, uint32_t k) {
    __shared__ uint32_t skeys[kShardMaxRanks * kShardKCap];
    shard_exchange_cta(sh, local, k, skeys);
}

// ------------------------------------------------------------------------------------------------------------
// Batched form of the same merge (sharded search_batch): `gathered` = [world][n_queries][k] candidates as an all-gather
// of the ranks' per-query lists leaves them (every list sorted, padding valid = 0 last); out = [n_queries][k_out], the
// k_out best of each query under (distance, GLOBAL row) -- the position rule of shard_exchange_cta, binary searches in
// global memory.  One CTA per query.  Replaces the host-side numpy merge (7.6 ms per 1024 x 8 x 10 batch, ten times the
// shard's tensor-core pass at 8 GPUs).
__device__ __forceinline__ uint32_t cand_dist_key(const wax_vs_candidate &c) {
    return c.valid ? orderable_u32(c.distance) : WAXVS_UKEY_NONE;
}
__global__ void __launch_bounds__(128) merge_gathered_kernel(const wax_vs_candidate *__restrict__ gathered, uint32_t world,
                                                             uint32_t n_queries, uint32_t k, uint32_t k_out,
                                                             wax_vs_candidate *__restrict__ out) {
    const uint32_t q = blockIdx.x;
    const size_t rank_stride = static_cast<size_t>(n_queries) * k;
    const wax_vs_candidate *mine = gathered + static_cast<size_t>(q) * k;           // rank 0's list of this query
    for (uint32_t t = threadIdx.x; t < world * k; t += blockDim.x) {
        const uint32_t r = t / k, j = t % k;
        const wax_vs_candidate c = mine[r * rank_stride + j];
        const uint32_t key = cand_dist_key(c);
        uint32_t pos = j;
        for (uint32_t o = 0; o < world; ++o) {
            if (o == r) continue;
            const wax_vs_candidate *lst = mine + o * rank_stride;
            uint32_t lo = 0, hi = k;                 // lower ranks: entries <= key come first; higher ranks: entries < key
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                const uint32_t km = cand_dist_key(lst[mid]);
                if (o < r ? km <= key : km < key) lo = mid + 1; else hi = mid;
            }
            pos += lo;
        }
        if (pos < k_out) out[static_cast<size_t>(q) * k_out + pos] = c;
    }
}

}  // namespace waxvs
