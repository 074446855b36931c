// waxvs_batch.cuh -- batched queries (BASELINE configs 3 and 5): a genuine dense contraction, so it runs on the
// 5th-gen tensor cores.  score'[b][n] = sum_d Q[b][d] * V[n][d] as a skinny GEMM with tcgen05.mma kind::tf32:
//
//   A (M = 128 queries)  : TMA 2-D box [128 x 32 floats], 128-byte swizzle, K-major          (16 KB / k-block)
//   B (N = 256 rows)     : TMA 2-D box [256 x 32 floats] of the row-major corpus, K-major     (32 KB / k-block)
//   D                    : fp32 accumulators in TMEM, 128 lanes x 256 columns, double buffered (all 512 columns)
//   warp roles           : warps 0-3 epilogue (thread t owns query t = TMEM lane t), warp 4 TMA producer,
//                          warp 5 TMEM allocator + single-thread MMA issuer;  smem ring of 4 k-block stages.
//
// The score matrix (B x N, 40 GB at 1024 x 10 M) is never written: each epilogue thread scans its query's 256
// accumulator columns straight out of TMEM (tcgen05.ld), scales by the row's cached 1/|v| (cosine) and keeps
// the k' best (score', row) pairs of its row slice in a private max-heap (inserts are rare: O(k' ln(N/k'))).
//
// TF32 keeps 10 mantissa bits, which is not enough for the parity bar (scores within 1e-4, identical order),
// so the tensor-core pass only NOMINATES candidates: batch_finish_kernel merges the slices' lists per query,
// re-scores the k' nominees EXACTLY in fp32 with the very same accumulation order as the single-query kernels
// (bit-identical results), and proves nothing was missed: every row that was not nominated has
// score' <= tau, hence exact score <= tau + eps with eps the TF32 worst-case bound
// (|a_t b_t - ab| <= 2^-9 |ab| termwise  =>  |err| <= 2^-9 |q||v|); if the exact k-th score does not clear
// tau + eps the query is flagged and the host re-runs it on the exact single-query path.  Results are
// therefore always identical to the non-batched path; the tensor cores only buy speed.
//
// The reference has no batched search at all (VectorSearchEngine.swift:13 takes one vector); this is the
// "batched-query case where it is a genuine dense contraction" of BASELINE.json's north_star.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "waxvs_common.cuh"
#include "waxvs_scan.cuh"

namespace waxvs {

constexpr int kBatchM = 128;            // queries per CTA  (UMMA M)
constexpr int kBatchN = 256;            // corpus rows per tile (UMMA N)
constexpr int kBatchKBlock = 32;        // floats per k-block = one 128-byte swizzle atom
constexpr int kBatchKBlockBf16 = 64;    // bf16 elements per k-block (the same 128 bytes)
// Two shapes of the same kernel share the 227 KB of shared memory differently (template <STAGES, HEAP>):
//   <4, 16>: four TMA stages (192 KB) + 16-entry nominee heaps (16 KB)  -- the pipeline is latency-bound on TMA
//            (profiles/ncu_batch_tf32_r01b_summary.csv: 3 stages keep the tensor pipe 54 % busy), so the 4th stage
//            matters; used whenever 16 nominees per slice are plenty (16 * slices >= 8 * k);
//   <3, 64>: three stages + 64-entry heaps, for few slices or large k;
//   <4, 24> (bf16 only: no scale area): four stages + 24-entry heaps, when k exceeds the slice count -- configs[4]'s
//            top-100 over 74 slices left 1 query in 1 000 unproven with 16 nominees per slice (a slice that holds 16 rows
//            within the bf16 bound of the 100th score), and every unproven query costs the batch a second pass.
// PAIR = true is the cta_group::2 form: two CTAs of a cluster (two query groups, the same row slice) issue ONE
// 256 x 256 x 8 MMA; each CTA stages only its own 128 query rows and HALF of the corpus tile (16 + 16 KB per
// k-block instead of 16 + 32), so six stages fit where four did and the L2->SM traffic per SM drops by a third.
constexpr int kBatchRescore = 256;       // nominees of the union re-scored exactly per query (TF32 nominations)
constexpr int kBatchRescoreMax = 1024;   // upper bound (BF16 nominations with larger k re-score more, see eps)
constexpr uint32_t kBatchABytes = kBatchM * 128u;   // 16 KB
constexpr uint32_t kBatchBBytes = kBatchN * 128u;   // 32 KB
constexpr uint32_t kBatchStageBytes = kBatchABytes + kBatchBBytes;
constexpr int kBatchStageSlots = 8;      // staged nominees per epilogue thread before a forced flush
__host__ __device__ constexpr uint32_t batch_stage_bytes(bool pair) { return kBatchABytes + (pair ? kBatchBBytes / 2 : kBatchBBytes); }
__host__ __device__ constexpr uint32_t batch_smem_bytes(int stages, int heap, bool pair = false) {
    return stages * batch_stage_bytes(pair) + 2048 /*scales*/ + 256 /*barriers*/ + kBatchStageSlots * kBatchM * 8 /*staging*/ +
           heap * kBatchM * 8 /*heaps*/ + 1024 /*align*/;
}
// ARES (queries resident in shared memory): `ares_kb` k-blocks of the 128 queries stay in shared memory for the whole
// kernel (16 KB each: 6 k-blocks = 96 KB at 384 bf16 dims) and the ring stages carry the corpus only.
__host__ __device__ constexpr uint32_t batch_ares_stage_bytes(bool pair) { return pair ? kBatchBBytes / 2 : kBatchBBytes; }
__host__ __device__ constexpr uint32_t batch_ares_smem_bytes(int stages, int heap, bool pair, int ares_kb) {
    return ares_kb * kBatchABytes + stages * batch_ares_stage_bytes(pair) + 256 + kBatchStageSlots * kBatchM * 8 +
           heap * kBatchM * 8 + 1024;       // (bf16 only: no epilogue scale area)
}
constexpr int kBatchThreads = 192;
constexpr float kTf32Eps = 1.25f * 0x1p-9f;
// BF16 nominations: both operands are ROUNDED to nearest; bf16 keeps 8 significand bits (7 stored), so the unit
// round-off is 2^-8: |x~ - x| <= 2^-8 |x| each, |q~ v~ - q v| <= (2^-7 + 2^-16) |q v| termwise, hence
// |err| <= (2^-7 + 2^-16) sum |q_i v_i| <= (2^-7 + 2^-16) |q||v|; the fp32 accumulation adds O(dims * 2^-24) and the
// pre-normalisation of the cosine shadow rows O(2^-23).  1.03 * 2^-7 covers all of it.
constexpr float kBf16Eps = 1.03f * 0x1p-7f;

struct BatchParams {
    uint32_t n_rows, dims, n_queries;
    uint32_t groups;        // ceil(n_queries / 128)
    uint32_t slices;        // row slices; CTA b -> (group b % groups, slice b / groups)
    uint32_t tiles_total;   // ceil(n_rows / 256)
    uint32_t kprime;        // == HEAP of the kernel shape in use
    int metric;             // kCosine or kDot
    const float *row_scale; // [n_rows] 1/|v| (cosine) or nullptr
    uint64_t *heaps;        // [slices*groups][HEAP][128]: each CTA's heaps, dumped entry-major at the end
    uint32_t *tau_global;   // [n_queries] orderable(score') of the best k'-th nominee any slice has reached (0 = none)
    uint32_t no_insert;     // instrumentation: skip nominations (timing floor of the GEMM pipeline)
    // FILTER form (level 2): no heaps -- every row whose score' beats the query's FIXED threshold is appended to the
    // query's candidate list (complete by construction, see filter level below)
    const float *tau_fixed; // [n_queries]
    uint32_t *cand_count;   // [n_queries] appended so far (may exceed cand_cap: overflow)
    uint32_t *cand_rows;    // [n_queries][cand_cap]
    uint32_t cand_cap;
    // filtered batches (wax_vs_search_batch_filtered): 1 bit per row, set = the row may be returned; nullptr = all rows.
    // Consulted only on the rare path (a chunk of 32 rows that holds a score above the query's threshold).
    const uint32_t *allow_bits;
};

// ---- PTX wrappers (tcgen05 / TMA tensor) ---------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int32_t c0,
                                            int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, one CTA.
__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind::f16 (here: bf16 x bf16 -> fp32), one CTA: K = 16 elements = the same 32 bytes per MMA as tf32's K = 8.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 columns of 32-bit accumulators -> 32 registers (thread = lane, register j = column j).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- 2-CTA (cta_group::2) variants: the CTA pair of a cluster works as one 256-row MMA --------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load issued by either CTA of the pair into its OWN shared memory; the transaction bytes are credited to the
// LEADER CTA's mbarrier (peer bit of the shared::cluster address cleared, as cute::SM100_TMA_2SM_LOAD_2D does).
__device__ __forceinline__ void tma_load_2d_pair(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int32_t c0,
                                                 int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
// mbarrier.arrive on the barrier at the same offset in CTA `rank` of the cluster.
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 remote;\n\t"
        "mapa.shared::cluster.u32 remote, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [remote];\n\t}" ::"r"(smem_u32(bar)),
        "r"(rank)
        : "memory");
}
__device__ __forceinline__ void tcgen05_commit_pair(uint64_t *bar) {   // arrives on `bar` in BOTH CTAs of the pair
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(static_cast<uint16_t>(3))
        : "memory");
}
__device__ __forceinline__ void umma_tf32_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                  uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void umma_bf16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                  uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Shared-memory matrix descriptor, K-major, 128-byte swizzle (canonical layout ((8,n),2):((8,SBO),1) in 16-byte
// units; rows 128 B apart, 8-row groups SBO = 1024 B apart).  Field layout: cute::UMMA::SmemDescriptor.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(const void *smem_tile) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_u32(smem_tile) >> 4) & 0x3FFFu);  // start address  [0,14)
    d |= static_cast<uint64_t>(1) << 16;                               // LBO (unused for swizzled K-major) [16,30)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;                       // SBO [32,46)
    d |= static_cast<uint64_t>(1) << 46;                               // descriptor version (Blackwell) [46,48)
    d |= static_cast<uint64_t>(2) << 61;                               // layout type SWIZZLE_128B [61,64)
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, M = 128, N = 256.
__host__ __device__ constexpr uint32_t umma_idesc_tf32_m128_n256() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
}
// cta_group::2: M = 256 (128 rows per CTA of the pair), N = 256.
__host__ __device__ constexpr uint32_t umma_idesc_tf32_m256_n256() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);
}

// kind::f16 with A = B = BF16 (format 1), D = F32.
__host__ __device__ constexpr uint32_t umma_idesc_bf16_m128_n256() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_m256_n256() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);
}

// max of three (FMNMX3 on sm_100; like fmaxf, a NaN input is dropped).
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// ---- per-thread nominee heap (max-heap on the ordering key: root = worst nominee) ----------------------------
// key = (orderable(-score') << 32) | row  -- smaller is better, exactly like the exact keys.
__device__ __forceinline__ uint64_t nominee_key(float score, uint32_t row) {
    return (static_cast<uint64_t>(orderable_u32(-score)) << 32) | row;
}
__device__ __forceinline__ float nominee_score(uint64_t key) { return -from_orderable_u32(static_cast<uint32_t>(key >> 32)); }

// `heap` points at this thread's node 0 in shared memory; node i lives at heap[i * kBatchM] (entry-major, so the
// 32 lanes of a warp touching the same level hit 32 different banks).  Returns the new root.
template <int HEAP>
__device__ __forceinline__ uint64_t heap_replace_root(uint64_t *heap, uint64_t x) {
    uint32_t i = 0;
#pragma unroll 1
    for (;;) {
        const uint32_t l = 2 * i + 1;
        if (l >= HEAP) break;
        const uint32_t r = l + 1;
        const uint64_t kl = heap[l * kBatchM];
        const uint64_t kr = (r < HEAP) ? heap[r * kBatchM] : 0ull;
        const uint32_t c = (kr > kl) ? r : l;
        const uint64_t kc = (kr > kl) ? kr : kl;
        if (kc <= x) break;
        heap[i * kBatchM] = kc;
        i = c;
    }
    heap[i * kBatchM] = x;
    return i == 0 ? x : heap[0];
}

// ---- row norms (cached per corpus version) --------------------------------------------------------------------
// One warp per row, same accumulation order as the scan kernels.  inv_norm = 1/sqrt(sum v^2) (0 for a zero
// row); max_norm_bits = max over finite rows of sqrt(sum v^2) as float bits (positive floats order as uints).
__global__ void __launch_bounds__(256) row_norms_kernel(const float *corpus, uint32_t n_rows, uint32_t dims,
                                                        float *inv_norm, uint32_t *max_norm_bits) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    const bool vec4 = (dims % 4u) == 0u;
    float local_max = 0.0f;
    for (uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n_rows; row += warps) {
        const float *v = corpus + static_cast<size_t>(row) * dims;
        float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
        if (vec4) {
            const float4 *v4 = reinterpret_cast<const float4 *>(v);
            for (uint32_t c = lane; c < dims / 4u; c += 32u) {
                const float4 x = __ldg(v4 + c);
                b0 = __fmaf_rn(x.x, x.x, b0); b1 = __fmaf_rn(x.y, x.y, b1);
                b2 = __fmaf_rn(x.z, x.z, b2); b3 = __fmaf_rn(x.w, x.w, b3);
            }
        } else {
            for (uint32_t base = 4u * lane; base < dims; base += 128u) {
                const float x = __ldg(v + base); b0 = __fmaf_rn(x, x, b0);
                if (base + 1 < dims) { const float y = __ldg(v + base + 1); b1 = __fmaf_rn(y, y, b1); }
                if (base + 2 < dims) { const float z = __ldg(v + base + 2); b2 = __fmaf_rn(z, z, b2); }
                if (base + 3 < dims) { const float w = __ldg(v + base + 3); b3 = __fmaf_rn(w, w, b3); }
            }
        }
        const float s = warp_butterfly_sum(__fadd_rn(__fadd_rn(b0, b1), __fadd_rn(b2, b3)));
        const float nrm = __fsqrt_rn(s);
        if (lane == 0) inv_norm[row] = (s == 0.0f) ? 0.0f : __fdiv_rn(1.0f, nrm);
        if (finite_f32(nrm)) local_max = fmaxf(local_max, nrm);
    }
    if (lane == 0 && local_max > 0.0f) atomicMax(max_norm_bits, __float_as_uint(local_max));
}

// ---- bf16 shadow of the corpus (cached per corpus version, like the norms) ------------------------------------------
// dst[row][d] = bf16_rn(src[row][d] * scale[row]) (cosine: scale = 1/|v|, so the nomination scores need no epilogue
// scaling; dot: scale == nullptr).  Only ever NOMINATES: every returned score is recomputed from the fp32 corpus.
__global__ void __launch_bounds__(256) shadow_bf16_kernel(const float *__restrict__ src, const float *__restrict__ scale,
                                                          uint64_t n_rows, uint32_t dims, __nv_bfloat16 *__restrict__ dst) {
    const uint32_t d4 = dims >> 2;                       // dims % 4 == 0 (the tensor path needs dims % 64 == 0)
    const uint64_t total = n_rows * d4;
    const float4 *s4 = reinterpret_cast<const float4 *>(src);
    uint2 *o = reinterpret_cast<uint2 *>(dst);
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        float4 x = __ldcs(s4 + i);
        if (scale) {
            const float w = __ldg(scale + i / d4);
            x.x *= w; x.y *= w; x.z *= w; x.w *= w;
        }
        const __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
        uint2 v;
        v.x = *reinterpret_cast<const uint32_t *>(&lo);
        v.y = *reinterpret_cast<const uint32_t *>(&hi);
        // A finite fp32 within 2^-8 of FLT_MAX rounds to bf16 infinity: clamp it to the largest finite bf16 instead
        // (relative error still <= 2^-8), so that rounding alone can never turn a finite row into inf / NaN scores.
        auto clamp_half = [](uint32_t h, float src) -> uint32_t {   // h: one bf16 in the low 16 bits
            return ((h & 0x7FFFu) == 0x7F80u && finite_f32(src)) ? ((h & 0x8000u) | 0x7F7Fu) : h;
        };
        v.x = clamp_half(v.x & 0xFFFFu, x.x) | (clamp_half(v.x >> 16, x.y) << 16);
        v.y = clamp_half(v.y & 0xFFFFu, x.z) | (clamp_half(v.y >> 16, x.w) << 16);
        o[i] = v;
    }
}

// ---- the tensor-core kernel ---------------------------------------------------------------------------------------
// BF16: operands are bf16 (the corpus shadow + converted queries; 64 elements per 128-byte k-block, kind::f16 MMAs at
//       twice the TF32 rate for the same bytes per cycle) -- nominations only, exactness comes from the finish kernel.
// ARES: the CTA's 128 queries stay resident in shared memory (dims/64 bf16 k-blocks of 16 KB, loaded once), the ring
//       stages carry only corpus tiles: a third less L2->SM and TMA->smem traffic per MMA.
// FILTER: the epilogue appends every row beating the query's fixed threshold to a per-query list instead of keeping
//       the k' best in a heap (the filter level: complete by construction).
template <int STAGES, int HEAP, bool PAIR, bool BF16 = false, bool ARES = false, bool FILTER = false>
__global__ void __launch_bounds__(kBatchThreads, 1)
batch_nominate_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
                      const BatchParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128B-swizzled tiles, computed as an OFFSET into the shared array so the compiler
    // keeps the shared address space (LDS/STS, not generic LD/ST).
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr uint32_t STAGE_BYTES = ARES ? batch_ares_stage_bytes(PAIR) : batch_stage_bytes(PAIR);
    constexpr uint32_t B_OFF = ARES ? 0u : kBatchABytes;            // corpus tile offset inside a stage
    constexpr uint32_t B_ROWS = PAIR ? kBatchN / 2 : kBatchN;       // corpus rows this CTA stages per tile
    constexpr uint32_t KB_ELEMS = BF16 ? kBatchKBlockBf16 : kBatchKBlock;   // elements per 128-byte k-block
    const uint32_t num_kb = p.dims / KB_ELEMS;
    const uint32_t ares_bytes = ARES ? num_kb * kBatchABytes : 0u;   // resident queries: [kb][128 rows x 128 B]
    uint8_t *a_res = smem;
    uint8_t *stages = smem + ares_bytes;                               // [stage][A 16 KB | B 32 KB], 1024-aligned
    // bf16 nominations never scale in the epilogue (the cosine shadow rows are pre-normalised): no scale area, which is
    // what lets the <4 stages, 24-entry heaps> shape fit
    constexpr uint32_t SCALE_BYTES = BF16 ? 0u : 2u * kBatchN * 4u;
    float *scale_smem = reinterpret_cast<float *>(stages + STAGES * STAGE_BYTES);   // [2][256]
    uint64_t *full = reinterpret_cast<uint64_t *>(stages + STAGES * STAGE_BYTES + SCALE_BYTES);   // [stages]
    uint64_t *empty = full + STAGES;                                                   // [stages]
    uint64_t *tmem_full = empty + STAGES;                                              // [2]
    uint64_t *tmem_empty = tmem_full + 2;                                                    // [2]
    uint64_t *a_full = tmem_empty + 2;                                                       // [1] (ARES)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(a_full + 1);
    uint64_t *stage_smem = reinterpret_cast<uint64_t *>(stages + STAGES * STAGE_BYTES + SCALE_BYTES + 256);  // [slots][128]
    uint64_t *heap_smem = stage_smem + kBatchStageSlots * kBatchM;                                            // [64][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // PAIR: cluster c = (pair of groups c % (groups/2), slice c / (groups/2)); the CTA's rank picks the group.
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const uint32_t unit = PAIR ? blockIdx.x / 2u : blockIdx.x;
    const uint32_t units_per_slice = PAIR ? p.groups / 2u : p.groups;
    const uint32_t group = PAIR ? (unit % units_per_slice) * 2u + rank : unit % units_per_slice;
    const uint32_t slice = unit / units_per_slice;
    const bool leader = rank == 0u;
    const uint32_t tile_lo = static_cast<uint32_t>(static_cast<uint64_t>(p.tiles_total) * slice / p.slices);
    const uint32_t tile_hi = static_cast<uint32_t>(static_cast<uint64_t>(p.tiles_total) * (slice + 1) / p.slices);

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_c);
        // PAIR: the leader's full[] collects both producers (its own expect_tx arrive + the peer's remote arrive) and
        // its tmem_empty[] collects the 4 epilogue warps of both CTAs.
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], PAIR ? 2 : 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], PAIR ? 8 : 4); }
        mbar_init(a_full, PAIR ? 2 : 1);
        mbar_fence_init();
    }
    if (warp == 5) {  // whole warp: allocate all 512 TMEM columns (2 accumulator buffers of 256)
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"(512)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"(512)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tcgen05_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ===== TMA producer =====
        if (lane == 0) {
            if (ARES && tile_lo < tile_hi) {      // the CTA's queries, once: num_kb boxes of 128 rows x 128 B
                if (PAIR) {
                    if (leader) mbar_arrive_expect_tx(a_full, 2u * ares_bytes);
                    else mbar_arrive_remote(a_full, 0u);
                } else {
                    mbar_arrive_expect_tx(a_full, ares_bytes);
                }
                for (uint32_t kb = 0; kb < num_kb; ++kb) {
                    if (PAIR) tma_load_2d_pair(a_res + kb * kBatchABytes, &tmap_q, a_full, static_cast<int32_t>(kb * KB_ELEMS),
                                               static_cast<int32_t>(group * kBatchM));
                    else tma_load_2d(a_res + kb * kBatchABytes, &tmap_q, a_full, static_cast<int32_t>(kb * KB_ELEMS),
                                     static_cast<int32_t>(group * kBatchM));
                }
            }
            uint32_t stage = 0, phase = 0;
            for (uint32_t tile = tile_lo; tile < tile_hi; ++tile) {
                for (uint32_t kb = 0; kb < num_kb; ++kb) {
                    mbar_wait_parity(&empty[stage], phase ^ 1u);
                    uint8_t *a = stages + stage * STAGE_BYTES;
                    if (PAIR) {
                        if (leader) mbar_arrive_expect_tx(&full[stage], 2u * STAGE_BYTES);   // both CTAs' bytes
                        else mbar_arrive_remote(&full[stage], 0u);
                        if (!ARES) tma_load_2d_pair(a, &tmap_q, &full[stage], static_cast<int32_t>(kb * KB_ELEMS),
                                                    static_cast<int32_t>(group * kBatchM));
                        tma_load_2d_pair(a + B_OFF, &tmap_c, &full[stage], static_cast<int32_t>(kb * KB_ELEMS),
                                         static_cast<int32_t>(tile * kBatchN + rank * B_ROWS));
                    } else {
                        mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
                        if (!ARES) tma_load_2d(a, &tmap_q, &full[stage], static_cast<int32_t>(kb * KB_ELEMS),
                                               static_cast<int32_t>(group * kBatchM));
                        tma_load_2d(a + B_OFF, &tmap_c, &full[stage], static_cast<int32_t>(kb * KB_ELEMS),
                                    static_cast<int32_t>(tile * kBatchN));
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0 && leader) {      // PAIR: the leader CTA issues for both
            constexpr uint32_t idesc = BF16 ? (PAIR ? umma_idesc_bf16_m256_n256() : umma_idesc_bf16_m128_n256())
                                            : (PAIR ? umma_idesc_tf32_m256_n256() : umma_idesc_tf32_m128_n256());
            if (ARES && tile_lo < tile_hi) { mbar_wait_parity(a_full, 0u); tcgen05_fence_after(); }
            uint32_t stage = 0, phase = 0, t = 0;
            for (uint32_t tile = tile_lo; tile < tile_hi; ++tile, ++t) {
                const uint32_t acc = t & 1u, acc_phase = (t >> 1) & 1u;
                mbar_wait_parity(&tmem_empty[acc], acc_phase ^ 1u);   // epilogue has drained this buffer
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * kBatchN;
                for (uint32_t kb = 0; kb < num_kb; ++kb) {
                    mbar_wait_parity(&full[stage], phase);            // TMA bytes have landed
                    tcgen05_fence_after();
                    const uint8_t *a = stages + stage * STAGE_BYTES;
                    const uint64_t da = umma_desc_k_sw128(ARES ? a_res + kb * kBatchABytes : a);
                    const uint64_t db = umma_desc_k_sw128(a + B_OFF);
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j) {   // UMMA K = 8 tf32 / 16 bf16 = 32 bytes = +2 in the address field
                        const uint32_t accum = (kb | j) != 0u ? 1u : 0u;
                        if (BF16) {
                            if (PAIR) umma_bf16_ss_pair(d_tmem, da + 2 * j, db + 2 * j, idesc, accum);
                            else umma_bf16_ss(d_tmem, da + 2 * j, db + 2 * j, idesc, accum);
                        } else {
                            if (PAIR) umma_tf32_ss_pair(d_tmem, da + 2 * j, db + 2 * j, idesc, accum);
                            else umma_tf32_ss(d_tmem, da + 2 * j, db + 2 * j, idesc, accum);
                        }
                    }
                    if (PAIR) tcgen05_commit_pair(&empty[stage]);     // frees the stage in both CTAs
                    else tcgen05_commit(&empty[stage]);               // frees the smem stage when the MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                if (PAIR) tcgen05_commit_pair(&tmem_full[acc]);       // accumulators complete -> both epilogues
                else tcgen05_commit(&tmem_full[acc]);                 // accumulator complete -> epilogue
            }
        }
    } else {
        // ===== epilogue: thread t <-> query group*128 + t <-> TMEM lane t =====
        // A nominee costs one compare against tau in the hot loop; winners are STAGED in shared memory and the
        // whole warp flushes together (every lane sifts its own heap concurrently) so a lane's insert never idles
        // the other 31.  tau is also shared across the slices of a query through tau_global: any slice's k'-th
        // best is a valid filter for all of them (the union then holds >= k' nominees at or above it).
        const uint32_t tid = threadIdx.x;                              // 0..127
        const uint32_t q = group * kBatchM + tid;
        const bool q_valid = q < p.n_queries && !p.no_insert;
        uint64_t *heap = heap_smem + tid;
        if (!FILTER) for (uint32_t i = 0; i < HEAP; ++i) heap[i * kBatchM] = WAXVS_KEY_NONE;
        uint64_t root = WAXVS_KEY_NONE;                               // heap[0]: this slice's k'-th best so far
        float tau = -INFINITY;
        if (FILTER && q_valid) tau = __ldg(p.tau_fixed + q);          // never changes: the list is a pure filter
        uint64_t *stage = stage_smem + tid;                           // slot i at stage[i * 128]
        uint32_t cnt = 0;
        bool improved = false;
        auto flush = [&]() {
            if (FILTER) {                                             // staged rows -> the query's global list
                if (cnt) {
                    const uint32_t base = atomicAdd(p.cand_count + q, cnt);
                    for (uint32_t i = 0; i < cnt; ++i)
                        if (base + i < p.cand_cap)
                            p.cand_rows[static_cast<size_t>(q) * p.cand_cap + base + i] = static_cast<uint32_t>(stage[i * kBatchM]);
                }
                cnt = 0;
                return;
            }
            for (uint32_t i = 0; i < cnt; ++i) {
                const uint64_t x = stage[i * kBatchM];
                if (x < root) { root = heap_replace_root<HEAP>(heap, x); improved = true; }
            }
            cnt = 0;
            if (root != WAXVS_KEY_NONE) tau = fmaxf(tau, nominee_score(root));
        };
        const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
        // The shared threshold is read one tile AHEAD (the L2 round trip of __ldcg would otherwise sit on every
        // tile's critical path: with bf16 MMAs a tile lasts ~3000 cycles and the epilogue has no slack to hide it).
        uint32_t g_next = 0;
        uint32_t t = 0;
        for (uint32_t tile = tile_lo; tile < tile_hi; ++tile, ++t) {
            const uint32_t acc = t & 1u, acc_phase = (t >> 1) & 1u;
            const uint32_t row0 = tile * kBatchN;
            float *sc = scale_smem + acc * kBatchN;
            if (!BF16 && p.row_scale) {
#pragma unroll
                for (uint32_t h = 0; h < 2; ++h) {
                    const uint32_t r = row0 + tid + h * 128u;
                    sc[tid + h * 128u] = (r < p.n_rows) ? __ldg(p.row_scale + r) : 0.0f;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");        // scales visible; previous use of sc[] finished
            }
            if (!FILTER && g_next) tau = fmaxf(tau, from_orderable_u32(g_next)); // adopt the best threshold any slice has published
            mbar_wait_parity(&tmem_full[acc], acc_phase);
            tcgen05_fence_after();
            if (!FILTER && q_valid) g_next = __ldcg(p.tau_global + q); // consumed at the next tile
            const uint32_t rows_here = min(static_cast<uint32_t>(kBatchN), p.n_rows - row0);
#pragma unroll 1
            for (uint32_t chunk = 0; chunk < kBatchN / 32; ++chunk) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + lane_base + acc * kBatchN + chunk * 32u, v);
                if (chunk * 32u >= rows_here) continue;               // warp-uniform
                // Hot path, branch-free: scale the 32 scores and take their max (max.f32 drops NaNs: FMNMX3, 17
                // instructions for 32 values); only a chunk whose max beats tau (rare once the heap has warmed up)
                // is examined column by column.  The rare path is deliberately COMPACT: the first version walked a
                // fully unrolled max tree with the flush inlined at every leaf -- 200 KB of SASS, so every entry
                // missed the instruction cache, which cost a quarter of the kernel once bf16 halved the MMA time.
                float sv[32];
                if (!BF16 && p.row_scale) {
                    const float4 *sc4 = reinterpret_cast<const float4 *>(sc + chunk * 32u);
#pragma unroll
                    for (uint32_t j4 = 0; j4 < 8; ++j4) {
                        const float4 w = sc4[j4];
                        sv[4 * j4 + 0] = __uint_as_float(v[4 * j4 + 0]) * w.x;
                        sv[4 * j4 + 1] = __uint_as_float(v[4 * j4 + 1]) * w.y;
                        sv[4 * j4 + 2] = __uint_as_float(v[4 * j4 + 2]) * w.z;
                        sv[4 * j4 + 3] = __uint_as_float(v[4 * j4 + 3]) * w.w;
                    }
                } else {
#pragma unroll
                    for (uint32_t j = 0; j < 32; ++j) sv[j] = __uint_as_float(v[j]);
                }
                float t11[11];
#pragma unroll
                for (uint32_t j = 0; j < 10; ++j) t11[j] = fmax3(sv[3 * j], sv[3 * j + 1], sv[3 * j + 2]);
                t11[10] = fmaxf(sv[30], sv[31]);
                const float u0 = fmax3(t11[0], t11[1], t11[2]), u1 = fmax3(t11[3], t11[4], t11[5]);
                const float u2 = fmax3(t11[6], t11[7], t11[8]), u3 = fmaxf(t11[9], t11[10]);
                const float cmax = fmaxf(fmax3(u0, u1, u2), u3);
                if (cmax > tau && q_valid) {
                    uint32_t mask = 0;
#pragma unroll
                    for (uint32_t j = 0; j < 32; ++j) mask |= (sv[j] > tau ? 1u : 0u) << j;
                    const uint32_t cols = rows_here - chunk * 32u;                 // >= 1 here
                    if (cols < 32u) mask &= (1u << cols) - 1u;
                    // row0 + chunk * 32 is a multiple of 32: the chunk's 32 rows are exactly one word of the row filter
                    if (p.allow_bits) mask &= __ldg(p.allow_bits + ((row0 + chunk * 32u) >> 5));
                    while (mask) {
                        if (cnt + __popc(mask) > kBatchStageSlots) flush();       // empties the slots, may raise tau
                        uint32_t take = mask;
                        if (__popc(mask) > kBatchStageSlots) {                     // warm-up only: lowest 8 set bits
                            take = 0;
#pragma unroll 1
                            for (int i = 0; i < kBatchStageSlots; ++i) { const uint32_t bit = mask & (0u - mask); take |= bit; mask ^= bit; }
                        } else {
                            mask = 0;
                        }
#pragma unroll
                        for (uint32_t j = 0; j < 32; ++j) {
                            if ((take >> j) & 1u) {
                                if (sv[j] > tau) { stage[cnt * kBatchM] = nominee_key(sv[j], row0 + chunk * 32u + j); ++cnt; }
                            }
                        }
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {                                          // accumulator buffer is free again
                if (PAIR) mbar_arrive_remote(&tmem_empty[acc], 0u);   // the leader's barrier gates the shared MMA
                else mbar_arrive(&tmem_empty[acc]);
            }
            if (__any_sync(WAXVS_FULL_MASK, cnt >= kBatchStageSlots / 2)) {
                flush();
                if (!FILTER && improved && q_valid && root != WAXVS_KEY_NONE) {  // heap full: publish this slice's k'-th best
                    atomicMax(p.tau_global + q, orderable_u32(nominee_score(root)));
                    improved = false;
                }
            }
        }
        flush();
        if (!FILTER) {
            // dump this CTA's heaps (entry-major, coalesced) for batch_finish_kernel
            uint64_t *dst = p.heaps + static_cast<size_t>(slice * p.groups + group) * HEAP * kBatchM + tid;
            for (uint32_t i = 0; i < HEAP; ++i) dst[i * kBatchM] = heap[i * kBatchM];
        }
    }
    tcgen05_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();   // PAIR: the peer may still arrive on / read this CTA's shared memory
    if (warp == 5) {
        tcgen05_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}


// ---- TS + pair kernel: the queries live in TMEM ---------------------------------------------------------------------
// Removes the shared-memory bandwidth bound of the SS shapes: the A operand (128 queries x dims
// tf32 per CTA) is written ONCE into tensor memory (tcgen05.st, `dims` <= 384 columns) and every MMA reads it from
// there (`tcgen05.mma ... [d], [a_tmem], b_desc`), so shared memory only carries the corpus: with the CTA pair each
// CTA stages 32 rows x 32 floats = 4 KB per k-block per 128 MMA cycles (64 B/clk in + out instead of 192).
//   TMEM columns: [0, dims) queries | [384, 448) accumulators 0 | [448, 512) accumulators 1   (N = 64 rows per tile)
constexpr int kTsN = 64;                 // corpus rows per tile (UMMA N); each CTA of the pair stages half
constexpr int kTsKbPerStage = 4;         // k-blocks per stage: 16 small MMAs per barrier round trip (a single thread
                                         // cannot wait + commit every 128 cycles), so dims % 128 == 0
constexpr int kTsStages = 6;
constexpr uint32_t kTsBoxBytes = (kTsN / 2) * 128u;               // one k-block of this CTA's half tile: 4 KB
constexpr uint32_t kTsStageBytes = kTsKbPerStage * kTsBoxBytes;  // 16 KB
constexpr uint32_t kTsAccCol = 384;
__host__ __device__ constexpr uint32_t batch_ts_smem_bytes(int heap) {
    return kTsStages * kTsStageBytes + 2 * kTsN * 4 /*scales*/ + 1024 /*barriers*/ + kBatchStageSlots * kBatchM * 8 +
           heap * kBatchM * 8 + 1024 /*align*/;
}
__host__ __device__ constexpr uint32_t umma_idesc_tf32_m256_n64() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((256u >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32_ts_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                                  uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}

template <int HEAP>
__global__ void __launch_bounds__(kBatchThreads, 1)
batch_tf32_ts_kernel(const __grid_constant__ CUtensorMap tmap_c, const float *__restrict__ queries, const BatchParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *stages = smem;                                                            // [stage][32 rows x 128 B]
    float *scale_smem = reinterpret_cast<float *>(smem + kTsStages * kTsStageBytes);   // [2][64]
    uint64_t *full = reinterpret_cast<uint64_t *>(scale_smem + 2 * kTsN);              // [stages]
    uint64_t *empty = full + kTsStages;                                                // [stages]
    uint64_t *tmem_full = empty + kTsStages;                                           // [2]
    uint64_t *tmem_empty = tmem_full + 2;                                              // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    uint64_t *stage_smem = reinterpret_cast<uint64_t *>(smem + kTsStages * kTsStageBytes + 2 * kTsN * 4 + 1024);
    uint64_t *heap_smem = stage_smem + kBatchStageSlots * kBatchM;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t unit = blockIdx.x / 2u, units_per_slice = p.groups / 2u;
    const uint32_t group = (unit % units_per_slice) * 2u + rank, slice = unit / units_per_slice;
    const bool leader = rank == 0u;
    const uint32_t tile_lo = static_cast<uint32_t>(static_cast<uint64_t>(p.tiles_total) * slice / p.slices);
    const uint32_t tile_hi = static_cast<uint32_t>(static_cast<uint64_t>(p.tiles_total) * (slice + 1) / p.slices);
    const uint32_t num_kb = p.dims / kBatchKBlock;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmap_c);
        for (int s = 0; s < kTsStages; ++s) { mbar_init(&full[s], 2); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 8); }
        mbar_fence_init();
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // queries -> TMEM: thread t writes query (group*128 + t) into lane t, columns [0, dims) (zeros if out of range)
        const uint32_t q = group * kBatchM + threadIdx.x;
        const float4 *src = reinterpret_cast<const float4 *>(queries + static_cast<size_t>(q) * p.dims);
        const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
        for (uint32_t c0 = 0; c0 < p.dims; c0 += 32u) {
            uint32_t v[32];
#pragma unroll
            for (uint32_t j4 = 0; j4 < 8; ++j4) {
                float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                if (q < p.n_queries) x = __ldg(src + (c0 >> 2) + j4);
                v[4 * j4 + 0] = __float_as_uint(x.x); v[4 * j4 + 1] = __float_as_uint(x.y);
                v[4 * j4 + 2] = __float_as_uint(x.z); v[4 * j4 + 3] = __float_as_uint(x.w);
            }
            tmem_st_32x32(tmem_base + lane_base + c0, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    cluster_sync_all();            // barriers initialised and both CTAs' queries resident before any MMA / remote arrive
    tcgen05_fence_after();

    if (warp == 4) {
        // ===== TMA producer (both CTAs: each stages its half of the corpus tile) =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t tile = tile_lo; tile < tile_hi; ++tile) {
                for (uint32_t kb = 0; kb < num_kb; kb += kTsKbPerStage) {
                    mbar_wait_parity(&empty[stage], phase ^ 1u);
                    if (leader) mbar_arrive_expect_tx(&full[stage], 2u * kTsStageBytes);
                    else mbar_arrive_remote(&full[stage], 0u);
#pragma unroll
                    for (uint32_t i = 0; i < kTsKbPerStage; ++i)
                        tma_load_2d_pair(stages + stage * kTsStageBytes + i * kTsBoxBytes, &tmap_c, &full[stage],
                                         static_cast<int32_t>((kb + i) * kBatchKBlock),
                                         static_cast<int32_t>(tile * kTsN + rank * (kTsN / 2)));
                    if (++stage == kTsStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer: the leader CTA's single thread issues the 256-row MMAs for the pair =====
        if (lane == 0 && leader) {
            constexpr uint32_t idesc = umma_idesc_tf32_m256_n64();
            uint32_t stage = 0, phase = 0, t = 0;
            for (uint32_t tile = tile_lo; tile < tile_hi; ++tile, ++t) {
                const uint32_t acc = t & 1u, acc_phase = (t >> 1) & 1u;
                mbar_wait_parity(&tmem_empty[acc], acc_phase ^ 1u);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + kTsAccCol + acc * kTsN;
                for (uint32_t kb = 0; kb < num_kb; kb += kTsKbPerStage) {
                    mbar_wait_parity(&full[stage], phase);
                    tcgen05_fence_after();
#pragma unroll
                    for (uint32_t i = 0; i < kTsKbPerStage; ++i) {
                        const uint64_t db = umma_desc_k_sw128(stages + stage * kTsStageBytes + i * kTsBoxBytes);
#pragma unroll
                        for (uint32_t j = 0; j < kBatchKBlock / 8; ++j)   // A: 8 tf32 = 8 TMEM columns per K step
                            umma_tf32_ts_pair(d_tmem, tmem_base + (kb + i) * kBatchKBlock + j * 8u, db + 2 * j, idesc,
                                              (kb | i | j) != 0u ? 1u : 0u);
                    }
                    tcgen05_commit_pair(&empty[stage]);
                    if (++stage == kTsStages) { stage = 0; phase ^= 1u; }
                }
                tcgen05_commit_pair(&tmem_full[acc]);
            }
        }
    } else {
        // ===== epilogue (same nomination scheme as batch_tf32_kernel, 64 columns per tile) =====
        const uint32_t tid = threadIdx.x;
        const uint32_t q = group * kBatchM + tid;
        const bool q_valid = q < p.n_queries && !p.no_insert;
        uint64_t *heap = heap_smem + tid;
        for (uint32_t i = 0; i < HEAP; ++i) heap[i * kBatchM] = WAXVS_KEY_NONE;
        uint64_t root = WAXVS_KEY_NONE;
        float tau = -INFINITY;
        uint64_t *stage = stage_smem + tid;
        uint32_t cnt = 0;
        bool improved = false;
        auto flush = [&]() {
            for (uint32_t i = 0; i < cnt; ++i) {
                const uint64_t x = stage[i * kBatchM];
                if (x < root) { root = heap_replace_root<HEAP>(heap, x); improved = true; }
            }
            cnt = 0;
            if (root != WAXVS_KEY_NONE) tau = fmaxf(tau, nominee_score(root));
        };
        const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
        // the row scales of the NEXT tile are fetched while the current one is processed
        float next_scale = 0.0f;
        if (p.row_scale && tid < kTsN && tile_lo < tile_hi) {
            const uint32_t r = tile_lo * kTsN + tid;
            next_scale = (r < p.n_rows) ? __ldg(p.row_scale + r) : 0.0f;
        }
        uint32_t t = 0;
        for (uint32_t tile = tile_lo; tile < tile_hi; ++tile, ++t) {
            const uint32_t acc = t & 1u, acc_phase = (t >> 1) & 1u;
            const uint32_t row0 = tile * kTsN;
            float *sc = scale_smem + acc * kTsN;
            if (p.row_scale && tid < kTsN) {
                sc[tid] = next_scale;
                const uint32_t r = row0 + kTsN + tid;
                next_scale = (tile + 1 < tile_hi && r < p.n_rows) ? __ldg(p.row_scale + r) : 0.0f;
            }
            if (q_valid && (t & 7u) == 0u) {
                const uint32_t g = __ldcg(p.tau_global + q);
                if (g) tau = fmaxf(tau, from_orderable_u32(g));
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            mbar_wait_parity(&tmem_full[acc], acc_phase);
            tcgen05_fence_after();
            const uint32_t rows_here = min(static_cast<uint32_t>(kTsN), p.n_rows - row0);
#pragma unroll 1
            for (uint32_t chunk = 0; chunk < kTsN / 32; ++chunk) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + lane_base + kTsAccCol + acc * kTsN + chunk * 32u, v);
                if (chunk * 32u >= rows_here) continue;
                float sv[32];
                if (p.row_scale) {
                    const float4 *sc4 = reinterpret_cast<const float4 *>(sc + chunk * 32u);
#pragma unroll
                    for (uint32_t j4 = 0; j4 < 8; ++j4) {
                        const float4 w = sc4[j4];
                        sv[4 * j4 + 0] = __uint_as_float(v[4 * j4 + 0]) * w.x;
                        sv[4 * j4 + 1] = __uint_as_float(v[4 * j4 + 1]) * w.y;
                        sv[4 * j4 + 2] = __uint_as_float(v[4 * j4 + 2]) * w.z;
                        sv[4 * j4 + 3] = __uint_as_float(v[4 * j4 + 3]) * w.w;
                    }
                } else {
#pragma unroll
                    for (uint32_t j = 0; j < 32; ++j) sv[j] = __uint_as_float(v[j]);
                }
                float m16[16], m8[8], m4[4], m2[2];
#pragma unroll
                for (uint32_t j = 0; j < 16; ++j) m16[j] = fmaxf(sv[j], sv[j + 16]);
#pragma unroll
                for (uint32_t j = 0; j < 8; ++j) m8[j] = fmaxf(m16[j], m16[j + 8]);
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j) m4[j] = fmaxf(m8[j], m8[j + 4]);
#pragma unroll
                for (uint32_t j = 0; j < 2; ++j) m2[j] = fmaxf(m4[j], m4[j + 2]);
                if (fmaxf(m2[0], m2[1]) > tau && q_valid) {
                    auto leaf = [&](uint32_t j, float sj) {
                        const uint32_t col = chunk * 32u + j;
                        if (sj > tau && col < rows_here) {
                            if (cnt == kBatchStageSlots) flush();
                            stage[cnt * kBatchM] = nominee_key(sj, row0 + col);
                            ++cnt;
                        }
                    };
#pragma unroll
                    for (uint32_t a = 0; a < 2; ++a) {
                        if (!(m2[a] > tau)) continue;
#pragma unroll
                        for (uint32_t b = 0; b < 2; ++b) {
                            const uint32_t i4 = a + 2 * b;
                            if (!(m4[i4] > tau)) continue;
#pragma unroll
                            for (uint32_t c = 0; c < 2; ++c) {
                                const uint32_t i8 = i4 + 4 * c;
                                if (!(m8[i8] > tau)) continue;
#pragma unroll
                                for (uint32_t d = 0; d < 2; ++d) {
                                    const uint32_t i16 = i8 + 8 * d;
                                    if (!(m16[i16] > tau)) continue;
                                    leaf(i16, sv[i16]);
                                    leaf(i16 + 16, sv[i16 + 16]);
                                }
                            }
                        }
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(&tmem_empty[acc], 0u);
            if (__any_sync(WAXVS_FULL_MASK, cnt >= kBatchStageSlots / 2)) {
                flush();
                if (improved && q_valid && root != WAXVS_KEY_NONE) {
                    atomicMax(p.tau_global + q, orderable_u32(nominee_score(root)));
                    improved = false;
                }
            }
        }
        flush();
        uint64_t *dst = p.heaps + static_cast<size_t>(slice * p.groups + group) * HEAP * kBatchM + tid;
        for (uint32_t i = 0; i < HEAP; ++i) dst[i * kBatchM] = heap[i * kBatchM];
    }
    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 5) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---- exact re-score + proof ------------------------------------------------------------------------------------------
// Exact distance of one row by one warp: the generic-dims code path of scan_ldg_kernel, same order, same bits.
template <int METRIC>
__device__ __forceinline__ float exact_row_distance(const float *q, const float *v, uint32_t dims, float a2,
                                                    float sqrt_a2, int lane) {
    float a0 = 0.f, a1 = 0.f, a2_ = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
    auto acc = [&](float qx, float vx, float &a, float &b) {
        if (METRIC == kL2) { const float dd = __fsub_rn(qx, vx); a = __fmaf_rn(dd, dd, a); }
        else { a = __fmaf_rn(qx, vx, a); if (METRIC == kCosine) b = __fmaf_rn(vx, vx, b); }
    };
    if ((dims % 4u) == 0u) {
        const float4 *v4 = reinterpret_cast<const float4 *>(v);
        const float4 *q4 = reinterpret_cast<const float4 *>(q);
        for (uint32_t c = lane; c < dims / 4u; c += 32u) {
            const float4 x = __ldg(v4 + c), y = __ldg(q4 + c);
            acc(y.x, x.x, a0, b0); acc(y.y, x.y, a1, b1); acc(y.z, x.z, a2_, b2); acc(y.w, x.w, a3, b3);
        }
    } else {
        for (uint32_t base = 4u * lane; base < dims; base += 128u) {
            acc(__ldg(q + base), __ldg(v + base), a0, b0);
            if (base + 1 < dims) acc(__ldg(q + base + 1), __ldg(v + base + 1), a1, b1);
            if (base + 2 < dims) acc(__ldg(q + base + 2), __ldg(v + base + 2), a2_, b2);
            if (base + 3 < dims) acc(__ldg(q + base + 3), __ldg(v + base + 3), a3, b3);
        }
    }
    const float s0 = warp_butterfly_sum(__fadd_rn(__fadd_rn(a0, a1), __fadd_rn(a2_, a3)));
    if (METRIC == kCosine) {
        const float s1 = warp_butterfly_sum(__fadd_rn(__fadd_rn(b0, b1), __fadd_rn(b2, b3)));
        return finish_cos(s0, a2, sqrt_a2, s1);
    }
    return METRIC == kDot ? finish_dot(s0) : finish_l2(s0);
}

// Four rows per warp pass: the re-score kernels are latency-bound (a warp that scores one row at a time has a single
// row's loads in flight), so four independent rows are interleaved.  Per row the operations and their order are exactly
// those of exact_row_distance -- same bits.
template <int METRIC>
__device__ __forceinline__ void exact_row_distance_x4(const float *q, const float *const (&v)[4], uint32_t dims, float a2,
                                                      float sqrt_a2, int lane, float (&d)[4]) {
    if ((dims % 4u) != 0u) {
#pragma unroll
        for (int r = 0; r < 4; ++r) d[r] = exact_row_distance<METRIC>(q, v[r], dims, a2, sqrt_a2, lane);
        return;
    }
    float a[4][4], b[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) { a[r][j] = 0.f; b[r][j] = 0.f; }
    auto acc = [&](float qx, float vx, float &aa, float &bb) {
        if (METRIC == kL2) { const float dd = __fsub_rn(qx, vx); aa = __fmaf_rn(dd, dd, aa); }
        else { aa = __fmaf_rn(qx, vx, aa); if (METRIC == kCosine) bb = __fmaf_rn(vx, vx, bb); }
    };
    const float4 *q4 = reinterpret_cast<const float4 *>(q);
    for (uint32_t c = lane; c < dims / 4u; c += 32u) {
        float4 x[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) x[r] = __ldg(reinterpret_cast<const float4 *>(v[r]) + c);
        const float4 y = __ldg(q4 + c);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            acc(y.x, x[r].x, a[r][0], b[r][0]); acc(y.y, x[r].y, a[r][1], b[r][1]);
            acc(y.z, x[r].z, a[r][2], b[r][2]); acc(y.w, x[r].w, a[r][3], b[r][3]);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const float s0 = warp_butterfly_sum(__fadd_rn(__fadd_rn(a[r][0], a[r][1]), __fadd_rn(a[r][2], a[r][3])));
        if (METRIC == kCosine) {
            const float s1 = warp_butterfly_sum(__fadd_rn(__fadd_rn(b[r][0], b[r][1]), __fadd_rn(b[r][2], b[r][3])));
            d[r] = finish_cos(s0, a2, sqrt_a2, s1);
        } else {
            d[r] = METRIC == kDot ? finish_dot(s0) : finish_l2(s0);
        }
    }
}

__device__ __forceinline__ void block_bitonic_sort(uint64_t *sk, uint32_t pow2) {
    for (uint32_t size = 2; size <= pow2; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = threadIdx.x; i < pow2 / 2; i += blockDim.x) {
                const uint32_t lo = (i / stride) * (2 * stride) + (i % stride), hi = lo + stride;
                const bool asc = ((lo & size) == 0);
                const uint64_t a = sk[lo], b = sk[hi];
                if ((a > b) == asc) { sk[lo] = b; sk[hi] = a; }
            }
            __syncthreads();
        }
    }
}

// ---- small allow-lists: score only the listed rows (O(n_allow), not O(N)) --------------------------------------------
// One warp per listed row, exact distance in the kernels' order; then one CTA sorts the keys.
template <int METRIC>
__global__ void __launch_bounds__(256) gather_score_kernel(const float *corpus, const float *queries, uint32_t dims,
                                                           const uint32_t *rows, uint32_t n, uint64_t *keys_all) {
    const int lane = threadIdx.x & 31;
    const float *query = queries + static_cast<size_t>(blockIdx.y) * dims;     // grid.y = queries of a filtered batch
    uint64_t *keys = keys_all + static_cast<size_t>(blockIdx.y) * n;
    float a2 = 0.0f, sqrt_a2 = 0.0f;
    if (METRIC == kCosine) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        for (uint32_t base = 4u * lane; base < dims; base += 128u) {
            const float x = __ldg(query + base); s0 = __fmaf_rn(x, x, s0);
            if (base + 1 < dims) { const float y = __ldg(query + base + 1); s1 = __fmaf_rn(y, y, s1); }
            if (base + 2 < dims) { const float z = __ldg(query + base + 2); s2 = __fmaf_rn(z, z, s2); }
            if (base + 3 < dims) { const float w = __ldg(query + base + 3); s3 = __fmaf_rn(w, w, s3); }
        }
        a2 = warp_butterfly_sum(__fadd_rn(__fadd_rn(s0, s1), __fadd_rn(s2, s3)));
        sqrt_a2 = __fsqrt_rn(a2);
    }
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i0 < n; i0 += 4u * warps) {
        uint32_t rr[4];
        const float *vp[4];
        float d[4];
#pragma unroll
        for (uint32_t r = 0; r < 4; ++r) {
            const uint32_t i = i0 + r * warps;
            rr[r] = rows[i < n ? i : i0];
            vp[r] = corpus + static_cast<size_t>(rr[r]) * dims;
        }
        exact_row_distance_x4<METRIC>(query, vp, dims, a2, sqrt_a2, lane, d);
        if (lane == 0) {
#pragma unroll
            for (uint32_t r = 0; r < 4; ++r) {
                const uint32_t i = i0 + r * warps;
                if (i < n) keys[i] = finite_f32(d[r]) ? make_key(d[r], rr[r]) : WAXVS_KEY_NONE;
            }
        }
    }
}

__global__ void __launch_bounds__(1024) gather_sort_kernel(const uint64_t *keys_all, uint32_t n, uint32_t pow2, ScanParams p) {
    extern __shared__ uint64_t gsk[];
    const uint64_t *keys = keys_all + static_cast<size_t>(blockIdx.x) * n;     // one CTA per query
    p.out += static_cast<size_t>(blockIdx.x) * p.k;
    for (uint32_t i = threadIdx.x; i < pow2; i += blockDim.x) gsk[i] = (i < n) ? keys[i] : WAXVS_KEY_NONE;
    __syncthreads();
    block_bitonic_sort(gsk, pow2);
    for (uint32_t i = threadIdx.x; i < p.k; i += blockDim.x)
        write_candidate(p, static_cast<int>(i), i < pow2 ? gsk[i] : WAXVS_KEY_NONE);
}

struct FinishParams {
    const float *corpus, *queries;
    uint32_t n_rows, dims, n_queries, groups, slices, kprime, k;
    int metric;
    const uint64_t *heaps;
    const uint32_t *max_norm_bits;
    wax_vs_candidate *out;      // [n_queries][k]
    uint32_t *ok;               // [n_queries] 1 = proven exact, 0 = re-run on the exact path
    const uint64_t *frame_ids;
    uint64_t id_base, row_offset;
    uint32_t pow2_all;          // next pow2 >= slices*kprime
    uint32_t rescore;           // nominees re-scored exactly per query: a power of two in [256, kBatchRescoreMax]
    float eps_rel;              // kTf32Eps or kBf16Eps: |score' - score| <= eps_rel * |q||v|
    float *tau_star;            // [2][n_queries] or nullptr: thresholds for the filter levels, in score' units:
                                // (exact k-th score of the re-scored nominees) - eps * |q| (* max|v|); -inf if none.
                                // [0][q]: eps of a TF32 filter pass, [1][q]: eps of a bf16-shadow filter pass
    uint32_t tau_stride;        // n_queries of the whole batch (distance between the two arrays)
};

// One CTA per query.  Union of the slices' nominee heaps -> best kBatchRescore by score' -> exact re-score ->
// proof.  Rows that were never nominated have score' <= tau_excl = max over slices of that slice's final heap
// root (a slice only ever filtered by its own root or by a root another slice had published); nominated rows
// beyond the first kBatchRescore have score' <= the (kBatchRescore+1)-th nominee.
template <int METRIC>
__global__ void __launch_bounds__(512, 2) batch_finish_kernel(const FinishParams p) {
    extern __shared__ uint64_t fsm[];
    uint64_t *sk = fsm;                 // [pow2_all] nominee keys of every slice
    uint64_t *ek = fsm + p.pow2_all;    // [rescore] exact keys
    __shared__ float s_a2, s_sqrt_a2;
    __shared__ uint32_t s_valid, s_excl;
    const uint32_t q = blockIdx.x, g = q / kBatchM, t = q % kBatchM;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float *qv = p.queries + static_cast<size_t>(q) * p.dims;

    if (threadIdx.x == 0) { s_valid = 0; s_excl = 0; }
    __syncthreads();
    const uint32_t total = p.slices * p.kprime;
    uint32_t cnt = 0;
    for (uint32_t i = threadIdx.x; i < p.pow2_all; i += blockDim.x) {
        uint64_t key = WAXVS_KEY_NONE;
        if (i < total) {
            const uint32_t s = i / p.kprime, e = i % p.kprime;
            key = p.heaps[(static_cast<size_t>(s * p.groups + g) * p.kprime + e) * kBatchM + t];
            if (key != WAXVS_KEY_NONE) {
                ++cnt;
                // node 0 is the slice's root: real only when its heap filled up, i.e. when it could exclude rows
                if (e == 0) atomicMax(&s_excl, orderable_u32(nominee_score(key)));
            }
        }
        sk[i] = key;
    }
    if (cnt) atomicAdd(&s_valid, cnt);
    if (warp == 0) {  // |q|^2 in the kernels' order
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        for (uint32_t base = 4u * lane; base < p.dims; base += 128u) {
            const float x = __ldg(qv + base); s0 = __fmaf_rn(x, x, s0);
            if (base + 1 < p.dims) { const float y = __ldg(qv + base + 1); s1 = __fmaf_rn(y, y, s1); }
            if (base + 2 < p.dims) { const float z = __ldg(qv + base + 2); s2 = __fmaf_rn(z, z, s2); }
            if (base + 3 < p.dims) { const float w = __ldg(qv + base + 3); s3 = __fmaf_rn(w, w, s3); }
        }
        const float a2 = warp_butterfly_sum(__fadd_rn(__fadd_rn(s0, s1), __fadd_rn(s2, s3)));
        if (lane == 0) { s_a2 = a2; s_sqrt_a2 = __fsqrt_rn(a2); }
    }
    __syncthreads();
    block_bitonic_sort(sk, p.pow2_all);   // best nominees first

    const uint32_t n_valid = s_valid;
    const uint32_t kpp = min(p.rescore, n_valid);
    float tau = s_excl ? from_orderable_u32(s_excl) : -INFINITY;                     // never-nominated rows
    if (n_valid > p.rescore) tau = fmaxf(tau, nominee_score(sk[p.rescore]));          // nominated, not re-scored
    const bool excluded_any = (s_excl != 0u) || (n_valid > p.rescore);

    for (uint32_t i = threadIdx.x; i < p.rescore; i += blockDim.x) ek[i] = WAXVS_KEY_NONE;
    __syncthreads();
    const uint32_t nwarps = blockDim.x >> 5;
    for (uint32_t i0 = warp; i0 < kpp; i0 += 4u * nwarps) {      // four nominees per warp pass (warp-uniform bounds)
        uint32_t rows[4];
        const float *vp[4];
        float d[4];
#pragma unroll
        for (uint32_t r = 0; r < 4; ++r) {
            const uint32_t i = i0 + r * nwarps;
            rows[r] = static_cast<uint32_t>(sk[i < kpp ? i : i0]);
            vp[r] = p.corpus + static_cast<size_t>(rows[r]) * p.dims;
        }
        exact_row_distance_x4<METRIC>(qv, vp, p.dims, s_a2, s_sqrt_a2, lane, d);
        if (lane == 0) {
#pragma unroll
            for (uint32_t r = 0; r < 4; ++r) {
                const uint32_t i = i0 + r * nwarps;
                if (i < kpp) ek[i] = finite_f32(d[r]) ? make_key(d[r], rows[r]) : WAXVS_KEY_NONE;
            }
        }
    }
    __syncthreads();
    block_bitonic_sort(ek, p.rescore);

    if (threadIdx.x == 0) {
        uint32_t n_exact = 0;
        while (n_exact < kpp && ek[n_exact] != WAXVS_KEY_NONE) ++n_exact;
        uint32_t ok = 1;
        float tau_star = -INFINITY, tau_star16 = -INFINITY;
        if (n_exact >= p.k) {
            const float dk = from_orderable_u32(static_cast<uint32_t>(ek[p.k - 1] >> 32));
            const float qn = s_sqrt_a2;
            const float scale = METRIC == kCosine ? qn : qn * __uint_as_float(*p.max_norm_bits);
            const float sk_exact = METRIC == kCosine ? (1.0f - dk) * qn : 1.0f - dk;
            // Slack on top of the operand-rounding bound: (a) fp32 accumulation error of the tensor-core pass and of
            // the exact re-score, at most ~dims * 2^-24 * |q||v| each; (b) sk_exact is rebuilt from the ROUNDED
            // distance dk (two roundings near 1.0), so a row excluded by less than that could still tie the k-th
            // result in distance and win on the row index -- a few ulps of max(1, |dk|) cover it.
            const float acc_slack = static_cast<float>(p.dims) * 0x1p-23f * scale;
            const float ulp_slack = 0x1p-21f * fmaxf(1.0f, fabsf(dk)) * (METRIC == kCosine ? qn : 1.0f);
            const float eps = p.eps_rel * scale * 1.01f + acc_slack + ulp_slack + 1e-30f;
            if (excluded_any && (!(sk_exact > tau + eps) || !finite_f32(eps))) ok = 0;
            // Filter level: the true top-k rows all have exact score >= the true k-th score >= sk_exact (the nominees
            // are a subset of the corpus), hence score' >= sk_exact - eps_filter: a pass that collects EVERY row above
            // that fixed threshold misses none of them.  A relative 2^-20 margin absorbs the rounding of this
            // subtraction and of the fp32 products sk_exact was built from.
            const float feps = kTf32Eps * scale * 1.01f + acc_slack + ulp_slack + 1e-30f;
            const float t = sk_exact - feps;
            tau_star = finite_f32(t) ? t - fabsf(t) * 0x1p-20f - 1e-30f : -INFINITY;
            const float feps16 = kBf16Eps * scale * 1.01f + acc_slack + ulp_slack + 1e-30f;
            const float t16 = sk_exact - feps16;
            tau_star16 = finite_f32(t16) ? t16 - fabsf(t16) * 0x1p-20f - 1e-30f : -INFINITY;
        } else if (excluded_any) {
            ok = 0;
        }
        p.ok[q] = ok;
        if (p.tau_star) { p.tau_star[q] = tau_star; p.tau_star[p.tau_stride + q] = tau_star16; }
    }
    ScanParams sp{};
    sp.out = p.out + static_cast<size_t>(q) * p.k;
    sp.frame_ids = p.frame_ids; sp.id_base = p.id_base; sp.row_offset = p.row_offset;
    for (uint32_t i = threadIdx.x; i < p.k; i += blockDim.x)
        write_candidate(sp, static_cast<int>(i), i < p.rescore ? ek[i] : WAXVS_KEY_NONE);
}

// ---- filter level (level 2): exact re-score of EVERY candidate above the fixed threshold, then top-k ------------------
// One warp per (query, candidate): exact distance in the kernels' order (bit-identical to the single-query path).
template <int METRIC>
__global__ void __launch_bounds__(256) filter_rescore_kernel(const float *corpus, const float *queries, uint32_t dims,
                                                             const uint32_t *cand_count, const uint32_t *cand_rows,
                                                             uint32_t cand_cap, uint64_t *keys) {
    const uint32_t q = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const uint32_t n = min(cand_count[q], cand_cap);
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const float *qv = queries + static_cast<size_t>(q) * dims;
    float a2 = 0.0f, sqrt_a2 = 0.0f;
    if (METRIC == kCosine) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        for (uint32_t base = 4u * lane; base < dims; base += 128u) {
            const float x = __ldg(qv + base); s0 = __fmaf_rn(x, x, s0);
            if (base + 1 < dims) { const float y = __ldg(qv + base + 1); s1 = __fmaf_rn(y, y, s1); }
            if (base + 2 < dims) { const float z = __ldg(qv + base + 2); s2 = __fmaf_rn(z, z, s2); }
            if (base + 3 < dims) { const float w = __ldg(qv + base + 3); s3 = __fmaf_rn(w, w, s3); }
        }
        a2 = warp_butterfly_sum(__fadd_rn(__fadd_rn(s0, s1), __fadd_rn(s2, s3)));
        sqrt_a2 = __fsqrt_rn(a2);
    }
    for (uint32_t i0 = i; i0 < n; i0 += 4u * warps) {
        uint32_t rr[4];
        const float *vp[4];
        float d[4];
#pragma unroll
        for (uint32_t r = 0; r < 4; ++r) {
            const uint32_t j = i0 + r * warps;
            rr[r] = cand_rows[static_cast<size_t>(q) * cand_cap + (j < n ? j : i0)];
            vp[r] = corpus + static_cast<size_t>(rr[r]) * dims;
        }
        exact_row_distance_x4<METRIC>(qv, vp, dims, a2, sqrt_a2, lane, d);
        if (lane == 0) {
#pragma unroll
            for (uint32_t r = 0; r < 4; ++r) {
                const uint32_t j = i0 + r * warps;
                if (j < n) keys[static_cast<size_t>(q) * cand_cap + j] = finite_f32(d[r]) ? make_key(d[r], rr[r]) : WAXVS_KEY_NONE;
            }
        }
    }
}

struct FilterSelectParams {
    const uint32_t *cand_count;
    const uint64_t *keys;       // [n_queries][cand_cap]
    uint32_t cand_cap, k;
    wax_vs_candidate *out;      // [n_queries][k]
    uint32_t *ok;               // [n_queries]: 1 = complete (the list did not overflow and holds >= k finite rows)
    const uint64_t *frame_ids;
    uint64_t id_base, row_offset;
};

// One CTA per query: sort its candidates' exact keys, the first k are the answer.
__global__ void __launch_bounds__(1024) filter_select_kernel(const FilterSelectParams p) {
    extern __shared__ uint64_t fsk[];
    const uint32_t q = blockIdx.x;
    const uint32_t total = p.cand_count[q];
    const uint32_t n = min(total, p.cand_cap);
    uint32_t pow2 = 64;
    while (pow2 < n || pow2 < p.k) pow2 <<= 1;
    for (uint32_t i = threadIdx.x; i < pow2; i += blockDim.x)
        fsk[i] = (i < n) ? p.keys[static_cast<size_t>(q) * p.cand_cap + i] : WAXVS_KEY_NONE;
    __syncthreads();
    block_bitonic_sort(fsk, pow2);
    if (threadIdx.x == 0) p.ok[q] = (total <= p.cand_cap && fsk[p.k - 1] != WAXVS_KEY_NONE) ? 1u : 0u;
    ScanParams sp{};
    sp.out = p.out + static_cast<size_t>(q) * p.k;
    sp.frame_ids = p.frame_ids; sp.id_base = p.id_base; sp.row_offset = p.row_offset;
    for (uint32_t i = threadIdx.x; i < p.k; i += blockDim.x) write_candidate(sp, static_cast<int>(i), fsk[i]);
}

}  // namespace waxvs
