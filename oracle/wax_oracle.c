/*
 * wax_oracle.c -- CPU ORACLE (test infrastructure only; see wax_oracle.h for the contract, the
 * reference file:line each function restates, and the "parity unpinned" statement).
 *
 * Build: see oracle/Makefile (gcc -O3 -march=native -ffp-contract=off -fno-math-errno -pthread).
 * -ffp-contract=off matters: every fused multiply-add in this file is an explicit fmaf(), every
 * other a*b+c is two roundings, so the three accumulation modes are exactly what the header says.
 */
#include "wax_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* VectorMetric.swift:32-43                                                                    */
float wax_oracle_score_from_distance(int metric, float d) {
    if (!isfinite(d)) return 0.0f;                     /* guard d.isFinite else { return 0 }   */
    if (metric == WAX_ORACLE_COSINE) return 1.0f - d;  /* case .cosine: return 1 - d            */
    return -d;                                         /* case .dot, .l2: return -d             */
}

/* MetalVectorEngine.swift:842-846 / USearchVectorEngine.swift:331-335                          */
int64_t wax_oracle_clamp_topk(int64_t top_k) {
    if (top_k < 1) return 1;
    if (top_k > WAX_ORACLE_MAX_RESULTS) return WAX_ORACLE_MAX_RESULTS;
    return top_k;
}

/* ------------------------------------------------------------------------------------------ */
/* Row statistics in the three accumulation orders.                                            */
typedef struct { float ab, a2, b2, l2; } stats_f32;

static void stats_seq(const float *a, const float *b, uint32_t n, int metric, stats_f32 *s) {
    float ab = 0.0f, a2 = 0.0f, b2 = 0.0f, l2 = 0.0f;
    if (metric == WAX_ORACLE_L2) {
        for (uint32_t i = 0; i < n; ++i) { float d = a[i] - b[i]; l2 += d * d; }
    } else if (metric == WAX_ORACLE_DOT) {
        for (uint32_t i = 0; i < n; ++i) ab += a[i] * b[i];
    } else {
        for (uint32_t i = 0; i < n; ++i) { ab += a[i] * b[i]; a2 += a[i] * a[i]; b2 += b[i] * b[i]; }
    }
    s->ab = ab; s->a2 = a2; s->b2 = b2; s->l2 = l2;
}

static void stats_f64(const float *a, const float *b, uint32_t n, int metric, double *ab_, double *a2_,
                      double *b2_, double *l2_) {
    double ab = 0, a2 = 0, b2 = 0, l2 = 0;
    for (uint32_t i = 0; i < n; ++i) {
        double x = a[i], y = b[i];
        if (metric == WAX_ORACLE_L2) { double d = x - y; l2 += d * d; }
        else { ab += x * y; a2 += x * x; b2 += y * y; }
    }
    *ab_ = ab; *a2_ = a2; *b2_ = b2; *l2_ = l2;
}

/* 128 interleaved accumulators -> ((a0+a1)+(a2+a3)) per lane -> xor butterfly over 32 lanes.    */
static float tree_reduce128(const float *acc) {
    float t[32];
    for (int l = 0; l < 32; ++l) t[l] = (acc[4 * l] + acc[4 * l + 1]) + (acc[4 * l + 2] + acc[4 * l + 3]);
    /* lane 0 of the xor butterfly (16,8,4,2,1): at every level lane l < off holds t[l] + t[l ^ off]; only those
       lanes feed the next level, so the other half need not be computed. */
    for (int off = 16; off >= 1; off >>= 1)
        for (int l = 0; l < off; ++l) t[l] = t[l] + t[l + off];
    return t[0];
}

static float tree_sumsq(const float *a, uint32_t n) {
    float acc[128];
    memset(acc, 0, sizeof acc);
    uint32_t full = n / 128u * 128u;
    for (uint32_t base = 0; base < full; base += 128)
        for (uint32_t j = 0; j < 128; ++j) acc[j] = fmaf(a[base + j], a[base + j], acc[j]);
    for (uint32_t i = full; i < n; ++i) acc[i - full] = fmaf(a[i], a[i], acc[i - full]);
    return tree_reduce128(acc);
}

static void stats_tree(const float *a, const float *b, uint32_t n, int metric, float a2_pre, stats_f32 *s) {
    float acc0[128], acc1[128];
    memset(acc0, 0, sizeof acc0);
    memset(acc1, 0, sizeof acc1);
    uint32_t full = n / 128u * 128u;
    s->ab = s->b2 = s->l2 = 0.0f;
    s->a2 = a2_pre;
    if (metric == WAX_ORACLE_L2) {
        for (uint32_t base = 0; base < full; base += 128)
            for (uint32_t j = 0; j < 128; ++j) { float d = a[base + j] - b[base + j]; acc0[j] = fmaf(d, d, acc0[j]); }
        for (uint32_t i = full; i < n; ++i) { float d = a[i] - b[i]; acc0[i - full] = fmaf(d, d, acc0[i - full]); }
        s->l2 = tree_reduce128(acc0);
    } else if (metric == WAX_ORACLE_DOT) {
        for (uint32_t base = 0; base < full; base += 128)
            for (uint32_t j = 0; j < 128; ++j) acc0[j] = fmaf(a[base + j], b[base + j], acc0[j]);
        for (uint32_t i = full; i < n; ++i) acc0[i - full] = fmaf(a[i], b[i], acc0[i - full]);
        s->ab = tree_reduce128(acc0);
    } else {
        for (uint32_t base = 0; base < full; base += 128)
            for (uint32_t j = 0; j < 128; ++j) {
                acc0[j] = fmaf(a[base + j], b[base + j], acc0[j]);
                acc1[j] = fmaf(b[base + j], b[base + j], acc1[j]);
            }
        for (uint32_t i = full; i < n; ++i) {
            acc0[i - full] = fmaf(a[i], b[i], acc0[i - full]);
            acc1[i - full] = fmaf(b[i], b[i], acc1[i - full]);
        }
        s->ab = tree_reduce128(acc0);
        s->b2 = tree_reduce128(acc1);
    }
}

/* USearch 2.23.0 index_plugins.hpp, metric_cos_gt / metric_ip_gt / metric_l2sq_gt, in fp32.   */
static float finish_f32(int metric, const stats_f32 *s) {
    float d;
    if (metric == WAX_ORACLE_L2) d = s->l2;
    else if (metric == WAX_ORACLE_DOT) d = 1.0f - s->ab;
    else {
        int az = (s->a2 == 0.0f), bz = (s->b2 == 0.0f);
        if (az && bz) d = 0.0f;                 /* result_if_zero[1][1] = 0 */
        else if (az || bz) d = 1.0f;            /* result_if_zero[0][1] = [1][0] = 1 */
        else d = 1.0f - s->ab / (sqrtf(s->a2) * sqrtf(s->b2));
    }
    return d + 0.0f;                            /* canonicalise -0 to +0 (ordering key) */
}

static float finish_f64(int metric, double ab, double a2, double b2, double l2) {
    double d;
    if (metric == WAX_ORACLE_L2) d = l2;
    else if (metric == WAX_ORACLE_DOT) d = 1.0 - ab;
    else {
        int az = (a2 == 0.0), bz = (b2 == 0.0);
        if (az && bz) d = 0.0;
        else if (az || bz) d = 1.0;
        else d = 1.0 - ab / (sqrt(a2) * sqrt(b2));
    }
    return (float)d + 0.0f;
}

/* Per-query precomputation (|q|^2 in the mode's order), so a scan does not redo it per row.    */
typedef struct { int metric, mode; uint32_t dims; const float *q; float a2_f32; } query_ctx;

static void query_ctx_init(query_ctx *c, int metric, int mode, const float *q, uint32_t dims) {
    c->metric = metric; c->mode = mode; c->dims = dims; c->q = q; c->a2_f32 = 0.0f;
    if (metric == WAX_ORACLE_COSINE && mode == WAX_ORACLE_ACC_F32_TREE) c->a2_f32 = tree_sumsq(q, dims);
}

static float row_distance(const query_ctx *c, const float *row) {
    if (c->mode == WAX_ORACLE_ACC_F64) {
        double ab, a2, b2, l2;
        stats_f64(c->q, row, c->dims, c->metric, &ab, &a2, &b2, &l2);
        return finish_f64(c->metric, ab, a2, b2, l2);
    }
    stats_f32 s;
    if (c->mode == WAX_ORACLE_ACC_F32_TREE) stats_tree(c->q, row, c->dims, c->metric, c->a2_f32, &s);
    else stats_seq(c->q, row, c->dims, c->metric, &s);
    return finish_f32(c->metric, &s);
}

float wax_oracle_distance(int metric, int mode, const float *a, const float *b, uint32_t dims) {
    query_ctx c;
    query_ctx_init(&c, metric, mode, a, dims);
    return row_distance(&c, b);
}

/* CosineDistance.metal:233-328 (SIMD8) / :152-229 (SIMD4): dot and |v|^2 only, |q| assumed 1.  */
float wax_oracle_metal_cosine_distance(const float *q, const float *v, uint32_t dims) {
    float dot = 0.0f, mag2 = 0.0f;
    for (uint32_t i = 0; i < dims; ++i) { dot = fmaf(q[i], v[i], dot); mag2 = fmaf(v[i], v[i], mag2); }
    float mag = sqrtf(mag2);
    float sim = (mag > 1e-6f) ? dot / mag : 0.0f;     /* :324-325 */
    return 1.0f - sim;
}

/* VectorMath.swift:15-33 */
void wax_oracle_normalize_l2(const float *in, float *out, uint32_t n) {
    float s = 0.0f;
    for (uint32_t i = 0; i < n; ++i) s += in[i] * in[i];   /* vDSP_svesq */
    float m = sqrtf(s);
    if (!(m > 0.0f)) { if (out != in) memcpy(out, in, (size_t)n * sizeof(float)); return; }
    float inv = 1.0f / m;                                   /* let inverseMagnitude = 1.0 / magnitude */
    for (uint32_t i = 0; i < n; ++i) out[i] = in[i] * inv;  /* vDSP_vsmul */
}

/* VectorMath.swift:123-127 */
int wax_oracle_is_normalized_l2(const float *v, uint32_t n, float tol) {
    if (n == 0) return 0;
    float s = 0.0f;
    for (uint32_t i = 0; i < n; ++i) s += v[i] * v[i];
    return fabsf(sqrtf(s) - 1.0f) <= tol;
}

/* ------------------------------------------------------------------------------------------ */
/* Bounded max-heap under the total order (distance asc, row asc); root = current worst.        */
typedef struct { float d; uint64_t row; } cand;
typedef struct { cand *h; uint32_t n, cap; } topk_heap;

static inline int cand_less(const cand *x, const cand *y) {  /* x strictly better than y */
    return x->d < y->d || (x->d == y->d && x->row < y->row);
}
static void heap_sift_down(topk_heap *t, uint32_t i) {
    for (;;) {
        uint32_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < t->n && cand_less(&t->h[m], &t->h[l])) m = l;
        if (r < t->n && cand_less(&t->h[m], &t->h[r])) m = r;
        if (m == i) return;
        cand tmp = t->h[i]; t->h[i] = t->h[m]; t->h[m] = tmp; i = m;
    }
}
static void heap_push(topk_heap *t, float d, uint64_t row) {
    if (!isfinite(d)) return;                 /* MetalVectorEngine.swift:597: non-finite dropped */
    cand c = { d, row };
    if (t->n < t->cap) {
        uint32_t i = t->n++;
        t->h[i] = c;
        while (i > 0) {
            uint32_t p = (i - 1) / 2;
            if (!cand_less(&t->h[p], &t->h[i])) break;
            cand tmp = t->h[i]; t->h[i] = t->h[p]; t->h[p] = tmp; i = p;
        }
    } else if (cand_less(&c, &t->h[0])) {
        t->h[0] = c;
        heap_sift_down(t, 0);
    }
}
static int cand_cmp(const void *a, const void *b) {
    const cand *x = (const cand *)a, *y = (const cand *)b;
    return cand_less(x, y) ? -1 : (cand_less(y, x) ? 1 : 0);
}

/* ------------------------------------------------------------------------------------------ */
/* Synthetic rows (RAGBenchmarkSupport.swift:130-156)                                           */
static uint64_t fnv1a64_16(uint64_t a, uint64_t b) {
    uint64_t h = 14695981039346656037ull;
    for (int i = 0; i < 8; ++i) { h ^= (a >> (8 * i)) & 0xff; h *= 1099511628211ull; }
    for (int i = 0; i < 8; ++i) { h ^= (b >> (8 * i)) & 0xff; h *= 1099511628211ull; }
    /* fnv1a over near-identical 16-byte keys leaves consecutive rows in arithmetic progression and an
       LCG preserves that, so finish with the splitmix64 avalanche (the reference hashes distinct
       document texts, which are already well separated). */
    h ^= h >> 30; h *= 0xbf58476d1ce4e5b9ull;
    h ^= h >> 27; h *= 0x94d049bb133111ebull;
    h ^= h >> 31;
    return h;
}

void wax_oracle_synth_row(uint64_t seed, uint64_t row, uint32_t dims, int normalize, float *out) {
    uint64_t state = fnv1a64_16(seed, row);
    float s = 0.0f;
    for (uint32_t i = 0; i < dims; ++i) {
        state = state * 6364136223846793005ull + 1442695040888963407ull;
        float x = (float)(int64_t)state * 0x1p-63f;    /* Float(signed) / Float(Int64.max) == *2^-63 */
        out[i] = x;
        s = fmaf(x, x, s);
    }
    if (normalize && s > 0.0f) {
        float inv = 1.0f / sqrtf(s);
        for (uint32_t i = 0; i < dims; ++i) out[i] *= inv;
    }
}

typedef struct { uint64_t seed, first, n; uint32_t dims; int normalize; float *out; } synth_job;
static void *synth_worker(void *p) {
    synth_job *j = (synth_job *)p;
    for (uint64_t r = 0; r < j->n; ++r)
        wax_oracle_synth_row(j->seed, j->first + r, j->dims, j->normalize, j->out + r * (uint64_t)j->dims);
    return NULL;
}
void wax_oracle_synth_rows(uint64_t seed, uint64_t first_row, uint64_t n_rows, uint32_t dims,
                           int normalize, int threads, float *out) {
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n_rows) threads = n_rows ? (int)n_rows : 1;
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    synth_job *jobs = (synth_job *)malloc(sizeof(synth_job) * (size_t)threads);
    uint64_t per = (n_rows + (uint64_t)threads - 1) / (uint64_t)threads;
    for (int t = 0; t < threads; ++t) {
        uint64_t lo = per * (uint64_t)t, hi = lo + per; if (lo > n_rows) lo = n_rows; if (hi > n_rows) hi = n_rows;
        jobs[t] = (synth_job){ seed, first_row + lo, hi - lo, dims, normalize, out + lo * (uint64_t)dims };
        if (threads == 1) synth_worker(&jobs[t]); else pthread_create(&tid[t], NULL, synth_worker, &jobs[t]);
    }
    if (threads > 1) for (int t = 0; t < threads; ++t) pthread_join(tid[t], NULL);
    free(tid); free(jobs);
}

/* ------------------------------------------------------------------------------------------ */
/* Scan                                                                                        */
typedef struct {
    const query_ctx *qc;
    const float *corpus;            /* NULL -> synthetic */
    uint64_t seed; int normalize;
    uint64_t lo, hi;                /* local row range */
    uint64_t row_base;              /* reported row = local + row_base (also synth row id) */
    topk_heap heap;
} scan_job;

static void *scan_worker(void *p) {
    scan_job *j = (scan_job *)p;
    uint32_t dims = j->qc->dims;
    float *tmp = NULL;
    if (!j->corpus) tmp = (float *)malloc(sizeof(float) * (size_t)dims);
    for (uint64_t r = j->lo; r < j->hi; ++r) {
        const float *row;
        if (j->corpus) row = j->corpus + r * (uint64_t)dims;
        else { wax_oracle_synth_row(j->seed, j->row_base + r, dims, j->normalize, tmp); row = tmp; }
        heap_push(&j->heap, row_distance(j->qc, row), r + j->row_base);
    }
    free(tmp);
    return NULL;
}

static int scan_common(int metric, int mode, const float *corpus, uint64_t seed, int normalize,
                       uint64_t n_rows, uint32_t dims, const float *query, int64_t top_k,
                       uint64_t row_base, int threads, uint64_t *out_rows, float *out_d, float *out_s,
                       uint32_t *out_n) {
    if (!query || !out_n || dims == 0 || dims > WAX_ORACLE_MAX_DIMS) return -1;
    if (metric < 0 || metric > 2 || mode < 0 || mode > 2) return -1;
    *out_n = 0;
    if (n_rows == 0) return 0;                       /* guard vectorCount > 0 else { return [] } */
    uint32_t k = (uint32_t)wax_oracle_clamp_topk(top_k);
    if ((uint64_t)k > n_rows) k = (uint32_t)n_rows;  /* topKCount = min(limit, vectorCount) */
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n_rows) threads = (int)n_rows;

    query_ctx qc;
    query_ctx_init(&qc, metric, mode, query, dims);
    scan_job *jobs = (scan_job *)calloc((size_t)threads, sizeof(scan_job));
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    uint64_t per = (n_rows + (uint64_t)threads - 1) / (uint64_t)threads;
    for (int t = 0; t < threads; ++t) {
        uint64_t lo = per * (uint64_t)t, hi = lo + per; if (lo > n_rows) lo = n_rows; if (hi > n_rows) hi = n_rows;
        jobs[t].qc = &qc; jobs[t].corpus = corpus; jobs[t].seed = seed; jobs[t].normalize = normalize;
        jobs[t].lo = lo; jobs[t].hi = hi; jobs[t].row_base = row_base;
        jobs[t].heap.h = (cand *)malloc(sizeof(cand) * k); jobs[t].heap.n = 0; jobs[t].heap.cap = k;
        if (threads == 1) scan_worker(&jobs[t]); else pthread_create(&tid[t], NULL, scan_worker, &jobs[t]);
    }
    if (threads > 1) for (int t = 0; t < threads; ++t) pthread_join(tid[t], NULL);

    size_t total = 0;
    for (int t = 0; t < threads; ++t) total += jobs[t].heap.n;
    cand *all = (cand *)malloc(sizeof(cand) * (total ? total : 1));
    size_t w = 0;
    for (int t = 0; t < threads; ++t) { memcpy(all + w, jobs[t].heap.h, sizeof(cand) * jobs[t].heap.n); w += jobs[t].heap.n; free(jobs[t].heap.h); }
    qsort(all, total, sizeof(cand), cand_cmp);
    uint32_t n = (uint32_t)(total < k ? total : k);
    for (uint32_t i = 0; i < n; ++i) {
        if (out_rows) out_rows[i] = all[i].row;
        if (out_d) out_d[i] = all[i].d;
        if (out_s) out_s[i] = wax_oracle_score_from_distance(metric, all[i].d);
    }
    *out_n = n;
    free(all); free(jobs); free(tid);
    return 0;
}

int wax_oracle_search(int metric, int mode, const float *corpus, uint64_t n_rows, uint32_t dims,
                      const float *query, int64_t top_k, uint64_t row_base, int threads,
                      uint64_t *out_rows, float *out_distances, float *out_scores, uint32_t *out_n) {
    if (!corpus && n_rows) return -1;
    return scan_common(metric, mode, corpus, 0, 0, n_rows, dims, query, top_k, row_base, threads,
                       out_rows, out_distances, out_scores, out_n);
}

int wax_oracle_search_synth(int metric, int mode, uint64_t seed, uint64_t first_row, uint64_t n_rows,
                            uint32_t dims, int normalize, const float *query, int64_t top_k,
                            int threads, uint64_t *out_rows, float *out_distances,
                            float *out_scores, uint32_t *out_n) {
    return scan_common(metric, mode, NULL, seed, normalize, n_rows, dims, query, top_k, first_row,
                       threads, out_rows, out_distances, out_scores, out_n);
}

/* ------------------------------------------------------------------------------------------ */
/* Multi-query scan: every row (generated once when the corpus is synthetic) is scored against all  */
/* n_queries queries; one heap per (thread, query).  Same arithmetic and total order as scan_common */
/* -- it exists so that full-size batched parity tests pay the row generator once, not per query.   */
typedef struct {
    const query_ctx *qcs; uint32_t n_queries;
    const float *corpus; uint64_t seed; int normalize;
    uint64_t lo, hi, row_base;
    topk_heap *heaps;               /* [n_queries] */
} multi_job;

static void *multi_worker(void *p) {
    multi_job *j = (multi_job *)p;
    uint32_t dims = j->qcs[0].dims;
    float *tmp = NULL;
    if (!j->corpus) tmp = (float *)malloc(sizeof(float) * (size_t)dims);
    for (uint64_t r = j->lo; r < j->hi; ++r) {
        const float *row;
        if (j->corpus) row = j->corpus + r * (uint64_t)dims;
        else { wax_oracle_synth_row(j->seed, j->row_base + r, dims, j->normalize, tmp); row = tmp; }
        for (uint32_t q = 0; q < j->n_queries; ++q)
            heap_push(&j->heaps[q], row_distance(&j->qcs[q], row), r + j->row_base);
    }
    free(tmp);
    return NULL;
}

static int multi_common(int metric, int mode, const float *corpus, uint64_t seed, int normalize, uint64_t n_rows,
                        uint32_t dims, const float *queries, uint32_t n_queries, int64_t top_k, uint64_t row_base,
                        int threads, uint64_t *out_rows, float *out_d, float *out_s, uint32_t *out_n) {
    if (!queries || !out_n || dims == 0 || dims > WAX_ORACLE_MAX_DIMS || n_queries == 0) return -1;
    if (metric < 0 || metric > 2 || mode < 0 || mode > 2) return -1;
    for (uint32_t q = 0; q < n_queries; ++q) out_n[q] = 0;
    if (n_rows == 0) return 0;
    uint32_t k = (uint32_t)wax_oracle_clamp_topk(top_k);
    if ((uint64_t)k > n_rows) k = (uint32_t)n_rows;
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n_rows) threads = (int)n_rows;

    query_ctx *qcs = (query_ctx *)malloc(sizeof(query_ctx) * n_queries);
    for (uint32_t q = 0; q < n_queries; ++q) query_ctx_init(&qcs[q], metric, mode, queries + (size_t)q * dims, dims);
    multi_job *jobs = (multi_job *)calloc((size_t)threads, sizeof(multi_job));
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    uint64_t per = (n_rows + (uint64_t)threads - 1) / (uint64_t)threads;
    for (int t = 0; t < threads; ++t) {
        uint64_t lo = per * (uint64_t)t, hi = lo + per; if (lo > n_rows) lo = n_rows; if (hi > n_rows) hi = n_rows;
        jobs[t].qcs = qcs; jobs[t].n_queries = n_queries; jobs[t].corpus = corpus; jobs[t].seed = seed;
        jobs[t].normalize = normalize; jobs[t].lo = lo; jobs[t].hi = hi; jobs[t].row_base = row_base;
        jobs[t].heaps = (topk_heap *)malloc(sizeof(topk_heap) * n_queries);
        for (uint32_t q = 0; q < n_queries; ++q) {
            jobs[t].heaps[q].h = (cand *)malloc(sizeof(cand) * k); jobs[t].heaps[q].n = 0; jobs[t].heaps[q].cap = k;
        }
        if (threads == 1) multi_worker(&jobs[t]); else pthread_create(&tid[t], NULL, multi_worker, &jobs[t]);
    }
    if (threads > 1) for (int t = 0; t < threads; ++t) pthread_join(tid[t], NULL);

    cand *all = (cand *)malloc(sizeof(cand) * (size_t)threads * k);
    for (uint32_t q = 0; q < n_queries; ++q) {
        size_t total = 0;
        for (int t = 0; t < threads; ++t) {
            memcpy(all + total, jobs[t].heaps[q].h, sizeof(cand) * jobs[t].heaps[q].n);
            total += jobs[t].heaps[q].n;
        }
        qsort(all, total, sizeof(cand), cand_cmp);
        uint32_t n = (uint32_t)(total < k ? total : k);
        for (uint32_t i = 0; i < n; ++i) {
            if (out_rows) out_rows[(size_t)q * k + i] = all[i].row;
            if (out_d) out_d[(size_t)q * k + i] = all[i].d;
            if (out_s) out_s[(size_t)q * k + i] = wax_oracle_score_from_distance(metric, all[i].d);
        }
        out_n[q] = n;
    }
    for (int t = 0; t < threads; ++t) {
        for (uint32_t q = 0; q < n_queries; ++q) free(jobs[t].heaps[q].h);
        free(jobs[t].heaps);
    }
    free(all); free(jobs); free(tid); free(qcs);
    return 0;
}

int wax_oracle_search_multi(int metric, int mode, const float *corpus, uint64_t n_rows, uint32_t dims,
                            const float *queries, uint32_t n_queries, int64_t top_k, uint64_t row_base, int threads,
                            uint64_t *out_rows, float *out_distances, float *out_scores, uint32_t *out_n) {
    if (!corpus && n_rows) return -1;
    return multi_common(metric, mode, corpus, 0, 0, n_rows, dims, queries, n_queries, top_k, row_base, threads,
                        out_rows, out_distances, out_scores, out_n);
}

int wax_oracle_search_synth_multi(int metric, int mode, uint64_t seed, uint64_t first_row, uint64_t n_rows,
                                  uint32_t dims, int normalize, const float *queries, uint32_t n_queries,
                                  int64_t top_k, int threads, uint64_t *out_rows, float *out_distances,
                                  float *out_scores, uint32_t *out_n) {
    return multi_common(metric, mode, NULL, seed, normalize, n_rows, dims, queries, n_queries, top_k, first_row,
                        threads, out_rows, out_distances, out_scores, out_n);
}

/* siftDown of MetalVectorEngine.swift:635-647 (distance-only max-heap). */
static void metal_sift_down(cand *h, uint32_t start, uint32_t end) {
    uint32_t root = start;
    for (;;) {
        uint32_t child = root * 2 + 1;
        if (child > end) break;
        uint32_t sw = root;
        if (h[sw].d < h[child].d) sw = child;
        if (child + 1 <= end && h[sw].d < h[child + 1].d) sw = child + 1;
        if (sw == root) return;
        cand tmp = h[root]; h[root] = h[sw]; h[sw] = tmp;
        root = sw;
    }
}

/* MetalVectorEngine.swift:630-680.  Max-heap on distance ONLY; a later row with an equal distance
   never displaces (`value >= heap[0].0 -> continue`, :671). */
uint32_t wax_oracle_metal_cpu_topk(const float *distances, uint64_t count, uint32_t k,
                                   uint64_t *out_rows, float *out_distances) {
    if (k == 0 || count == 0) return 0;
    uint32_t initial = (uint32_t)((uint64_t)k < count ? k : count);
    cand *h = (cand *)malloc(sizeof(cand) * initial);
    for (uint32_t i = 0; i < initial; ++i) { h[i].d = distances[i]; h[i].row = i; }
    for (int64_t i = initial / 2; i >= 0; --i) metal_sift_down(h, (uint32_t)i, initial - 1);
    for (uint64_t i = initial; i < count; ++i) {
        float v = distances[i];
        if (v >= h[0].d) continue;                 /* :671 */
        h[0].d = v; h[0].row = i;
        metal_sift_down(h, 0u, initial - 1);
    }
    qsort(h, initial, sizeof(cand), cand_cmp);
    for (uint32_t i = 0; i < initial; ++i) { out_rows[i] = h[i].row; out_distances[i] = h[i].d; }
    free(h);
    return initial;
}

/* ------------------------------------------------------------------------------------------ */
/* MV2V encoding = 2 (little-endian throughout; this oracle assumes a little-endian host).      */
uint64_t wax_oracle_mv2v_length(uint32_t dims, uint64_t count) {
    return 36ull + count * (uint64_t)dims * 4ull + 8ull + count * 8ull;
}

int wax_oracle_mv2v_encode(uint8_t similarity, uint32_t dims, uint64_t count, const float *vectors,
                           const uint64_t *frame_ids, uint8_t *dst, uint64_t cap, uint64_t *out_len) {
    uint64_t need = wax_oracle_mv2v_length(dims, count);
    if (out_len) *out_len = need;
    if (!dst || cap < need) return -1;
    uint8_t *p = dst;
    const uint8_t magic[4] = { 0x4D, 0x56, 0x32, 0x56 };            /* "MV2V"  :686 */
    memcpy(p, magic, 4); p += 4;
    uint16_t ver = 1; memcpy(p, &ver, 2); p += 2;                   /* :687-688 */
    *p++ = 2;                                                       /* encoding :689 */
    *p++ = similarity;                                              /* :690 */
    memcpy(p, &dims, 4); p += 4;                                    /* :691-692 */
    memcpy(p, &count, 8); p += 8;                                   /* :693-694 */
    uint64_t vbytes = count * (uint64_t)dims * 4ull;
    memcpy(p, &vbytes, 8); p += 8;                                  /* :697-699 */
    memset(p, 0, 8); p += 8;                                        /* reserved :700 */
    if (vbytes) memcpy(p, vectors, vbytes);
    p += vbytes;                                                    /* :703-705 */
    uint64_t ibytes = count * 8ull;
    memcpy(p, &ibytes, 8); p += 8;                                  /* :707-709 */
    if (ibytes) memcpy(p, frame_ids, ibytes);                       /* :710 */
    return 0;
}

int wax_oracle_mv2v_decode(const uint8_t *src, uint64_t len, uint8_t expect_similarity,
                           uint32_t expect_dims, uint64_t *out_count, const uint8_t **out_vectors,
                           const uint8_t **out_ids) {
    if (!src || len < 36) return -1;                                /* :718 */
    const uint8_t magic[4] = { 0x4D, 0x56, 0x32, 0x56 };
    if (memcmp(src, magic, 4) != 0) return -2;                      /* :727 */
    uint16_t ver; memcpy(&ver, src + 4, 2);
    if (ver != 1) return -3;                                        /* :736 */
    if (src[6] != 2) return -4;                                     /* :743 */
    if (src[7] > 2 || src[7] != expect_similarity) return -5;       /* :750-753 */
    uint32_t dims; memcpy(&dims, src + 8, 4);
    if (dims != expect_dims) return -6;                             /* :760 */
    uint64_t count, vbytes; memcpy(&count, src + 12, 8); memcpy(&vbytes, src + 20, 8);
    for (int i = 0; i < 8; ++i) if (src[28 + i] != 0) return -7;    /* :778 */
    if (count > UINT64_MAX / 4 / (dims ? dims : 1)) return -8;
    if (vbytes != count * (uint64_t)dims * 4ull) return -8;         /* :782 */
    if (len < 36 + vbytes + 8 || 36 + vbytes + 8 < vbytes) return -9; /* :785 */
    uint64_t ibytes; memcpy(&ibytes, src + 36 + vbytes, 8);
    if (ibytes != count * 8ull) return -10;                         /* :806 */
    if (len != 36 + vbytes + 8 + ibytes) return -11;                /* VectorSerializer.swift:141-144 */
    if (out_count) *out_count = count;
    if (out_vectors) *out_vectors = src + 36;
    if (out_ids) *out_ids = src + 36 + vbytes + 8;
    return 0;
}
