/*
 * jv_cpu_simd.c — tuned CPU arm of the benchmark: what the reference's Panama-Vector-API path does on the host cores.
 *
 * THIS IS BENCH / TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rule as jv_oracle.c: only tests/, smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it).  It is NOT the bit-exact checker — that stays jv_oracle.c, whose
 * canonical 128-partial reductions and scalar gather loop exist for reproducibility, not speed.  This file is the same
 * algorithm (fp32 ADC table, best-first GraphSearcher loop with a candidate max-heap and a bounded
 * result min-heap, exact rerank of the approximate list, tie -> lower id) written the way jVector's SIMD provider
 * (PanamaVectorUtilSupport: lane-wise partial sums, gathers in assembleAndSum) runs it: AVX-512 (or AVX2) gathers over
 * the table with 16 (8) subspaces per instruction, 4 independent FMA accumulators in the exact scorers, native-width
 * reductions (free summation order), software prefetch of the next code rows, an epoch-stamped visited array instead of a
 * hash set.  The instruction set is picked at run time (__builtin_cpu_supports), so one binary runs on the build
 * container and on the GPU box.  Gate: recall equal to the checker's within sampling error (tests/test_cpu_simd.py).
 *
 * Scope: unfiltered queries, threshold 0, rerank floor 0, similarities EUCLIDEAN / DOT / COSINE / MIP, with or without PQ
 * (JVectorReader.java:130-210 with the default collector parameters) — what the benchmark times.
 */
#include <immintrin.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/jvgpu.h"

#define JVS_EXPORT __attribute__((visibility("default")))
#define T512 __attribute__((target("avx512f,avx512bw,avx512vl,avx512dq,fma,avx2")))
#define T256 __attribute__((target("avx2,fma")))

/* ------------------------------------------------------------------------------------------ exact scorers */
T512 static float dot_512(const float *a, const float *b, int n) {
    __m512 s0 = _mm512_setzero_ps(), s1 = s0, s2 = s0, s3 = s0;
    int i = 0;
    for (; i + 64 <= n; i += 64) {
        s0 = _mm512_fmadd_ps(_mm512_loadu_ps(a + i), _mm512_loadu_ps(b + i), s0);
        s1 = _mm512_fmadd_ps(_mm512_loadu_ps(a + i + 16), _mm512_loadu_ps(b + i + 16), s1);
        s2 = _mm512_fmadd_ps(_mm512_loadu_ps(a + i + 32), _mm512_loadu_ps(b + i + 32), s2);
        s3 = _mm512_fmadd_ps(_mm512_loadu_ps(a + i + 48), _mm512_loadu_ps(b + i + 48), s3);
    }
    for (; i + 16 <= n; i += 16) s0 = _mm512_fmadd_ps(_mm512_loadu_ps(a + i), _mm512_loadu_ps(b + i), s0);
    float r = _mm512_reduce_add_ps(_mm512_add_ps(_mm512_add_ps(s0, s1), _mm512_add_ps(s2, s3)));
    for (; i < n; i++) r += a[i] * b[i];
    return r;
}
T512 static float l2_512(const float *a, const float *b, int n) {
    __m512 s0 = _mm512_setzero_ps(), s1 = s0;
    int i = 0;
    for (; i + 32 <= n; i += 32) {
        const __m512 d0 = _mm512_sub_ps(_mm512_loadu_ps(a + i), _mm512_loadu_ps(b + i));
        const __m512 d1 = _mm512_sub_ps(_mm512_loadu_ps(a + i + 16), _mm512_loadu_ps(b + i + 16));
        s0 = _mm512_fmadd_ps(d0, d0, s0);
        s1 = _mm512_fmadd_ps(d1, d1, s1);
    }
    for (; i + 16 <= n; i += 16) {
        const __m512 d0 = _mm512_sub_ps(_mm512_loadu_ps(a + i), _mm512_loadu_ps(b + i));
        s0 = _mm512_fmadd_ps(d0, d0, s0);
    }
    float r = _mm512_reduce_add_ps(_mm512_add_ps(s0, s1));
    for (; i < n; i++) r += (a[i] - b[i]) * (a[i] - b[i]);
    return r;
}
T256 static inline float hsum256(__m256 v) {
    __m128 x = _mm_add_ps(_mm256_castps256_ps128(v), _mm256_extractf128_ps(v, 1));
    x = _mm_add_ps(x, _mm_movehl_ps(x, x));
    x = _mm_add_ss(x, _mm_shuffle_ps(x, x, 1));
    return _mm_cvtss_f32(x);
}
T256 static float dot_256(const float *a, const float *b, int n) {
    __m256 s0 = _mm256_setzero_ps(), s1 = s0, s2 = s0, s3 = s0;
    int i = 0;
    for (; i + 32 <= n; i += 32) {
        s0 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i), s0);
        s1 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i + 8), _mm256_loadu_ps(b + i + 8), s1);
        s2 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i + 16), _mm256_loadu_ps(b + i + 16), s2);
        s3 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i + 24), _mm256_loadu_ps(b + i + 24), s3);
    }
    for (; i + 8 <= n; i += 8) s0 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i), s0);
    float r = hsum256(_mm256_add_ps(_mm256_add_ps(s0, s1), _mm256_add_ps(s2, s3)));
    for (; i < n; i++) r += a[i] * b[i];
    return r;
}
T256 static float l2_256(const float *a, const float *b, int n) {
    __m256 s0 = _mm256_setzero_ps(), s1 = s0;
    int i = 0;
    for (; i + 16 <= n; i += 16) {
        const __m256 d0 = _mm256_sub_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i));
        const __m256 d1 = _mm256_sub_ps(_mm256_loadu_ps(a + i + 8), _mm256_loadu_ps(b + i + 8));
        s0 = _mm256_fmadd_ps(d0, d0, s0);
        s1 = _mm256_fmadd_ps(d1, d1, s1);
    }
    float r = hsum256(_mm256_add_ps(s0, s1));
    for (; i < n; i++) r += (a[i] - b[i]) * (a[i] - b[i]);
    return r;
}

/* ------------------------------------------------------------------------------------------ ADC: assembleAndSum
 * sum_m lut[m*K + code[m]] with K = 256 (every production shape: K = min(256, n), n >= 1024). */
T512 static float adc_512(const float *lut, const uint8_t *code, int M) {
    __m512 acc0 = _mm512_setzero_ps(), acc1 = acc0;
    const __m512i step = _mm512_set1_epi32(16 * 256);
    __m512i base = _mm512_mullo_epi32(_mm512_set_epi32(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0), _mm512_set1_epi32(256));
    int m = 0;
    for (; m + 32 <= M; m += 32) {
        const __m512i c0 = _mm512_cvtepu8_epi32(_mm_loadu_si128((const __m128i *)(code + m)));
        const __m512i c1 = _mm512_cvtepu8_epi32(_mm_loadu_si128((const __m128i *)(code + m + 16)));
        acc0 = _mm512_add_ps(acc0, _mm512_i32gather_ps(_mm512_add_epi32(base, c0), lut, 4));
        base = _mm512_add_epi32(base, step);
        acc1 = _mm512_add_ps(acc1, _mm512_i32gather_ps(_mm512_add_epi32(base, c1), lut, 4));
        base = _mm512_add_epi32(base, step);
    }
    for (; m + 16 <= M; m += 16) {
        const __m512i c0 = _mm512_cvtepu8_epi32(_mm_loadu_si128((const __m128i *)(code + m)));
        acc0 = _mm512_add_ps(acc0, _mm512_i32gather_ps(_mm512_add_epi32(base, c0), lut, 4));
        base = _mm512_add_epi32(base, step);
    }
    float r = _mm512_reduce_add_ps(_mm512_add_ps(acc0, acc1));
    for (; m < M; m++) r += lut[m * 256 + code[m]];
    return r;
}
T256 static float adc_256(const float *lut, const uint8_t *code, int M) {
    __m256 acc0 = _mm256_setzero_ps(), acc1 = acc0;
    const __m256i step = _mm256_set1_epi32(8 * 256);
    __m256i base = _mm256_mullo_epi32(_mm256_set_epi32(7, 6, 5, 4, 3, 2, 1, 0), _mm256_set1_epi32(256));
    int m = 0;
    for (; m + 16 <= M; m += 16) {
        const __m256i c0 = _mm256_cvtepu8_epi32(_mm_loadl_epi64((const __m128i *)(code + m)));
        const __m256i c1 = _mm256_cvtepu8_epi32(_mm_loadl_epi64((const __m128i *)(code + m + 8)));
        acc0 = _mm256_add_ps(acc0, _mm256_i32gather_ps(lut, _mm256_add_epi32(base, c0), 4));
        base = _mm256_add_epi32(base, step);
        acc1 = _mm256_add_ps(acc1, _mm256_i32gather_ps(lut, _mm256_add_epi32(base, c1), 4));
        base = _mm256_add_epi32(base, step);
    }
    float r = hsum256(_mm256_add_ps(acc0, acc1));
    for (; m < M; m++) r += lut[m * 256 + code[m]];
    return r;
}

typedef float (*pair_fn)(const float *, const float *, int);
typedef float (*adc_fn)(const float *, const uint8_t *, int);
static pair_fn g_dot, g_l2;
static adc_fn g_adc;
static int g_isa; /* 512, 256 */

static void pick_isa(void) {
    if (g_isa) return;
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
        __builtin_cpu_supports("avx512dq")) {
        g_dot = dot_512, g_l2 = l2_512, g_adc = adc_512, g_isa = 512;
    } else {
        g_dot = dot_256, g_l2 = l2_256, g_adc = adc_256, g_isa = 256;
    }
}

JVS_EXPORT int32_t jvs_isa(void) {
    pick_isa();
    return g_isa;
}

/* ------------------------------------------------------------------------------------------ heaps of (score, ~id) keys */
static inline uint32_t f2ord(float f) {
    uint32_t b;
    memcpy(&b, &f, 4);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
static inline float ord2f(uint32_t u) {
    uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    float f;
    memcpy(&f, &b, 4);
    return f;
}
static inline uint64_t mk_key(float s, int32_t id) { return ((uint64_t)f2ord(s) << 32) | (uint32_t)(~id); }
static inline int32_t key_id(uint64_t k) { return (int32_t)(~(uint32_t)k); }
static inline float key_score(uint64_t k) { return ord2f((uint32_t)(k >> 32)); }

typedef struct {
    uint64_t *a;
    int n, cap;
} heap_t;
static void heap_reserve(heap_t *h, int need) {
    if (need > h->cap) {
        h->cap = need * 2 + 64;
        h->a = (uint64_t *)realloc(h->a, sizeof(uint64_t) * (size_t)h->cap);
    }
}
static void max_push(heap_t *h, uint64_t k) { /* candidates: best on top */
    heap_reserve(h, h->n + 1);
    int i = h->n++;
    while (i > 0) {
        const int p = (i - 1) >> 1;
        if (h->a[p] >= k) break;
        h->a[i] = h->a[p];
        i = p;
    }
    h->a[i] = k;
}
static uint64_t max_pop(heap_t *h) {
    const uint64_t top = h->a[0], last = h->a[--h->n];
    int i = 0;
    for (;;) {
        int c = 2 * i + 1;
        if (c >= h->n) break;
        if (c + 1 < h->n && h->a[c + 1] > h->a[c]) c++;
        if (h->a[c] <= last) break;
        h->a[i] = h->a[c];
        i = c;
    }
    if (h->n > 0) h->a[i] = last;
    return top;
}
static void min_push_bounded(heap_t *h, int cap, uint64_t k) { /* results: worst on top, at most cap entries */
    if (h->n >= cap) {
        if (k <= h->a[0]) return;
        int i = 0; /* replace the worst */
        for (;;) {
            int c = 2 * i + 1;
            if (c >= h->n) break;
            if (c + 1 < h->n && h->a[c + 1] < h->a[c]) c++;
            if (h->a[c] >= k) break;
            h->a[i] = h->a[c];
            i = c;
        }
        h->a[i] = k;
        return;
    }
    heap_reserve(h, h->n + 1);
    int i = h->n++;
    while (i > 0) {
        const int p = (i - 1) >> 1;
        if (h->a[p] <= k) break;
        h->a[i] = h->a[p];
        i = p;
    }
    h->a[i] = k;
}
static int cmp_desc(const void *x, const void *y) {
    const uint64_t a = *(const uint64_t *)x, b = *(const uint64_t *)y;
    return a < b ? 1 : a > b ? -1 : 0;
}

/* ------------------------------------------------------------------------------------------ per-thread scratch */
typedef struct {
    float *lut, *normlut, *qc;
    uint32_t *stamp; /* visited: stamp[node] == epoch */
    uint32_t epoch;
    heap_t cand, res, fin;
    uint64_t *sorted;
    int sorted_cap;
} scratch_t;

typedef struct jvs_index {
    jv_index_desc d;
    int sub; /* uniform sub-vector size (dim / M) */
    float *cbT;       /* codebooks transposed per subspace: [M][sub][256], so the table build vectorises over the centroids */
    float *node_norm; /* cosine: ||decode(code)||^2 per node */
    float *vec_norm;  /* cosine: ||x||^2 per node */
} jvs_index;

JVS_EXPORT jvs_index *jvs_create(const jv_index_desc *desc) {
    pick_isa();
    if (desc->pq_m > 0 && (desc->pq_k != 256 || desc->dim % desc->pq_m != 0)) return NULL; /* tuned for the production shapes */
    jvs_index *ix = (jvs_index *)calloc(1, sizeof(*ix));
    ix->d = *desc;
    ix->sub = desc->pq_m > 0 ? desc->dim / desc->pq_m : 0;
    const int64_t n = desc->n;
    const int dim = desc->dim, M = desc->pq_m;
    if (M > 0) {
        ix->cbT = (float *)aligned_alloc(64, sizeof(float) * (size_t)M * ix->sub * 256);
        for (int m = 0; m < M; m++)
            for (int c = 0; c < 256; c++)
                for (int j = 0; j < ix->sub; j++) ix->cbT[((size_t)m * ix->sub + j) * 256 + c] = desc->pq_codebooks[((size_t)m * 256 + c) * ix->sub + j];
    }
    if (desc->similarity == JV_SIM_COSINE) {
        ix->vec_norm = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; i++) ix->vec_norm[i] = g_dot(desc->vectors + i * dim, desc->vectors + i * dim, dim);
        if (M > 0) {
            float *cn = (float *)malloc(sizeof(float) * (size_t)M * 256); /* ||C_m[c]||^2 */
            for (int m = 0; m < M; m++)
                for (int c = 0; c < 256; c++) {
                    const float *cv = desc->pq_codebooks + ((size_t)m * 256 + c) * ix->sub;
                    float s = 0.f;
                    for (int j = 0; j < ix->sub; j++) s += cv[j] * cv[j];
                    cn[m * 256 + c] = s;
                }
            ix->node_norm = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < n; i++) ix->node_norm[i] = g_adc(cn, desc->pq_codes + i * M, M);
            free(cn);
        }
    }
    return ix;
}

JVS_EXPORT void jvs_destroy(jvs_index *ix) {
    if (!ix) return;
    free(ix->cbT);
    free(ix->node_norm);
    free(ix->vec_norm);
    free(ix);
}

/* K1: fp32 ADC table, PQVectors.precomputedScoreFunctionFor (JVectorReader.java:354) */
static void build_lut(const jvs_index *ix, scratch_t *S, const float *q) {
    const jv_index_desc *d = &ix->d;
    const int M = d->pq_m, sub = ix->sub;
    const float *qq = q;
    if (d->similarity == JV_SIM_EUCLIDEAN && d->pq_global_centroid) {
        for (int i = 0; i < d->dim; i++) S->qc[i] = q[i] - d->pq_global_centroid[i];
        qq = S->qc;
    }
    const int l2 = d->similarity == JV_SIM_EUCLIDEAN;
    for (int m = 0; m < M; m++) { /* inner loops run over the 256 centroids: unit stride, vectorised by the compiler */
        const float *cb = ix->cbT + (size_t)m * sub * 256;
        const float *qs = qq + m * sub;
        float *restrict out = S->lut + m * 256;
        for (int c = 0; c < 256; c++) out[c] = 0.f;
        for (int j = 0; j < sub; j++) {
            const float qj = qs[j];
            const float *restrict col = cb + (size_t)j * 256;
            if (l2) {
                for (int c = 0; c < 256; c++) {
                    const float t = qj - col[c];
                    out[c] += t * t;
                }
            } else {
                for (int c = 0; c < 256; c++) out[c] += qj * col[c];
            }
        }
    }
}

static inline float finish(int sim, float raw, float qnorm, float xnorm) { /* SURVEY A.2 */
    if (sim == JV_SIM_EUCLIDEAN) return 1.0f / (1.0f + raw);
    if (sim == JV_SIM_COSINE) return (1.0f + raw / sqrtf(qnorm * xnorm)) * 0.5f;
    return (1.0f + raw) * 0.5f;
}

static void search_one(const jvs_index *ix, scratch_t *S, const float *q, int k, int rerank_k, int32_t *out_doc, float *out_score,
                       int32_t *out_count, jv_query_stats *st) {
    const jv_index_desc *d = &ix->d;
    const int sim = d->similarity, dim = d->dim, R = d->max_degree, M = d->pq_m;
    const int use_pq = M > 0, l2 = sim == JV_SIM_EUCLIDEAN;
    const float qnorm = sim == JV_SIM_COSINE ? g_dot(q, q, dim) : 0.f;
    const float mip_mul = (sim == JV_SIM_MIP && !use_pq) ? 2.0f : 1.0f; /* wrapExactScoreFunction, JVectorReader.java:220-239 */
    if (use_pq) build_lut(ix, S, q);
    if (++S->epoch == 0) {
        memset(S->stamp, 0, sizeof(uint32_t) * (size_t)d->n);
        S->epoch = 1;
    }
#define EXACT(node)                                                                                                                    \
    finish(sim, l2 ? g_l2(q, d->vectors + (int64_t)(node)*dim, dim) : g_dot(q, d->vectors + (int64_t)(node)*dim, dim), qnorm,           \
           sim == JV_SIM_COSINE ? ix->vec_norm[node] : 0.f)
#define APPROX(node)                                                                                                                   \
    (use_pq ? finish(sim, g_adc(S->lut, d->pq_codes + (int64_t)(node)*M, M), qnorm, sim == JV_SIM_COSINE ? ix->node_norm[node] : 0.f)   \
            : EXACT(node) * mip_mul)
    S->cand.n = S->res.n = 0;
    int visited = 0, expanded = 0, reranked = 0;
    if (d->n > 0 && d->entry_node >= 0) {
        S->stamp[d->entry_node] = S->epoch;
        visited++;
        max_push(&S->cand, mk_key(APPROX(d->entry_node), d->entry_node));
    }
    while (S->cand.n > 0) {
        const uint64_t top = S->cand.a[0];
        if (S->res.n >= rerank_k && key_score(top) < key_score(S->res.a[0])) break;
        max_pop(&S->cand);
        const int32_t c = key_id(top);
        min_push_bounded(&S->res, rerank_k, top);
        expanded++;
        const int32_t *nb = d->adjacency + (int64_t)c * R;
        int32_t fresh[128];
        int nf = 0;
        for (int j = 0; j < R; j++) { /* visited test first, so the code rows of all fresh neighbours can be prefetched */
            const int32_t nn = nb[j];
            if (nn < 0) break;
            if (S->stamp[nn] == S->epoch) continue;
            S->stamp[nn] = S->epoch;
            fresh[nf++] = nn;
            if (use_pq) {
                const char *row = (const char *)(d->pq_codes + (int64_t)nn * M);
                _mm_prefetch(row, _MM_HINT_T0);
                if (M > 64) _mm_prefetch(row + 64, _MM_HINT_T0);
                if (M > 128) _mm_prefetch(row + 128, _MM_HINT_T0);
            } else {
                _mm_prefetch((const char *)(d->vectors + (int64_t)nn * dim), _MM_HINT_T0);
            }
        }
        visited += nf;
        for (int j = 0; j < nf; j++) max_push(&S->cand, mk_key(APPROX(fresh[j]), fresh[j]));
    }
    const int na = S->res.n;
    if (na > S->sorted_cap) {
        S->sorted_cap = na * 2;
        S->sorted = (uint64_t *)realloc(S->sorted, sizeof(uint64_t) * (size_t)S->sorted_cap);
    }
    memcpy(S->sorted, S->res.a, sizeof(uint64_t) * (size_t)na);
    qsort(S->sorted, (size_t)na, sizeof(uint64_t), cmp_desc);
    S->fin.n = 0;
    if (use_pq) /* the reranker reads the inline vectors: start all the gathers before the first dot product */
        for (int i = 0; i < na; i++) _mm_prefetch((const char *)(d->vectors + (int64_t)key_id(S->sorted[i]) * dim), _MM_HINT_T1);
    for (int i = 0; i < na; i++) {
        const int32_t node = key_id(S->sorted[i]);
        float s = key_score(S->sorted[i]);
        if (use_pq) {
            s = EXACT(node); /* reranker is NOT x2-wrapped (JVectorReader.java:353-356) */
            reranked++;
        }
        const int32_t doc = d->ord_to_doc ? d->ord_to_doc[node] : node;
        if (doc >= 0) min_push_bounded(&S->fin, k, mk_key(s, doc));
    }
#undef APPROX
#undef EXACT
    qsort(S->fin.a, (size_t)S->fin.n, sizeof(uint64_t), cmp_desc);
    for (int i = 0; i < k; i++) {
        out_doc[i] = i < S->fin.n ? key_id(S->fin.a[i]) : -1;
        out_score[i] = i < S->fin.n ? key_score(S->fin.a[i]) : 0.0f;
    }
    if (out_count) *out_count = S->fin.n;
    if (st) {
        st->visited = visited;
        st->expanded = expanded;
        st->expanded_base = expanded;
        st->reranked = reranked;
    }
}

/* one query per thread (JVectorReader.search is single-threaded per query); threads <= 0: all the cores the process may use */
JVS_EXPORT int32_t jvs_search_batch(const jvs_index *ix, const float *queries, int32_t nq, int32_t k, int32_t rerank_k, int32_t *out_doc,
                                    float *out_score, int32_t *out_count, jv_query_stats *stats, int32_t threads) {
    if (!ix || !queries || k < 1 || rerank_k < k || ix->d.max_degree > 128) return -1;
    const jv_index_desc *d = &ix->d;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_num_procs();
#else
    threads = 1;
#endif
#pragma omp parallel num_threads(threads)
    {
        scratch_t S;
        memset(&S, 0, sizeof(S));
        S.lut = (float *)aligned_alloc(64, sizeof(float) * (size_t)(d->pq_m > 0 ? d->pq_m : 1) * 256);
        S.qc = (float *)malloc(sizeof(float) * (size_t)d->dim);
        S.stamp = (uint32_t *)calloc((size_t)(d->n > 0 ? d->n : 1), sizeof(uint32_t));
#pragma omp for schedule(dynamic, 4)
        for (int32_t i = 0; i < nq; i++)
            search_one(ix, &S, queries + (int64_t)i * d->dim, k, rerank_k, out_doc + (int64_t)i * k, out_score + (int64_t)i * k,
                       out_count ? out_count + i : NULL, stats ? stats + i : NULL);
        free(S.lut);
        free(S.qc);
        free(S.stamp);
        free(S.cand.a);
        free(S.res.a);
        free(S.fin.a);
        free(S.sorted);
    }
    return 0;
}

JVS_EXPORT int32_t jvs_num_procs(void) {
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}
