/*
 * jv_oracle.c — CPU restatement of the opensearch-jvector query hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (opensearch-jvector_b200/) never links, imports or calls anything in oracle/.
 *
 * PARITY STATUS: the arithmetic of this path lives in the third-party library
 * io.github.jbellis:jvector:4.0.0-rc.9 (reference build.gradle:362, gradle.properties:8) whose
 * source is NOT under /root/reference and which cannot be built here (no JVM).  This file restates
 * its published algorithm and the plugin's glue around it.  It is pinned
 * against every known-answer test the reference holds for this path (tests/test_oracle_kat.py:
 * KNNJVectorTests analytic top-3 ids/scores, the CommonTestUtils score formulas, seeded recall
 * floors), but PQ codes / LUT values / traversal order have no golden vectors in the reference:
 * for those, PARITY IS UNPINNED and claims read "vs. CPU restatement of jVector 4.0.0-rc.9".
 *
 * Citations: paths relative to /root/reference/src/main/java/org/opensearch/knn/index/codec/jvector/.
 *
 * Floating point: compiled with -ffp-contract=off; every fused multiply-add is an explicit fmaf()
 * so the GPU kernels (which use __fmaf_rn in the same order) can match bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/jvgpu.h" /* jv_index_desc, jv_query_stats, JV_SIM_* (types only) */

#define JVO_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * java.util.Random — the reference's fixtures are `new Random(seed).nextFloat()` row-major
 * (src/testFixtures/java/org/opensearch/knn/TestUtils.java:108-124).  Documented 48-bit LCG.
 * ------------------------------------------------------------------------------------------ */
JVO_EXPORT void jvo_java_random_floats(int64_t seed, int64_t count, float *out) {
    const uint64_t mask = (1ULL << 48) - 1;
    uint64_t state = ((uint64_t)seed ^ 0x5DEECE66DULL) & mask;
    for (int64_t i = 0; i < count; i++) {
        state = (state * 0x5DEECE66DULL + 0xBULL) & mask;
        int32_t bits24 = (int32_t)(state >> (48 - 24));
        out[i] = (float)bits24 / (float)(1 << 24);
    }
}

/* ------------------------------------------------------------------------------------------
 * Canonical fp32 reductions.  jVector's VectorUtil (Panama) keeps lane-wise partial sums and
 * reduces horizontally, so its exact rounding depends on the host's SIMD width; any fixed order
 * is within dim*eps of it (hence the 1e-5 relative gate).  We fix ONE order that CPU and GPU both
 * implement exactly: 128 strided partial sums (element i -> partial i mod 128, fmaf), then
 * ((p0+p1)+(p2+p3)) inside each group of 4, then a halving tree over the 32 groups.
 * ------------------------------------------------------------------------------------------ */
static inline float reduce128(const float *acc) {
    float v[32];
    for (int j = 0; j < 32; j++) v[j] = (acc[4 * j] + acc[4 * j + 1]) + (acc[4 * j + 2] + acc[4 * j + 3]);
    for (int off = 16; off >= 1; off >>= 1)
        for (int j = 0; j < off; j++) v[j] = v[j] + v[j + off];
    return v[0];
}

static inline float canon_dot(const float *a, const float *b, int dim) {
    float acc[128];
    memset(acc, 0, sizeof(acc));
    int i = 0;
    for (; i + 128 <= dim; i += 128)
        for (int c = 0; c < 128; c++) acc[c] = fmaf(a[i + c], b[i + c], acc[c]);
    for (int c = 0; i + c < dim; c++) acc[c] = fmaf(a[i + c], b[i + c], acc[c]);
    return reduce128(acc);
}

static inline float canon_l2sq(const float *a, const float *b, int dim) {
    float acc[128];
    memset(acc, 0, sizeof(acc));
    int i = 0;
    for (; i + 128 <= dim; i += 128)
        for (int c = 0; c < 128; c++) {
            float d = a[i + c] - b[i + c];
            acc[c] = fmaf(d, d, acc[c]);
        }
    for (int c = 0; i + c < dim; c++) {
        float d = a[i + c] - b[i + c];
        acc[c] = fmaf(d, d, acc[c]);
    }
    return reduce128(acc);
}

/* jVector VectorSimilarityFunction.compare (pinned by CommonTestUtils.java:84-93):
 * EUCLIDEAN 1/(1+d2), DOT (1+dot)/2, COSINE (1+cos)/2.  `qnorm` = canon_dot(q,q). */
static inline float exact_score(int sim, const float *q, float qnorm, const float *x, int dim) {
    switch (sim) {
    case JV_SIM_EUCLIDEAN:
        return 1.0f / (1.0f + canon_l2sq(q, x, dim));
    case JV_SIM_COSINE: {
        float d = canon_dot(q, x, dim);
        float xn = canon_dot(x, x, dim);
        float c = (float)((double)d / sqrt((double)qnorm * (double)xn));
        return (1.0f + c) / 2.0f;
    }
    default: /* DOT, MIP */
        return (1.0f + canon_dot(q, x, dim)) / 2.0f;
    }
}

JVO_EXPORT float jvo_exact_score(int32_t sim, const float *q, const float *x, int32_t dim) {
    return exact_score(sim, q, canon_dot(q, q, dim), x, dim);
}

/* ------------------------------------------------------------------------------------------
 * (score, node) ordering.  jVector NodeQueue packs (floatToSortableInt(score) << 32) | ~node so
 * ties prefer the LOWER node id.  We use the unsigned-orderable transform so plain
 * uint64 comparison works: larger key = better.
 * ------------------------------------------------------------------------------------------ */
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
static inline uint64_t mk_key(float score, int32_t node) { return ((uint64_t)f2ord(score) << 32) | (uint32_t)(~node); }
static inline int32_t key_node(uint64_t k) { return (int32_t)(~(uint32_t)k); }
static inline float key_score(uint64_t k) { return ord2f((uint32_t)(k >> 32)); }

/* growable binary heaps over uint64 keys */
typedef struct {
    uint64_t *a;
    int n, cap;
} heap_t;
static void heap_reserve(heap_t *h, int cap) {
    if (cap > h->cap) {
        h->cap = cap < 64 ? 64 : cap;
        h->a = (uint64_t *)realloc(h->a, sizeof(uint64_t) * (size_t)h->cap);
    }
}
static void maxheap_push(heap_t *h, uint64_t k) {
    if (h->n == h->cap) heap_reserve(h, h->cap * 2 + 64);
    int i = h->n++;
    while (i > 0) {
        int p = (i - 1) >> 1;
        if (h->a[p] >= k) break;
        h->a[i] = h->a[p];
        i = p;
    }
    h->a[i] = k;
}
static uint64_t maxheap_pop(heap_t *h) {
    uint64_t top = h->a[0], last = h->a[--h->n];
    int i = 0;
    for (;;) {
        int l = 2 * i + 1, r = l + 1;
        if (l >= h->n) break;
        int c = (r < h->n && h->a[r] > h->a[l]) ? r : l;
        if (h->a[c] <= last) break;
        h->a[i] = h->a[c];
        i = c;
    }
    if (h->n > 0) h->a[i] = last;
    return top;
}
static void minheap_push(heap_t *h, uint64_t k) {
    if (h->n == h->cap) heap_reserve(h, h->cap * 2 + 64);
    int i = h->n++;
    while (i > 0) {
        int p = (i - 1) >> 1;
        if (h->a[p] <= k) break;
        h->a[i] = h->a[p];
        i = p;
    }
    h->a[i] = k;
}
static void minheap_replace_top(heap_t *h, uint64_t k) {
    int i = 0;
    for (;;) {
        int l = 2 * i + 1, r = l + 1;
        if (l >= h->n) break;
        int c = (r < h->n && h->a[r] < h->a[l]) ? r : l;
        if (h->a[c] >= k) break;
        h->a[i] = h->a[c];
        i = c;
    }
    h->a[i] = k;
}
/* bounded min-heap push: keep the `bound` best keys */
static inline void bounded_push(heap_t *h, int bound, uint64_t k) {
    if (h->n < bound)
        minheap_push(h, k);
    else if (k > h->a[0])
        minheap_replace_top(h, k);
}
static int cmp_u64_desc(const void *x, const void *y) {
    uint64_t a = *(const uint64_t *)x, b = *(const uint64_t *)y;
    return a < b ? 1 : (a > b ? -1 : 0);
}

/* ------------------------------------------------------------------------------------------
 * Product quantisation (call sites JVectorIndexQuantization.java:114-140,
 * JVectorWriter.java:1117-1124, JVectorReader.java:352-357).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int dim, M, K;
    int *size, *off;     /* sub-vector sizes / offsets into the vector */
    int64_t *cb_off;     /* offset of subspace m's codebook in the flat codebook array */
} pq_shape;

static void pq_shape_init(pq_shape *s, int dim, int M, int K) {
    s->dim = dim;
    s->M = M;
    s->K = K;
    s->size = (int *)malloc(sizeof(int) * (size_t)M);
    s->off = (int *)malloc(sizeof(int) * (size_t)M);
    s->cb_off = (int64_t *)malloc(sizeof(int64_t) * (size_t)M);
    int base = dim / M, rem = dim % M, o = 0;
    int64_t co = 0;
    for (int m = 0; m < M; m++) {
        s->size[m] = base + (m < rem ? 1 : 0); /* first dim%M sub-vectors get base+1 */
        s->off[m] = o;
        s->cb_off[m] = co;
        o += s->size[m];
        co += (int64_t)K * s->size[m];
    }
}
static void pq_shape_free(pq_shape *s) {
    free(s->size);
    free(s->off);
    free(s->cb_off);
}

JVO_EXPORT void jvo_pq_subspaces(int32_t dim, int32_t M, int32_t *sizes, int32_t *offsets) {
    int base = dim / M, rem = dim % M, o = 0;
    for (int m = 0; m < M; m++) {
        sizes[m] = base + (m < rem ? 1 : 0);
        offsets[m] = o;
        o += sizes[m];
    }
}

/* JVectorIndexQuantization.PQ.defaultNumSubspaces, JVectorIndexQuantization.java:428-446 */
JVO_EXPORT int32_t jvo_default_num_subspaces(int32_t d) {
    if (d <= 32) return d;
    if (d <= 64) return 32;
    if (d <= 200) return (int32_t)(d * 0.5);
    if (d <= 400) return 100;
    if (d <= 768) return (int32_t)(d * 0.25);
    if (d <= 1536) return 192;
    return (int32_t)(d * 0.125);
}

static inline float sub_l2sq(const float *x, const float *c, int len) {
    float acc = 0.0f;
    for (int j = 0; j < len; j++) {
        float d = x[j] - c[j];
        acc = fmaf(d, d, acc);
    }
    return acc;
}
static inline float sub_dot(const float *x, const float *c, int len) {
    float acc = 0.0f;
    for (int j = 0; j < len; j++) acc = fmaf(x[j], c[j], acc);
    return acc;
}

/* K6: ProductQuantization.encode — x' = x - g; code[m] = first argmin_c ||x'_m - C_m[c]||^2 (strict <). */
JVO_EXPORT void jvo_pq_encode(const float *vectors, int64_t n, int32_t dim, int32_t M, int32_t K, const float *codebooks,
                              const float *gcent, uint8_t *out_codes, int32_t threads) {
    pq_shape s;
    pq_shape_init(&s, dim, M, K);
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
    for (int64_t i = 0; i < n; i++) {
        float *xc = (float *)malloc(sizeof(float) * (size_t)dim);
        const float *x = vectors + i * dim;
        for (int d = 0; d < dim; d++) xc[d] = gcent ? x[d] - gcent[d] : x[d];
        for (int m = 0; m < M; m++) {
            const float *cb = codebooks + s.cb_off[m];
            float best = INFINITY;
            int idx = 0;
            for (int c = 0; c < K; c++) {
                float d2 = sub_l2sq(xc + s.off[m], cb + (int64_t)c * s.size[m], s.size[m]);
                if (d2 < best) {
                    best = d2;
                    idx = c;
                }
            }
            out_codes[i * M + m] = (uint8_t)idx;
        }
        free(xc);
    }
    pq_shape_free(&s);
}

/* K1: PQDecoder precomputed table.  DOT/COSINE/MIP: lut[m][c] = q_m . C_m[c];
 * EUCLIDEAN: lut[m][c] = ||(q-g)_m - C_m[c]||^2. */
static void pq_build_lut(const pq_shape *s, int sim, const float *codebooks, const float *gcent, const float *q, float *lut,
                         float *qc_scratch) {
    const float *qq = q;
    if (sim == JV_SIM_EUCLIDEAN && gcent) {
        for (int d = 0; d < s->dim; d++) qc_scratch[d] = q[d] - gcent[d];
        qq = qc_scratch;
    }
    for (int m = 0; m < s->M; m++) {
        const float *cb = codebooks + s->cb_off[m];
        for (int c = 0; c < s->K; c++) {
            const float *cv = cb + (int64_t)c * s->size[m];
            lut[m * s->K + c] = (sim == JV_SIM_EUCLIDEAN) ? sub_l2sq(qq + s->off[m], cv, s->size[m])
                                                          : sub_dot(qq + s->off[m], cv, s->size[m]);
        }
    }
}

JVO_EXPORT void jvo_pq_lut(int32_t sim, int32_t dim, int32_t M, int32_t K, const float *codebooks, const float *gcent,
                           const float *queries, int32_t nq, float *out_lut) {
    pq_shape s;
    pq_shape_init(&s, dim, M, K);
    float *scratch = (float *)malloc(sizeof(float) * (size_t)dim);
    for (int i = 0; i < nq; i++)
        pq_build_lut(&s, sim, codebooks, gcent, queries + (int64_t)i * dim, out_lut + (int64_t)i * M * K, scratch);
    free(scratch);
    pq_shape_free(&s);
}

/* ||decode(code)||^2 = sum_m ||C_m[code_m]||^2 : the cosine decoder's second table, folded per node
 * because it does not depend on the query. */
static float pq_node_norm(const pq_shape *s, const float *codebooks, const uint8_t *code) {
    float acc = 0.0f;
    for (int m = 0; m < s->M; m++) {
        const float *cv = codebooks + s->cb_off[m] + (int64_t)code[m] * s->size[m];
        acc += sub_dot(cv, cv, s->size[m]);
    }
    return acc;
}

/* a4: assembleAndSum + decoder mapping.  The order in which the M table entries are summed is not
 * observable in the reference beyond rounding (Panama lane-wise partials, host dependent).  Two
 * fixed orders are provided:
 *   order 0  four interleaved sequential accumulators (scalar-loop flavour)
 *   order N  (N = 1,2,4,..,32 lanes): lane l sums subspaces 4w..4w+3 for code words w = l, l+N, ..
 *            in increasing m, then a halving tree over the N lanes.  N = 32 is the strict sm_100a kernel,
 *            N = adc_lanes(M) the fast kernel (<= 4 code words per lane),
 *            so fp32 traversals can be compared bit for bit. */
static inline float adc_sum(const float *lut, int M, int K, const uint8_t *code, int order) {
    if (order >= 1) { /* order = number of lanes (power of two <= 32) sharing one code row */
        float v[32];
        const int lanes = order > 32 ? 32 : order;
        const int nwords = (M + 3) >> 2;
        for (int l = 0; l < lanes; l++) {
            float s = 0.f;
            for (int w = l; w < nwords; w += lanes)
                for (int b = 0; b < 4; b++)
                    if (4 * w + b < M) s += lut[(4 * w + b) * K + code[4 * w + b]];
            v[l] = s;
        }
        for (int off = lanes >> 1; off >= 1; off >>= 1)
            for (int j = 0; j < off; j++) v[j] = v[j] + v[j + off];
        return v[0];
    }
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int m = 0;
    for (; m + 4 <= M; m += 4) {
        s0 += lut[(m + 0) * K + code[m + 0]];
        s1 += lut[(m + 1) * K + code[m + 1]];
        s2 += lut[(m + 2) * K + code[m + 2]];
        s3 += lut[(m + 3) * K + code[m + 3]];
    }
    for (; m < M; m++) s0 += lut[m * K + code[m]];
    return (s0 + s1) + (s2 + s3);
}

static inline float adc_score(int sim, const float *lut, int M, int K, const uint8_t *code, float node_norm, float qnorm,
                              int order) {
    float s = adc_sum(lut, M, K, code, order);
    switch (sim) {
    case JV_SIM_EUCLIDEAN:
        return 1.0f / (1.0f + s);
    case JV_SIM_COSINE:
        return (1.0f + s / sqrtf(node_norm * qnorm)) * 0.5f;
    default:
        return (1.0f + s) * 0.5f;
    }
}

/* ------------------------------------------------------------------------------------------
 * 8-bit quantised ADC table ("adc_order = -8"): the GPU production traversal keeps the per-query table as
 * bytes (one shared scale per query) and sums integers, as jVector's fused-ADC / FAISS fast-scan style
 * scorers do.  Scores only steer the traversal; returned scores come from the exact rerank.
 *   ball_m   = bounding ball (centre, radius) of subspace m's centroids            (index constant)
 *   [lo_m, hi_m] = bound on the table row m from the ball: dot  q_m.ctr_m -/+ |q_m| r_m,
 *                                                          L2   (max(0, |q_m-ctr_m| - r_m))^2 .. (|q_m-ctr_m| + r_m)^2
 *   range = max_m (hi_m - lo_m);  inv = 255/range;  delta = range/255;  base = sum_m lo_m  (in m order)
 *   q8[m][c] = sat_u8(rint(fmaf(lut[m][c], inv, -(lo_m*inv))))
 *   partial-sum estimate = fmaf(delta, (float)sum_m q8[m][code_m], base)
 * The device kernel (jv_lut_q8.cu) performs exactly these operations, so tables and sums are bit-identical.
 * ------------------------------------------------------------------------------------------ */
static void pq_ball(const pq_shape *s, const float *codebooks, float *ctr, float *rad) {
    for (int m = 0; m < s->M; m++) {
        const float *cb = codebooks + s->cb_off[m];
        const int len = s->size[m];
        for (int j = 0; j < len; j++) {
            double acc = 0.0;
            for (int c = 0; c < s->K; c++) acc += (double)cb[(int64_t)c * len + j];
            ctr[s->off[m] + j] = (float)(acc / (double)s->K);
        }
        double r2max = 0.0;
        for (int c = 0; c < s->K; c++) {
            double r2 = 0.0;
            for (int j = 0; j < len; j++) {
                double d = (double)cb[(int64_t)c * len + j] - (double)ctr[s->off[m] + j];
                r2 += d * d;
            }
            if (r2 > r2max) r2max = r2;
        }
        rad[m] = (float)sqrt(r2max) * 1.0009765625f; /* 2^-10 slack over the rounded radius */
    }
}

/* lo[m], returns range; qq = (centred) query */
static float q8_bounds(const pq_shape *s, int sim, const float *qq, const float *ctr, const float *rad, float *lo) {
    float range = 0.f;
    for (int m = 0; m < s->M; m++) {
        const float *x = qq + s->off[m], *c = ctr + s->off[m];
        const int len = s->size[m];
        float l, h;
        if (sim == JV_SIM_EUCLIDEAN) {
            const float d = sqrtf(sub_l2sq(x, c, len));
            float a = d - rad[m];
            if (a < 0.f) a = 0.f;
            const float b = d + rad[m];
            l = a * a;
            h = b * b;
        } else {
            const float w = sqrtf(sub_dot(x, x, len)) * rad[m];
            const float qc = sub_dot(x, c, len);
            l = qc - w;
            h = qc + w;
        }
        lo[m] = l;
        const float r = h - l;
        if (r > range) range = r;
    }
    return range;
}

static inline uint8_t sat_u8_rn(float x) {
    if (!(x > 0.f)) return 0; /* also NaN */
    if (x >= 255.f) return 255;
    return (uint8_t)(int)rintf(x);
}

/* lut (fp32, [M][K]) -> q8 [M][K]; params[0] = delta, params[1] = base */
static void q8_quantise(const pq_shape *s, int sim, const float *qq, const float *ctr, const float *rad, const float *lut,
                        uint8_t *q8, float *lo_scratch, float *params) {
    const float range = q8_bounds(s, sim, qq, ctr, rad, lo_scratch);
    const float inv = range > 0.f ? 255.0f / range : 0.f;
    float base = 0.f;
    for (int m = 0; m < s->M; m++) base += lo_scratch[m];
    for (int m = 0; m < s->M; m++) {
        const float nlo = -(lo_scratch[m] * inv);
        for (int c = 0; c < s->K; c++) q8[m * s->K + c] = sat_u8_rn(fmaf(lut[m * s->K + c], inv, nlo));
    }
    params[0] = range / 255.0f;
    params[1] = base;
}

static inline float adc_score_q8(int sim, const uint8_t *q8, const float *params, int M, int K, const uint8_t *code,
                                 float node_norm, float qnorm) {
    uint32_t isum = 0;
    for (int m = 0; m < M; m++) isum += q8[m * K + code[m]];
    const float s = fmaf(params[0], (float)isum, params[1]);
    switch (sim) {
    case JV_SIM_EUCLIDEAN:
        return 1.0f / (1.0f + s);
    case JV_SIM_COSINE:
        return (1.0f + s / sqrtf(node_norm * qnorm)) * 0.5f;
    default:
        return (1.0f + s) * 0.5f;
    }
}

/* ------------------------------------------------------------------------------------------
 * NVQ (non-uniform vector quantisation) of the inline vectors, "nvq+pq" segments.
 * Decoder: literal restatement of the IN-TREE code JVectorIndexQuantization.java:316-361
 * (nvqDequantize, logisticNQT, logitNQT) — Math.fma -> fmaf, Math.round(float) -> floor(x + 1/2)
 * evaluated exactly, Float.floatToIntBits / intBitsToFloat -> memcpy.
 * Encoder: FIXTURE ONLY.  jVector trains (growthRate, midpoint) per sub-vector with an optimiser that
 * is out of tree; any (growthRate, midpoint, minValue, maxValue, bytes) tuple is a valid input of
 * the decoder, so the fixture uses fixed shape parameters and picks every byte as the best
 * reconstruction under that decoder.
 * ------------------------------------------------------------------------------------------ */
static inline int32_t f2i_bits(float f) {
    int32_t i;
    memcpy(&i, &f, 4);
    return i;
}
static inline float i2f_bits(int32_t i) {
    float f;
    memcpy(&f, &i, 4);
    return f;
}
static inline int java_round_f(float a) { return (int)floor((double)a + 0.5); } /* Math.round(float), exact */

/* JVectorIndexQuantization.java:344-351 */
static float logistic_nqt(float value, float alpha, float x0) {
    float temp = fmaf(value, alpha, -alpha * x0);
    int p = java_round_f(temp + 0.5f);
    int m = f2i_bits(fmaf(temp - (float)p, 0.5f, 1.0f));
    temp = i2f_bits(m + (int32_t)((uint32_t)p << 23));
    return temp / (temp + 1.0f);
}
/* JVectorIndexQuantization.java:354-361 */
static float logit_nqt(float scaled, float inverse_alpha, float x0) {
    float z = scaled / (1.0f - scaled);
    int32_t temp = f2i_bits(z);
    int32_t e = temp & 0x7f800000;
    float p = (float)((e >> 23) - 128);
    float m = i2f_bits((temp & 0x007fffff) + 0x3f800000);
    return (m + p) * inverse_alpha + x0;
}
typedef struct {
    float scale, bias, inv_alpha, mid;
} nvq_sub_t;
/* the per-sub-vector constants of nvqDequantize, JVectorIndexQuantization.java:320-326 */
static nvq_sub_t nvq_sub(const float *prm /* growthRate, midpoint, minValue, maxValue */) {
    nvq_sub_t r;
    float delta = prm[3] - prm[2];
    float sgr = prm[0] / delta;
    r.mid = prm[1] * delta;
    r.bias = logistic_nqt(prm[2], sgr, r.mid);
    r.scale = (logistic_nqt(prm[3], sgr, r.mid) - r.bias) / 255.0f;
    r.inv_alpha = 1.0f / sgr;
    return r;
}
static inline float nvq_component(const nvq_sub_t *c, int b) { return logit_nqt(fmaf((float)b, c->scale, c->bias), c->inv_alpha, c->mid); }

/* nvqDequantize for one vector: out[dim] = decoded + globalMean */
static void nvq_dequantize_one(const pq_shape *sh, const uint8_t *bytes, const float *params, const float *gmean, float *out) {
    for (int m = 0; m < sh->M; m++) {
        nvq_sub_t c = nvq_sub(params + 4 * m);
        for (int d = 0; d < sh->size[m]; d++) out[sh->off[m] + d] = nvq_component(&c, bytes[sh->off[m] + d]);
    }
    for (int i = 0; i < sh->dim; i++) out[i] = out[i] + gmean[i];
}

JVO_EXPORT void jvo_nvq_dequantize(int64_t n, int32_t dim, int32_t nvq_m, const uint8_t *bytes, const float *params,
                                   const float *gmean, float *out) {
    pq_shape sh;
    pq_shape_init(&sh, dim, nvq_m, 1);
    for (int64_t i = 0; i < n; i++) nvq_dequantize_one(&sh, bytes + i * dim, params + i * nvq_m * 4, gmean, out + i * dim);
    pq_shape_free(&sh);
}

/* FIXTURE encoder: globalMean = per-dimension mean (double sums); per sub-vector minValue/maxValue = extremes of the
 * centred components, growthRate/midpoint as given; every byte = the code whose reconstruction is closest. */
JVO_EXPORT void jvo_nvq_encode(const float *vectors, int64_t n, int32_t dim, int32_t nvq_m, float growth_rate, float midpoint,
                               uint8_t *out_bytes, float *out_params, float *out_gmean) {
    pq_shape sh;
    pq_shape_init(&sh, dim, nvq_m, 1);
    for (int d = 0; d < dim; d++) {
        double acc = 0.0;
        for (int64_t i = 0; i < n; i++) acc += (double)vectors[i * dim + d];
        out_gmean[d] = n > 0 ? (float)(acc / (double)n) : 0.f;
    }
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int64_t i = 0; i < n; i++) {
        for (int m = 0; m < nvq_m; m++) {
            float *prm = out_params + (i * nvq_m + m) * 4;
            float lo = INFINITY, hi = -INFINITY;
            for (int d = 0; d < sh.size[m]; d++) {
                float x = vectors[i * dim + sh.off[m] + d] - out_gmean[sh.off[m] + d];
                if (x < lo) lo = x;
                if (x > hi) hi = x;
            }
            if (!(hi > lo)) hi = lo + 1e-6f;
            prm[0] = growth_rate, prm[1] = midpoint, prm[2] = lo, prm[3] = hi;
            nvq_sub_t c = nvq_sub(prm);
            float table[256];
            for (int b = 0; b < 256; b++) table[b] = nvq_component(&c, b);
            for (int d = 0; d < sh.size[m]; d++) {
                float x = vectors[i * dim + sh.off[m] + d] - out_gmean[sh.off[m] + d];
                int best = 0;
                float be = INFINITY;
                for (int b = 0; b < 256; b++) {
                    float e = fabsf(table[b] - x);
                    if (e < be) be = e, best = b;
                }
                out_bytes[i * dim + sh.off[m] + d] = (uint8_t)best;
            }
        }
    }
    pq_shape_free(&sh);
}

/* ------------------------------------------------------------------------------------------
 * Index view + per-thread search scratch
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    jv_index_desc d;
    pq_shape pq;
    int has_pq;
    int adc_order;    /* see adc_sum */
    float *node_norm; /* cosine + PQ only */
    float *ball_ctr, *ball_rad; /* adc_order -8 only */
    int has_nvq;
    pq_shape nvq; /* sub-vector split of the NVQ-inline vectors */
} jvo_index;

JVO_EXPORT jvo_index *jvo_index_create(const jv_index_desc *desc) {
    jvo_index *ix = (jvo_index *)calloc(1, sizeof(jvo_index));
    memset(&ix->d, 0, sizeof(ix->d)); /* struct_size 96 = layout without the NVQ fields */
    memcpy(&ix->d, desc, (desc->struct_size > 0 && (size_t)desc->struct_size < sizeof(ix->d)) ? (size_t)desc->struct_size : sizeof(ix->d));
    ix->has_pq = desc->pq_m > 0 && desc->pq_codes && desc->pq_codebooks;
    ix->has_nvq = ix->d.nvq_m > 0 && ix->d.nvq_bytes && ix->d.nvq_params && ix->d.nvq_global_mean;
    if (!ix->has_nvq) ix->d.nvq_m = 0;
    if (ix->has_nvq) pq_shape_init(&ix->nvq, desc->dim, desc->nvq_m, 1);
    if (ix->has_pq) {
        pq_shape_init(&ix->pq, desc->dim, desc->pq_m, desc->pq_k);
        if (desc->similarity == JV_SIM_COSINE) {
            ix->node_norm = (float *)malloc(sizeof(float) * (size_t)desc->n);
            for (int64_t i = 0; i < desc->n; i++)
                ix->node_norm[i] = pq_node_norm(&ix->pq, desc->pq_codebooks, desc->pq_codes + i * desc->pq_m);
        }
    }
    return ix;
}
JVO_EXPORT void jvo_index_set_adc_order(jvo_index *ix, int32_t order) {
    ix->adc_order = order;
    if (order == -8 && ix->has_pq && !ix->ball_ctr) {
        ix->ball_ctr = (float *)malloc(sizeof(float) * (size_t)ix->d.dim);
        ix->ball_rad = (float *)malloc(sizeof(float) * (size_t)ix->d.pq_m);
        pq_ball(&ix->pq, ix->d.pq_codebooks, ix->ball_ctr, ix->ball_rad);
    }
}
JVO_EXPORT void jvo_index_destroy(jvo_index *ix) {
    if (!ix) return;
    if (ix->has_pq) pq_shape_free(&ix->pq);
    if (ix->has_nvq) pq_shape_free(&ix->nvq);
    free(ix->node_norm);
    free(ix->ball_ctr);
    free(ix->ball_rad);
    free(ix);
}

typedef struct {
    heap_t cand, res;
    uint64_t *visited; /* bitmap over ordinals */
    int32_t *touched;
    int ntouched, touched_cap;
    float *lut, *qc;
    uint8_t *lut8; /* adc_order -8 */
    float *lo8, q8p[2];
    float *deq; /* NVQ: dequantised vector of the node being reranked */
    uint64_t *sorted;
} scratch_t;

static scratch_t *scratch_new(const jvo_index *ix) {
    scratch_t *s = (scratch_t *)calloc(1, sizeof(scratch_t));
    s->visited = (uint64_t *)calloc((size_t)((ix->d.n + 63) / 64), 8);
    s->touched_cap = 4096;
    s->touched = (int32_t *)malloc(sizeof(int32_t) * (size_t)s->touched_cap);
    if (ix->has_pq) {
        s->lut = (float *)malloc(sizeof(float) * (size_t)ix->d.pq_m * ix->d.pq_k);
        s->lut8 = (uint8_t *)malloc((size_t)ix->d.pq_m * ix->d.pq_k);
        s->lo8 = (float *)malloc(sizeof(float) * (size_t)ix->d.pq_m);
    }
    s->qc = (float *)malloc(sizeof(float) * (size_t)ix->d.dim);
    s->deq = (float *)malloc(sizeof(float) * (size_t)ix->d.dim);
    return s;
}
static void scratch_free(scratch_t *s) {
    free(s->cand.a);
    free(s->res.a);
    free(s->visited);
    free(s->touched);
    free(s->lut);
    free(s->lut8);
    free(s->lo8);
    free(s->qc);
    free(s->deq);
    free(s->sorted);
    free(s);
}
static inline int visit(scratch_t *s, int32_t node) {
    uint64_t bit = 1ULL << (node & 63);
    uint64_t *w = &s->visited[node >> 6];
    if (*w & bit) return 0;
    *w |= bit;
    if (s->ntouched == s->touched_cap) {
        s->touched_cap *= 2;
        s->touched = (int32_t *)realloc(s->touched, sizeof(int32_t) * (size_t)s->touched_cap);
    }
    s->touched[s->ntouched++] = node;
    return 1;
}
static void visited_reset(scratch_t *s) {
    for (int i = 0; i < s->ntouched; i++) s->visited[s->touched[i] >> 6] = 0;
    s->ntouched = 0;
}

/* a9: accept-bits lambda, JVectorReader.java:157-163 + GraphNodeIdToDocMap.java:159-161 */
static inline int accepted(const jvo_index *ix, const uint64_t *bits, int32_t ord) {
    int32_t doc = ix->d.ord_to_doc ? ix->d.ord_to_doc[ord] : ord;
    if (doc == -1) return 0; /* deleted / no-vector ordinals are never returned */
    if (!bits) return 1;
    return (int)((bits[doc >> 6] >> (doc & 63)) & 1ULL);
}

/* builds S->lut (and, for adc_order -8, S->lut8 + S->q8p) for one query */
static void build_tables(const jvo_index *ix, scratch_t *S, const float *q) {
    const jv_index_desc *d = &ix->d;
    pq_build_lut(&ix->pq, d->similarity, d->pq_codebooks, d->pq_global_centroid, q, S->lut, S->qc);
    if (ix->adc_order == -8) {
        const float *qq = (d->similarity == JV_SIM_EUCLIDEAN && d->pq_global_centroid) ? S->qc : q;
        q8_quantise(&ix->pq, d->similarity, qq, ix->ball_ctr, ix->ball_rad, S->lut, S->lut8, S->lo8, S->q8p);
    }
}
static inline float approx_score(const jvo_index *ix, const scratch_t *S, int32_t node, float qnorm) {
    const jv_index_desc *d = &ix->d;
    const float nn = ix->node_norm ? ix->node_norm[node] : 0.f;
    const uint8_t *code = d->pq_codes + (int64_t)node * d->pq_m;
    if (ix->adc_order == -8) return adc_score_q8(d->similarity, S->lut8, S->q8p, d->pq_m, d->pq_k, code, nn, qnorm);
    return adc_score(d->similarity, S->lut, d->pq_m, d->pq_k, code, nn, qnorm, ix->adc_order);
}

/* ------------------------------------------------------------------------------------------
 * GraphSearcher.search for ONE query on the flat (level-0) graph, followed by the
 * rerank step, ordinal->doc mapping and collector ordering — i.e. JVectorReader.search,
 * JVectorReader.java:130-210.  approx_out (nullable, [rerank_k] keys) receives the approximate
 * result list before rerank, best first, for the graph builder and for tests.
 * ------------------------------------------------------------------------------------------ */
static void search_one(const jvo_index *ix, scratch_t *S, const float *q, int k, int rerank_k, float threshold,
                       float rerank_floor, const uint64_t *bits, int32_t *out_doc, float *out_score, int32_t *out_count,
                       jv_query_stats *st, uint64_t *approx_out, int32_t *approx_count, int entry_override,
                       int64_t n_limit) {
    const jv_index_desc *d = &ix->d;
    const int sim = d->similarity, dim = d->dim, R = d->max_degree;
    const int use_pq = ix->has_pq;
    const float qnorm = canon_dot(q, q, dim);
    const float mip_mul = (sim == JV_SIM_MIP && !use_pq) ? 2.0f : 1.0f; /* wrapExactScoreFunction, :220-239 */
    if (use_pq) build_tables(ix, S, q);

#define APPROX(node)                                                                                                    \
    (use_pq ? approx_score(ix, S, (node), qnorm)                                                                        \
            : exact_score(sim, q, qnorm, d->vectors + (int64_t)(node) * dim, dim) * mip_mul)

    S->cand.n = 0;
    S->res.n = 0;
    int visited = 0, expanded = 0, reranked = 0;
    int32_t entry = entry_override >= 0 ? entry_override : d->entry_node;
    if (d->n > 0 && entry >= 0) {
        visit(S, entry);
        visited++;
        maxheap_push(&S->cand, mk_key(APPROX(entry), entry));
    }
    while (S->cand.n > 0) {
        uint64_t top = S->cand.a[0];
        float sc = key_score(top);
        if (S->res.n >= rerank_k && sc < key_score(S->res.a[0])) break;
        maxheap_pop(&S->cand);
        int32_t c = key_node(top);
        if (accepted(ix, bits, c) && sc >= threshold) bounded_push(&S->res, rerank_k, top);
        expanded++;
        const int32_t *nb = d->adjacency + (int64_t)c * R;
        for (int j = 0; j < R; j++) {
            int32_t nn = nb[j];
            if (nn < 0) break;
            if (nn >= n_limit) continue; /* builder: nodes not inserted yet */
            if (!visit(S, nn)) continue;
            visited++;
            maxheap_push(&S->cand, mk_key(APPROX(nn), nn));
        }
    }
#undef APPROX
    visited_reset(S);

    /* approximate results, best first */
    int na = S->res.n;
    S->sorted = (uint64_t *)realloc(S->sorted, sizeof(uint64_t) * (size_t)(na > 0 ? na : 1));
    memcpy(S->sorted, S->res.a, sizeof(uint64_t) * (size_t)na);
    qsort(S->sorted, (size_t)na, sizeof(uint64_t), cmp_u64_desc);
    if (approx_out) {
        memcpy(approx_out, S->sorted, sizeof(uint64_t) * (size_t)na);
        *approx_count = na;
    }

    /* rerank (a5) — only when a reranker exists, i.e. the PQ path (JVectorReader.java:352-356) */
    heap_t fin = {0};
    if (out_doc) {
        for (int i = 0; i < na; i++) {
            int32_t node = key_node(S->sorted[i]);
            float s = key_score(S->sorted[i]);
            if (use_pq) {
                if (s < rerank_floor) continue;
                const float *x = d->vectors + (int64_t)node * dim;
                if (ix->has_nvq) { /* nvq+pq: view.rerankerFor scores the dequantised inline vector (JVectorReader.java:352-358) */
                    nvq_dequantize_one(&ix->nvq, d->nvq_bytes + (int64_t)node * dim, d->nvq_params + (int64_t)node * d->nvq_m * 4,
                                       d->nvq_global_mean, S->deq);
                    x = S->deq;
                }
                s = exact_score(sim, q, qnorm, x, dim); /* reranker is NOT x2-wrapped */
                reranked++;
            }
            int32_t doc = d->ord_to_doc ? d->ord_to_doc[node] : node;
            bounded_push(&fin, k, mk_key(s, doc)); /* collector: tie -> lower docId (a10) */
        }
        qsort(fin.a, (size_t)fin.n, sizeof(uint64_t), cmp_u64_desc);
        for (int i = 0; i < k; i++) {
            out_doc[i] = i < fin.n ? key_node(fin.a[i]) : -1;
            out_score[i] = i < fin.n ? key_score(fin.a[i]) : 0.0f;
        }
        if (out_count) *out_count = fin.n;
        free(fin.a);
    }
    if (st) {
        st->visited = visited;
        st->expanded = expanded;
        st->expanded_base = expanded;
        st->reranked = reranked;
    }
}

JVO_EXPORT int32_t jvo_search_batch(const jvo_index *ix, const float *queries, int32_t nq, int32_t k, int32_t rerank_k,
                                    float threshold, float rerank_floor, const uint64_t *accept_bits,
                                    int64_t accept_stride_words, int32_t *out_doc, float *out_score, int32_t *out_count,
                                    jv_query_stats *stats, int32_t threads) {
    if (rerank_k < k) return -1; /* GraphSearcher requires rerankK >= topK */
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel num_threads(threads)
#endif
    {
        scratch_t *S = scratch_new(ix);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
        for (int i = 0; i < nq; i++) {
            const uint64_t *bits = accept_bits ? accept_bits + (int64_t)i * accept_stride_words : NULL;
            search_one(ix, S, queries + (int64_t)i * ix->d.dim, k, rerank_k, threshold, rerank_floor, bits,
                       out_doc + (int64_t)i * k, out_score + (int64_t)i * k, out_count ? out_count + i : NULL,
                       stats ? stats + i : NULL, NULL, NULL, -1, ix->d.n);
        }
        scratch_free(S);
    }
    return 0;
}

/* approximate result lists only (before rerank): nodes/scores [nq*rerank_k], -1 padded */
JVO_EXPORT int32_t jvo_search_approx(const jvo_index *ix, const float *queries, int32_t nq, int32_t rerank_k,
                                     float threshold, const uint64_t *accept_bits, int64_t accept_stride_words,
                                     int32_t *out_node, float *out_score, int32_t *out_count, jv_query_stats *stats) {
    scratch_t *S = scratch_new(ix);
    uint64_t *keys = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)rerank_k);
    for (int i = 0; i < nq; i++) {
        int32_t cnt = 0;
        const uint64_t *bits = accept_bits ? accept_bits + (int64_t)i * accept_stride_words : NULL;
        search_one(ix, S, queries + (int64_t)i * ix->d.dim, rerank_k, rerank_k, threshold, 0.f, bits, NULL, NULL, NULL,
                   stats ? stats + i : NULL, keys, &cnt, -1, ix->d.n);
        for (int j = 0; j < rerank_k; j++) {
            out_node[(int64_t)i * rerank_k + j] = j < cnt ? key_node(keys[j]) : -1;
            out_score[(int64_t)i * rerank_k + j] = j < cnt ? key_score(keys[j]) : 0.f;
        }
        if (out_count) out_count[i] = cnt;
    }
    free(keys);
    scratch_free(S);
    return 0;
}

/* ADC score of explicit (query,node) pairs (a4) */
JVO_EXPORT void jvo_pq_adc_scores(const jvo_index *ix, const float *queries, int32_t nq, const int32_t *nodes,
                                  int32_t per_query, float *out) {
    scratch_t *S = scratch_new(ix);
    const jv_index_desc *d = &ix->d;
    for (int i = 0; i < nq; i++) {
        const float *q = queries + (int64_t)i * d->dim;
        float qnorm = canon_dot(q, q, d->dim);
        build_tables(ix, S, q);
        for (int j = 0; j < per_query; j++) {
            int32_t node = nodes[(int64_t)i * per_query + j];
            out[(int64_t)i * per_query + j] = approx_score(ix, S, node, qnorm);
        }
    }
    scratch_free(S);
}

/* K1 (8-bit): out_q8 [nq][M][K] in logical (m, c) order, out_params [nq][2] = (delta, base) */
JVO_EXPORT int32_t jvo_pq_lut_q8(const jvo_index *ix, const float *queries, int32_t nq, uint8_t *out_q8, float *out_params) {
    if (!ix->has_pq || ix->adc_order != -8) return -1;
    scratch_t *S = scratch_new(ix);
    const size_t mk = (size_t)ix->d.pq_m * ix->d.pq_k;
    for (int i = 0; i < nq; i++) {
        build_tables(ix, S, queries + (int64_t)i * ix->d.dim);
        memcpy(out_q8 + (size_t)i * mk, S->lut8, mk);
        out_params[2 * i] = S->q8p[0];
        out_params[2 * i + 1] = S->q8p[1];
    }
    scratch_free(S);
    return 0;
}

/* K5: brute-force exact top-k = Lucene exactSearch over JVectorVectorScorer.score()
 * (JVectorVectorScorer.java:36-53): jVector score, x2 for MIP; deleted ordinals skipped; tie -> lower doc. */
JVO_EXPORT int32_t jvo_exact_topk(const jvo_index *ix, const float *queries, int32_t nq, int32_t k,
                                  const uint64_t *accept_bits, int64_t accept_stride_words, int32_t *out_doc,
                                  float *out_score, int32_t *out_count, int32_t threads) {
    const jv_index_desc *d = &ix->d;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
    for (int i = 0; i < nq; i++) {
        const float *q = queries + (int64_t)i * d->dim;
        const uint64_t *bits = accept_bits ? accept_bits + (int64_t)i * accept_stride_words : NULL;
        float qnorm = canon_dot(q, q, d->dim);
        float mul = d->similarity == JV_SIM_MIP ? 2.0f : 1.0f;
        heap_t h = {0};
        for (int64_t o = 0; o < d->n; o++) {
            int32_t doc = d->ord_to_doc ? d->ord_to_doc[o] : (int32_t)o;
            if (doc == -1) continue;
            if (bits && !((bits[doc >> 6] >> (doc & 63)) & 1ULL)) continue;
            float s = exact_score(d->similarity, q, qnorm, d->vectors + o * d->dim, d->dim) * mul;
            bounded_push(&h, k, mk_key(s, doc));
        }
        qsort(h.a, (size_t)h.n, sizeof(uint64_t), cmp_u64_desc);
        for (int j = 0; j < k; j++) {
            out_doc[(int64_t)i * k + j] = j < h.n ? key_node(h.a[j]) : -1;
            out_score[(int64_t)i * k + j] = j < h.n ? key_score(h.a[j]) : 0.0f;
        }
        if (out_count) out_count[i] = h.n;
        free(h.a);
    }
    return 0;
}

/* K7: merge g per-shard lists [g][nq][k] -> top-k, tie -> lower doc (Lucene TopDocs.merge analogue) */
JVO_EXPORT void jvo_merge_topk(int32_t g, int32_t nq, int32_t k, const int32_t *docs, const float *scores, int32_t *out_doc,
                               float *out_score, int32_t *out_count) {
    uint64_t *keys = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)g * k);
    for (int i = 0; i < nq; i++) {
        int n = 0;
        for (int s = 0; s < g; s++)
            for (int j = 0; j < k; j++) {
                int64_t idx = ((int64_t)s * nq + i) * k + j;
                if (docs[idx] >= 0) keys[n++] = mk_key(scores[idx], docs[idx]);
            }
        qsort(keys, (size_t)n, sizeof(uint64_t), cmp_u64_desc);
        for (int j = 0; j < k; j++) {
            out_doc[(int64_t)i * k + j] = j < n ? key_node(keys[j]) : -1;
            out_score[(int64_t)i * k + j] = j < n ? key_score(keys[j]) : 0.f;
        }
        if (out_count) out_count[i] = n < k ? n : k;
    }
    free(keys);
}

/* ------------------------------------------------------------------------------------------
 * FIXTURE: PQ codebook training (ProductQuantization.compute, JVectorIndexQuantization.java:123-131;
 * SURVEY A.3 [M]): optional global centring, per-subspace k-means++ seeding + `iters` Lloyd
 * iterations.  Upstream training is randomised, so codebooks are only comparable given the same
 * generator; ours is splitmix64 so the device trainer (jv_pq_train) can reproduce it exactly.
 * Accumulations are sequential in double in ordinal order.
 * ------------------------------------------------------------------------------------------ */
static inline uint64_t splitmix64(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
#define JVO_KMPP_BLOCK 256
static inline double u01(uint64_t *s) { return (double)(splitmix64(s) >> 11) * (1.0 / 9007199254740992.0); }

JVO_EXPORT void jvo_pq_train(const float *vectors, int64_t n, int32_t dim, int32_t M, int32_t K, int32_t center,
                             int32_t iters, uint64_t seed, float *out_codebooks, float *out_gcent) {
    pq_shape s;
    pq_shape_init(&s, dim, M, K);
    float *g = NULL;
    if (center) {
        g = out_gcent;
        for (int d = 0; d < dim; d++) {
            double acc = 0.0;
            for (int64_t i = 0; i < n; i++) acc += (double)vectors[i * dim + d];
            g[d] = (float)(acc / (double)n);
        }
    }
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int m = 0; m < M; m++) {
        const int len = s.size[m], off = s.off[m];
        float *cb = out_codebooks + s.cb_off[m];
        float *x = (float *)malloc(sizeof(float) * (size_t)n * len); /* centred sub-vectors */
        for (int64_t i = 0; i < n; i++)
            for (int j = 0; j < len; j++) x[i * len + j] = g ? vectors[i * dim + off + j] - g[off + j] : vectors[i * dim + off + j];
        float *d2 = (float *)malloc(sizeof(float) * (size_t)n);
        int32_t *assign = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
        uint64_t rng = seed * 0x9E3779B97F4A7C15ULL + (uint64_t)m * 0xD1B54A32D192ED03ULL + 1;
        /* k-means++ seeding */
        int64_t first = (int64_t)(splitmix64(&rng) % (uint64_t)n);
        memcpy(cb, x + first * len, sizeof(float) * (size_t)len);
        for (int64_t i = 0; i < n; i++) d2[i] = sub_l2sq(x + i * len, cb, len);
        const int64_t nblk = (n + JVO_KMPP_BLOCK - 1) / JVO_KMPP_BLOCK;
        double *bsum = (double *)malloc(sizeof(double) * (size_t)nblk);
        for (int c = 1; c < K; c++) {
            /* D^2 sampling over a two-level sum (blocks of 256 in ordinal order) so that the device trainer,
             * which sums blocks in parallel, reproduces the same pick */
            double total = 0.0;
            for (int64_t b = 0; b < nblk; b++) {
                double acc = 0.0;
                const int64_t e = (b + 1) * JVO_KMPP_BLOCK < n ? (b + 1) * JVO_KMPP_BLOCK : n;
                for (int64_t i = b * JVO_KMPP_BLOCK; i < e; i++) acc += (double)d2[i];
                bsum[b] = acc;
                total += acc;
            }
            const double r = u01(&rng) * total;
            double run = 0.0;
            int64_t pick = n - 1;
            for (int64_t b = 0; b < nblk; b++) {
                if (run + bsum[b] > r || b == nblk - 1) {
                    const int64_t e = (b + 1) * JVO_KMPP_BLOCK < n ? (b + 1) * JVO_KMPP_BLOCK : n;
                    pick = e - 1;
                    for (int64_t i = b * JVO_KMPP_BLOCK; i < e; i++) {
                        run += (double)d2[i];
                        if (run > r) {
                            pick = i;
                            break;
                        }
                    }
                    break;
                }
                run += bsum[b];
            }
            float *cc = cb + (int64_t)c * len;
            memcpy(cc, x + pick * len, sizeof(float) * (size_t)len);
            for (int64_t i = 0; i < n; i++) {
                float dd = sub_l2sq(x + i * len, cc, len);
                if (dd < d2[i]) d2[i] = dd;
            }
        }
        /* Lloyd */
        double *sum = (double *)malloc(sizeof(double) * (size_t)K * len);
        int64_t *cnt = (int64_t *)malloc(sizeof(int64_t) * (size_t)K);
        for (int it = 0; it < iters; it++) {
            for (int64_t i = 0; i < n; i++) {
                float best = INFINITY;
                int idx = 0;
                for (int c = 0; c < K; c++) {
                    float dd = sub_l2sq(x + i * len, cb + (int64_t)c * len, len);
                    if (dd < best) {
                        best = dd;
                        idx = c;
                    }
                }
                assign[i] = idx;
            }
            memset(sum, 0, sizeof(double) * (size_t)K * len);
            memset(cnt, 0, sizeof(int64_t) * (size_t)K);
            for (int64_t i = 0; i < n; i++) {
                cnt[assign[i]]++;
                for (int j = 0; j < len; j++) sum[(int64_t)assign[i] * len + j] += (double)x[i * len + j];
            }
            for (int c = 0; c < K; c++)
                if (cnt[c] > 0) /* empty cluster keeps its previous centroid */
                    for (int j = 0; j < len; j++) cb[(int64_t)c * len + j] = (float)(sum[(int64_t)c * len + j] / (double)cnt[c]);
        }
        free(sum);
        free(cnt);
        free(x);
        free(d2);
        free(bsum);
        free(assign);
    }
    pq_shape_free(&s);
}

/* ------------------------------------------------------------------------------------------
 * FIXTURE: Vamana construction (GraphIndexBuilder, JVectorWriter.java:1383-1422; SURVEY A.4 [M]).
 * Batched-insert schedule shared with the device builder (jv_graph_build) so adjacency can be
 * compared exactly:
 *   entry = medoid (node whose exact score against the mean vector is best)
 *   batches: entry alone first; then, over the remaining ordinals in order, batches of size
 *            min(inserted, max(1, min(max_batch, frac * n)))  ("prefix doubling", ParlayANN style)
 *   per batch, against the graph FROZEN at batch start:
 *     1. search(q = vector[p], L = beamWidth, exact scores) -> candidates (best first, p itself excluded)
 *     2. out[p] = retainDiverse(candidates, R, alpha)
 *   then, in ordinal order of the batch: for nb in out[p]: append p to adj[nb];
 *   then every node touched whose degree > R*overflow is sorted best first, cut to the best 1024 and
 *   re-pruned with retainDiverse to R.
 *   cleanup: every node with degree > R re-pruned to R.
 * retainDiverse (jVector ConcurrentNeighborMap.retainDiverse [M]): for a = 1.0, 1.2, .. <= alpha: walk
 * candidates best first, select c unless some already-selected s has score(c,s) > score(c,p) * a.
 * ------------------------------------------------------------------------------------------ */
#define JVO_PRUNE_MAX_CANDS 1024
#define JVO_APPEND_CAP 2048

static int retain_diverse(const float *vecs, int dim, int sim, const int32_t *cn, const float *cs, int nc, int R,
                          float alpha, int32_t *out_nodes, float *out_scores) {
    int nsel = 0;
    uint8_t *taken = (uint8_t *)calloc((size_t)(nc > 0 ? nc : 1), 1);
    for (float a = 1.0f; a <= alpha + 1e-6f && nsel < R; a += 0.2f) {
        for (int i = 0; i < nc && nsel < R; i++) {
            if (taken[i]) continue;
            const float *cv = vecs + (int64_t)cn[i] * dim;
            float cnorm = canon_dot(cv, cv, dim);
            int diverse = 1;
            for (int j = 0; j < nsel; j++) {
                float sb = exact_score(sim == JV_SIM_MIP ? JV_SIM_DOT : sim, cv, cnorm, vecs + (int64_t)out_nodes[j] * dim, dim);
                if (sb > cs[i] * a) {
                    diverse = 0;
                    break;
                }
            }
            if (diverse) {
                taken[i] = 1;
                out_nodes[nsel] = cn[i];
                out_scores[nsel] = cs[i];
                nsel++;
            }
        }
    }
    free(taken);
    return nsel;
}

static void sort_by_key_desc(int32_t *nodes, float *scores, int n) {
    uint64_t *k = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) k[i] = mk_key(scores[i], nodes[i]);
    qsort(k, (size_t)n, sizeof(uint64_t), cmp_u64_desc);
    for (int i = 0; i < n; i++) {
        nodes[i] = key_node(k[i]);
        scores[i] = key_score(k[i]);
    }
    free(k);
}

/* n0 > 0: the first n0 ordinals already form a graph (seed_adj [n0][R], -1 padded at the end of each row; seed_entry):
 * "leading segment merge", JVectorWriter.java:1166-1341 (insert-only part) — the leading segment's graph is loaded with its
 * cached neighbour scores (here: recomputed, they are exact pair scores) and the other segments' vectors are added with
 * builder.addGraphNode; the entry node stays the leading graph's.  n0 == 0: build from scratch (medoid entry). */
static int32_t graph_build_impl(const float *vectors, int64_t n, int32_t dim, int32_t sim, int32_t R, int32_t beam,
                                float overflow, float alpha, int32_t max_batch, float frac, int64_t n0, const int32_t *seed_adj,
                                int32_t seed_entry, int32_t *out_adj, int32_t *out_entry) {
    const int bsim = sim == JV_SIM_MIP ? JV_SIM_DOT : sim;
    const int cap = (int)ceilf((float)R * overflow) + 1; /* list capacity incl. one overflow slot */
    /* medoid */
    float *mean = (float *)calloc((size_t)dim, sizeof(float));
    for (int d = 0; d < dim; d++) {
        double acc = 0;
        for (int64_t i = 0; i < n; i++) acc += vectors[i * dim + d];
        mean[d] = (float)(acc / (double)n);
    }
    float mnorm = canon_dot(mean, mean, dim);
    int32_t entry = 0;
    uint64_t bestk = 0;
    for (int64_t i = 0; i < n; i++) {
        uint64_t kk = mk_key(exact_score(bsim, mean, mnorm, vectors + i * dim, dim), (int32_t)i);
        if (kk > bestk) {
            bestk = kk;
            entry = (int32_t)i;
        }
    }
    free(mean);
    if (n0 > 0) entry = seed_entry; /* the leading graph keeps its entry node */
    *out_entry = entry;

    /* insertion order: entry first, then ordinals ascending (seeded: ordinals n0.. in order) */
    int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    if (n0 > 0) {
        for (int64_t i = 0; i < n; i++) order[i] = (int32_t)i;
    } else {
        order[0] = entry;
        for (int64_t i = 0, j = 1; i < n; i++)
            if (i != entry) order[j++] = (int32_t)i;
    }

    /* dynamic adjacency with scores (score of neighbour w.r.t. owner) */
    int32_t **adj = (int32_t **)calloc((size_t)n, sizeof(int32_t *));
    float **ads = (float **)calloc((size_t)n, sizeof(float *));
    int *deg = (int *)calloc((size_t)n, sizeof(int));
    int *acap = (int *)calloc((size_t)n, sizeof(int));
    uint8_t *inserted = (uint8_t *)calloc((size_t)n, 1);
    /* frozen flat view for the searcher: stride Rb = floor(R*overflow), the largest degree that
     * survives a batch (longer lists are re-pruned to R), so searches see the overflow edges too */
    const int Rb = (int)floorf((float)R * overflow) > R ? (int)floorf((float)R * overflow) : R;
    int32_t *flat = (int32_t *)malloc(sizeof(int32_t) * (size_t)n * Rb);
    for (int64_t i = 0; i < n * Rb; i++) flat[i] = -1;

    jv_index_desc d;
    memset(&d, 0, sizeof(d));
    d.similarity = bsim;
    d.dim = dim;
    d.max_degree = Rb;
    d.n = n;
    d.entry_node = entry;
    d.adjacency = flat;
    d.vectors = vectors;
    jvo_index *ix = jvo_index_create(&d);

    inserted[entry] = 1;
    int64_t done = 1;
    if (n0 > 0) { /* seed: lists of the leading graph, scores = exact pair scores (the neighbours-score cache) */
        for (int64_t u = 0; u < n0; u++) {
            int du = 0;
            while (du < R && seed_adj[u * R + du] >= 0) du++;
            acap[u] = cap + 8;
            adj[u] = (int32_t *)malloc(sizeof(int32_t) * (size_t)acap[u]);
            ads[u] = (float *)malloc(sizeof(float) * (size_t)acap[u]);
            const float *uv = vectors + u * dim;
            const float unorm = canon_dot(uv, uv, dim);
            for (int j = 0; j < du; j++) {
                adj[u][j] = seed_adj[u * R + j];
                ads[u][j] = exact_score(bsim, uv, unorm, vectors + (int64_t)adj[u][j] * dim, dim);
            }
            deg[u] = du;
            inserted[u] = 1;
            for (int j = 0; j < Rb; j++) flat[u * Rb + j] = j < du ? adj[u][j] : -1;
        }
        done = n0;
    }
    /* prefix doubling, capped at n/divisor (integer: the device builder must compute the same cap) and max_batch */
    const int64_t divisor = frac > 0.f ? (int64_t)(1.0f / frac + 0.5f) : 50;
    int64_t bcap = n / (divisor > 0 ? divisor : 50);
    if (bcap > max_batch) bcap = max_batch;
    if (bcap < 1) bcap = 1;
    while (done < n) {
        int64_t bs = done < bcap ? done : bcap;
        if (bs > n - done) bs = n - done;
        /* 1+2: search + prune against the frozen graph */
        int32_t *newn = (int32_t *)malloc(sizeof(int32_t) * (size_t)bs * R);
        float *news = (float *)malloc(sizeof(float) * (size_t)bs * R);
        int *newc = (int *)malloc(sizeof(int) * (size_t)bs);
#ifdef _OPENMP
#pragma omp parallel
#endif
        {
            scratch_t *S = scratch_new(ix);
            uint64_t *keys = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)beam);
            int32_t *cn = (int32_t *)malloc(sizeof(int32_t) * (size_t)beam);
            float *cs = (float *)malloc(sizeof(float) * (size_t)beam);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 8)
#endif
            for (int64_t b = 0; b < bs; b++) {
                int32_t p = order[done + b];
                int32_t cnt = 0;
                search_one(ix, S, vectors + (int64_t)p * dim, beam, beam, 0.f, 0.f, NULL, NULL, NULL, NULL, NULL, keys,
                           &cnt, entry, n);
                int nc = 0;
                for (int i = 0; i < cnt; i++) {
                    if (key_node(keys[i]) == p) continue;
                    cn[nc] = key_node(keys[i]);
                    cs[nc] = key_score(keys[i]);
                    nc++;
                }
                newc[b] = retain_diverse(vectors, dim, bsim, cn, cs, nc, R, alpha, newn + b * R, news + b * R);
            }
            free(keys);
            free(cn);
            free(cs);
            scratch_free(S);
        }
        /* commit out-edges, then backlinks in batch order */
        int32_t *touched = (int32_t *)malloc(sizeof(int32_t) * (size_t)bs * (R + 1));
        int nt = 0;
        for (int64_t b = 0; b < bs; b++) {
            int32_t p = order[done + b];
            inserted[p] = 1;
            if (acap[p] < newc[b] + 1) {
                acap[p] = cap + 8;
                adj[p] = (int32_t *)realloc(adj[p], sizeof(int32_t) * (size_t)acap[p]);
                ads[p] = (float *)realloc(ads[p], sizeof(float) * (size_t)acap[p]);
            }
            deg[p] = newc[b];
            memcpy(adj[p], newn + b * R, sizeof(int32_t) * (size_t)newc[b]);
            memcpy(ads[p], news + b * R, sizeof(float) * (size_t)newc[b]);
            touched[nt++] = p;
        }
        for (int64_t b = 0; b < bs; b++) {
            int32_t p = order[done + b];
            for (int j = 0; j < newc[b]; j++) {
                int32_t nb = newn[b * R + j];
                int dup = 0;
                for (int t = 0; t < deg[nb]; t++)
                    if (adj[nb][t] == p) {
                        dup = 1;
                        break;
                    }
                if (dup) continue;
                if (deg[nb] >= JVO_APPEND_CAP) continue; /* old ++ incoming is cut at 2048 entries (device scratch size) */
                if (deg[nb] + 1 > acap[nb]) {
                    acap[nb] = (deg[nb] + 1) * 2 + 8;
                    adj[nb] = (int32_t *)realloc(adj[nb], sizeof(int32_t) * (size_t)acap[nb]);
                    ads[nb] = (float *)realloc(ads[nb], sizeof(float) * (size_t)acap[nb]);
                }
                adj[nb][deg[nb]] = p;
                ads[nb][deg[nb]] = news[b * R + j]; /* symmetric similarity */
                deg[nb]++;
                touched[nt++] = nb;
            }
        }
        /* re-prune overflowing lists */
        for (int t = 0; t < nt; t++) {
            int32_t u = touched[t];
            if ((float)deg[u] > (float)R * overflow) {
                sort_by_key_desc(adj[u], ads[u], deg[u]);
                if (deg[u] > JVO_PRUNE_MAX_CANDS) deg[u] = JVO_PRUNE_MAX_CANDS; /* keep the best; same cap on the device */
                int32_t *tn = (int32_t *)malloc(sizeof(int32_t) * (size_t)R);
                float *ts = (float *)malloc(sizeof(float) * (size_t)R);
                int c = retain_diverse(vectors, dim, bsim, adj[u], ads[u], deg[u], R, alpha, tn, ts);
                memcpy(adj[u], tn, sizeof(int32_t) * (size_t)c);
                memcpy(ads[u], ts, sizeof(float) * (size_t)c);
                deg[u] = c;
                free(tn);
                free(ts);
            }
        }
        /* refresh the frozen view */
        for (int t = 0; t < nt; t++) {
            int32_t u = touched[t];
            for (int j = 0; j < Rb; j++) flat[(int64_t)u * Rb + j] = j < deg[u] ? adj[u][j] : -1;
        }
        free(touched);
        free(newn);
        free(news);
        free(newc);
        done += bs;
    }
    /* cleanup: enforce degree <= R */
    for (int64_t u = 0; u < n; u++) {
        if (deg[u] > R) {
            sort_by_key_desc(adj[u], ads[u], deg[u]);
            if (deg[u] > JVO_PRUNE_MAX_CANDS) deg[u] = JVO_PRUNE_MAX_CANDS;
            int32_t *tn = (int32_t *)malloc(sizeof(int32_t) * (size_t)R);
            float *ts = (float *)malloc(sizeof(float) * (size_t)R);
            int c = retain_diverse(vectors, dim, bsim, adj[u], ads[u], deg[u], R, alpha, tn, ts);
            memcpy(adj[u], tn, sizeof(int32_t) * (size_t)c);
            deg[u] = c;
            free(tn);
            free(ts);
        }
        for (int j = 0; j < R; j++) out_adj[u * R + j] = j < deg[u] ? adj[u][j] : -1;
    }
    for (int64_t u = 0; u < n; u++) {
        free(adj[u]);
        free(ads[u]);
    }
    free(adj);
    free(ads);
    free(deg);
    free(acap);
    free(inserted);
    free(flat);
    free(order);
    jvo_index_destroy(ix);
    return 0;
}

JVO_EXPORT int32_t jvo_graph_build(const float *vectors, int64_t n, int32_t dim, int32_t sim, int32_t R, int32_t beam,
                                   float overflow, float alpha, int32_t max_batch, float frac, int32_t *out_adj,
                                   int32_t *out_entry) {
    return graph_build_impl(vectors, n, dim, sim, R, beam, overflow, alpha, max_batch, frac, 0, NULL, 0, out_adj, out_entry);
}

JVO_EXPORT int32_t jvo_graph_extend(const float *vectors, int64_t n, int64_t n0, const int32_t *seed_adj, int32_t seed_entry,
                                    int32_t dim, int32_t sim, int32_t R, int32_t beam, float overflow, float alpha,
                                    int32_t max_batch, float frac, int32_t *out_adj) {
    int32_t entry = 0;
    if (n0 < 1 || n0 > n || seed_entry < 0 || seed_entry >= n0) return -1;
    return graph_build_impl(vectors, n, dim, sim, R, beam, overflow, alpha, max_batch, frac, n0, seed_adj, seed_entry, out_adj, &entry);
}

/* ------------------------------------------------------------------------------------------
 * FIXTURE: delete consolidation (GraphIndexBuilder.removeDeletedNodes behind builder.markNodeDeleted + cleanup,
 * JVectorWriter.java:1318-1327; the algorithm is FreshDiskANN section 4.2 / Algorithm 4 [M]):
 *   for every live node u with at least one deleted out-neighbour:
 *     C = live out-neighbours of u, then for every deleted out-neighbour j (row order) the live out-neighbours k != u of j
 *         (row order); the list is cut at 2048 entries
 *     scored with the exact build score against u, sorted best first (ties -> lower ordinal), duplicates dropped,
 *     cut at 1024, re-pruned with retainDiverse(R, alpha)
 *   rows of deleted nodes are emptied.  Ordinals are NOT compacted here (the writer does that).
 *   entry deleted -> the best-scoring live candidate of the same construction around the old entry, else the lowest
 *   live ordinal (jVector picks an approximate medoid; documented deviation).
 * ------------------------------------------------------------------------------------------ */
static int consolidate_candidates(const float *vectors, int dim, int bsim, const int32_t *adj, int R, const uint8_t *deleted,
                                  int32_t u, uint64_t *keys /* [JVO_APPEND_CAP] */) {
    int32_t ids[JVO_APPEND_CAP];
    int m = 0;
    const int32_t *row = adj + (int64_t)u * R;
    for (int j = 0; j < R && m < JVO_APPEND_CAP; j++)
        if (row[j] >= 0 && !deleted[row[j]]) ids[m++] = row[j];
    for (int j = 0; j < R; j++) {
        if (row[j] < 0 || !deleted[row[j]]) continue;
        const int32_t *r2 = adj + (int64_t)row[j] * R;
        for (int t = 0; t < R && m < JVO_APPEND_CAP; t++)
            if (r2[t] >= 0 && r2[t] != u && !deleted[r2[t]]) ids[m++] = r2[t];
    }
    const float *uv = vectors + (int64_t)u * dim;
    const float unorm = canon_dot(uv, uv, dim);
    for (int i = 0; i < m; i++) keys[i] = mk_key(exact_score(bsim, uv, unorm, vectors + (int64_t)ids[i] * dim, dim), ids[i]);
    qsort(keys, (size_t)m, sizeof(uint64_t), cmp_u64_desc);
    int w = 0;
    for (int i = 0; i < m; i++)
        if (i == 0 || keys[i] != keys[i - 1]) keys[w++] = keys[i];
    return w < JVO_PRUNE_MAX_CANDS ? w : JVO_PRUNE_MAX_CANDS;
}

JVO_EXPORT int32_t jvo_graph_remove_deleted(const float *vectors, int64_t n, int32_t dim, int32_t sim, int32_t R, float alpha,
                                            const int32_t *adj, const uint8_t *deleted, int32_t entry, int32_t *out_adj,
                                            int32_t *out_entry) {
    const int bsim = sim == JV_SIM_MIP ? JV_SIM_DOT : sim;
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
        uint64_t *keys = (uint64_t *)malloc(sizeof(uint64_t) * JVO_APPEND_CAP);
        int32_t *cn = (int32_t *)malloc(sizeof(int32_t) * JVO_PRUNE_MAX_CANDS);
        float *cs = (float *)malloc(sizeof(float) * JVO_PRUNE_MAX_CANDS);
        int32_t *tn = (int32_t *)malloc(sizeof(int32_t) * (size_t)R);
        float *ts = (float *)malloc(sizeof(float) * (size_t)R);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 64)
#endif
        for (int64_t u = 0; u < n; u++) {
            int32_t *o = out_adj + u * R;
            const int32_t *row = adj + u * R;
            if (deleted[u]) {
                for (int j = 0; j < R; j++) o[j] = -1;
                continue;
            }
            int hit = 0;
            for (int j = 0; j < R; j++)
                if (row[j] >= 0 && deleted[row[j]]) hit = 1;
            if (!hit) {
                for (int j = 0; j < R; j++) o[j] = row[j];
                continue;
            }
            int nc = consolidate_candidates(vectors, dim, bsim, adj, R, deleted, (int32_t)u, keys);
            for (int i = 0; i < nc; i++) {
                cn[i] = key_node(keys[i]);
                cs[i] = key_score(keys[i]);
            }
            int c = retain_diverse(vectors, dim, bsim, cn, cs, nc, R, alpha, tn, ts);
            for (int j = 0; j < R; j++) o[j] = j < c ? tn[j] : -1;
        }
        free(keys);
        free(cn);
        free(cs);
        free(tn);
        free(ts);
    }
    int32_t e = entry;
    if (entry >= 0 && entry < n && deleted[entry]) {
        uint64_t *keys = (uint64_t *)malloc(sizeof(uint64_t) * JVO_APPEND_CAP);
        int nc = consolidate_candidates(vectors, dim, bsim, adj, R, deleted, entry, keys);
        e = -1;
        if (nc > 0) e = key_node(keys[0]);
        for (int64_t i = 0; e < 0 && i < n; i++)
            if (!deleted[i]) e = (int32_t)i;
        free(keys);
    }
    *out_entry = e;
    return 0;
}

JVO_EXPORT int32_t jvo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
