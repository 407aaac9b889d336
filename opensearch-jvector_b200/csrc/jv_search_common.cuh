// jv_search_common.cuh — device helpers shared by the traversal kernels (strict: jv_search.cu, fast: jv_search_fast.cu).
#pragma once
#include "jv_internal.h"

#include <cuda_fp16.h>

namespace jv {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxR = 128; // max graph degree supported by the per-step scratch
constexpr uint32_t kEmpty = 0xffffffffu;

struct SearchParams {
    // index
    const int32_t *adjacency;
    const float *vectors;
    const float *vec_norm;
    const int32_t *ord_to_doc;
    const uint8_t *codes;
    const float *codebooks;
    const __half *codebooks_h; // fp16 copy of the codebooks (fast kernel with the fp16 table): half the L2 bytes of the table build
    const float *gcent;
    const int32_t *pq_size, *pq_off, *pq_cboff;
    const float *node_norm;
    int64_t n, n_limit;
    int R, entry, sim, dim, M, K, code_stride, sub_uniform;
    float mip_mul;
    // batch
    const float *queries;
    const int32_t *query_ids; // optional: query i is vectors[query_ids[i]] (graph builder)
    const uint64_t *accept;
    int64_t accept_stride;
    uint64_t *approx_keys;
    int32_t *approx_count;
    jv_query_stats *stats;
    int *work_counter;
    int *dbg;
    int nq, L;
    float threshold;
    // shared-memory geometry
    int cand_cap, hash_log2;
    // fast kernel
    int expand_width, list_cap, adc_lanes_log2;
};

template <typename T> __device__ __forceinline__ float lut_get(const T *lut, int i);
template <> __device__ __forceinline__ float lut_get<float>(const float *lut, int i) { return lut[i]; }
template <> __device__ __forceinline__ float lut_get<__half>(const __half *lut, int i) { return __half2float(lut[i]); }
template <typename T> __device__ __forceinline__ void lut_put(T *lut, int i, float v);
template <> __device__ __forceinline__ void lut_put<float>(float *lut, int i, float v) { lut[i] = v; }
template <> __device__ __forceinline__ void lut_put<__half>(__half *lut, int i, float v) { lut[i] = __float2half_rn(v); }

// K1: lut[m][c] = q_m . C_m[c]  (DOT/COSINE/MIP)  or  ||(q-g)_m - C_m[c]||^2 (EUCLIDEAN); sequential fmaf over
// the sub-vector, identical to oracle pq_build_lut (PQVectors.precomputedScoreFunctionFor, JVectorReader.java:354).
// Uniform sub-vector sizes S in {2,4,8}: the codebook (dim*K*4 B; 786 KB at 768-d) streams from L2 with U
// independent vector loads per thread in flight; entries idx = m*K + c are contiguous S-float centroids.
template <typename LutT, int S>
__device__ __forceinline__ void build_lut_uniform(const SearchParams &p, const float *sq, LutT *lut, int tid, int nthreads) {
    const int total = p.M * p.K;
    const bool l2 = p.sim == JV_SIM_EUCLIDEAN;
    constexpr int U = S <= 4 ? 8 : 4;
    constexpr int V = S == 2 ? 2 : 4; // floats per vector load
    constexpr int NV = S / V;         // vector loads per centroid
    for (int base = tid; base < total; base += nthreads * U) {
        float cc[U][S];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int idx = base + u * nthreads;
#pragma unroll
            for (int v = 0; v < NV; v++) {
                if (V == 2) {
                    const float2 t = idx < total ? __ldg(reinterpret_cast<const float2 *>(p.codebooks) + (size_t)idx * NV + v)
                                                 : make_float2(0.f, 0.f);
                    cc[u][v * V] = t.x, cc[u][v * V + 1] = t.y;
                } else {
                    const float4 t = idx < total ? __ldg(reinterpret_cast<const float4 *>(p.codebooks) + (size_t)idx * NV + v)
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
                    cc[u][v * V] = t.x, cc[u][v * V + 1] = t.y, cc[u][v * V + 2] = t.z, cc[u][v * V + 3] = t.w;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int idx = base + u * nthreads;
            if (idx >= total) break;
            const int m = idx / p.K;
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < S; j++) {
                float qv = sq[S * m + j];
                if (l2) {
                    if (p.gcent) qv -= __ldg(p.gcent + S * m + j);
                    const float d = qv - cc[u][j];
                    acc = __fmaf_rn(d, d, acc);
                } else {
                    acc = __fmaf_rn(qv, cc[u][j], acc);
                }
            }
            lut_put(lut, idx, acc);
        }
    }
}

// K == 256 with a 256-thread CTA (every production config: K = min(256, n)): thread c owns centroid c of every
// subspace, so there is no index arithmetic at all — per entry one coalesced vector load (16 B * 32 lanes), one
// broadcast read of q_m from shared memory, S FMAs and one store.  Same fmaf order as build_lut_uniform.
template <typename LutT, int S>
__device__ __forceinline__ void build_lut_k256(const SearchParams &p, const float *sq, LutT *lut, int tid) {
    const bool l2 = p.sim == JV_SIM_EUCLIDEAN;
    // the loop is bound by the L2 round trip (~1.6 k cycles per iteration measured): keep 64 registers of centroids
    // (16 loads of 16 B at sub-dim 4) in flight per thread
    constexpr int U = S <= 4 ? 8 : 4;
    constexpr int V = S == 2 ? 2 : 4;
    constexpr int NV = S / V;
    const int M = p.M;
    for (int m0 = 0; m0 < M; m0 += U) {
        float cc[U][S];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int m = m0 + u;
            const size_t e = ((size_t)m * 256 + tid) * NV;
#pragma unroll
            for (int v = 0; v < NV; v++) {
                if (V == 2) {
                    const float2 t = m < M ? __ldg(reinterpret_cast<const float2 *>(p.codebooks) + e + v) : make_float2(0.f, 0.f);
                    cc[u][v * V] = t.x, cc[u][v * V + 1] = t.y;
                } else {
                    const float4 t = m < M ? __ldg(reinterpret_cast<const float4 *>(p.codebooks) + e + v) : make_float4(0.f, 0.f, 0.f, 0.f);
                    cc[u][v * V] = t.x, cc[u][v * V + 1] = t.y, cc[u][v * V + 2] = t.z, cc[u][v * V + 3] = t.w;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int m = m0 + u;
            if (m >= M) break;
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < S; j++) {
                float qv = sq[S * m + j]; // warp-wide broadcast
                if (l2) {
                    if (p.gcent) qv -= __ldg(p.gcent + S * m + j);
                    const float d = qv - cc[u][j];
                    acc = __fmaf_rn(d, d, acc);
                } else {
                    acc = __fmaf_rn(qv, cc[u][j], acc);
                }
            }
            lut_put(lut, m * 256 + tid, acc);
        }
    }
}

// fp16-table variant of build_lut_k256 reading the fp16 codebook copy: the table build is bound by the bytes streamed from
// L2 (786 KB of fp32 centroids per query at 768-d), so halving them halves its time.  Entries are rounded to fp16 anyway
// and only steer the traversal; products and sums stay in fp32.
template <int S>
__device__ __forceinline__ void build_lut_k256_h(const SearchParams &p, const float *sq, __half *lut, int tid) {
    const bool l2 = p.sim == JV_SIM_EUCLIDEAN;
    constexpr int U = S >= 8 ? 8 : 16; // <= 64 registers of centroids in flight per thread
    const int M = p.M;
    for (int m0 = 0; m0 < M; m0 += U) {
        float cc[U][S];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int m = m0 + u;
            const __half *src = p.codebooks_h + ((size_t)m * 256 + tid) * S;
            if (m < M) {
                if (S == 2) {
                    const float2 f = __half22float2(__ldg(reinterpret_cast<const __half2 *>(src)));
                    cc[u][0] = f.x, cc[u][1] = f.y;
                } else if (S == 4) {
                    const uint2 raw = __ldg(reinterpret_cast<const uint2 *>(src));
                    const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
                    const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
                    cc[u][0] = a.x, cc[u][1] = a.y, cc[u][2] = b.x, cc[u][3] = b.y;
                } else {
                    const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(src));
                    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                    for (int h = 0; h < 4; h++) {
                        const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[h]));
                        cc[u][2 * h] = f.x, cc[u][2 * h + 1] = f.y;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < S; j++) cc[u][j] = 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int m = m0 + u;
            if (m >= M) break;
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < S; j++) {
                float qv = sq[S * m + j];
                if (l2) {
                    if (p.gcent) qv -= __ldg(p.gcent + S * m + j);
                    const float d = qv - cc[u][j];
                    acc = __fmaf_rn(d, d, acc);
                } else {
                    acc = __fmaf_rn(qv, cc[u][j], acc);
                }
            }
            lut[m * 256 + tid] = __float2half_rn(acc);
        }
    }
}

template <typename LutT> struct IsHalf { static constexpr bool value = false; };
template <> struct IsHalf<__half> { static constexpr bool value = true; };

template <typename LutT>
__device__ __forceinline__ void build_lut(const SearchParams &p, const float *sq, LutT *lut, int tid, int nthreads) {
    const int total = p.M * p.K;
    const bool l2 = p.sim == JV_SIM_EUCLIDEAN;
    if constexpr (IsHalf<LutT>::value) {
        if (p.codebooks_h && p.sub_uniform && p.K == 256 && nthreads == 256) {
            const int S = p.dim / p.M;
            if (S == 4) return build_lut_k256_h<4>(p, sq, lut, tid);
            if (S == 2) return build_lut_k256_h<2>(p, sq, lut, tid);
            if (S == 8) return build_lut_k256_h<8>(p, sq, lut, tid);
        }
    }
    if (p.sub_uniform && p.K == 256 && nthreads == 256) {
        const int S = p.dim / p.M;
        if (S == 4) return build_lut_k256<LutT, 4>(p, sq, lut, tid);
        if (S == 2) return build_lut_k256<LutT, 2>(p, sq, lut, tid);
        if (S == 8) return build_lut_k256<LutT, 8>(p, sq, lut, tid);
    }
    if (p.sub_uniform) {
        const int S = p.dim / p.M;
        if (S == 4) return build_lut_uniform<LutT, 4>(p, sq, lut, tid, nthreads);
        if (S == 2) return build_lut_uniform<LutT, 2>(p, sq, lut, tid, nthreads);
        if (S == 8) return build_lut_uniform<LutT, 8>(p, sq, lut, tid, nthreads);
    }
    for (int idx = tid; idx < total; idx += nthreads) {
        const int m = idx / p.K, c = idx - m * p.K;
        const int len = __ldg(p.pq_size + m), off = __ldg(p.pq_off + m);
        const float *cv = p.codebooks + __ldg(p.pq_cboff + m) + (int64_t)c * len;
        float acc = 0.f;
        for (int j = 0; j < len; j++) {
            float qv = sq[off + j];
            if (l2) {
                if (p.gcent) qv -= __ldg(p.gcent + off + j);
                float d = qv - __ldg(cv + j);
                acc = __fmaf_rn(d, d, acc);
            } else {
                acc = __fmaf_rn(qv, __ldg(cv + j), acc);
            }
        }
        lut_put(lut, idx, acc);
    }
}

// a4: decoder mapping of the summed partials (PQDecoder.*Decoder)
__device__ __forceinline__ float adc_finish(int sim, float s, float node_norm, float qnorm) {
    if (sim == JV_SIM_EUCLIDEAN) return __fdiv_rn(1.0f, __fadd_rn(1.0f, s));
    if (sim == JV_SIM_COSINE) return __fmul_rn(__fadd_rn(1.0f, __fdiv_rn(s, __fsqrt_rn(__fmul_rn(node_norm, qnorm)))), 0.5f);
    return __fmul_rn(__fadd_rn(1.0f, s), 0.5f);
}

// one warp sums the table entries selected by one code row (assembleAndSum).  Lane l owns the 4-byte code
// words l, l+32, .. and adds their subspaces in increasing m; the 32 lane partials go through the halving
// tree.  This is "order 1 / warp32" of oracle adc_sum, so fp32 traversals match the oracle bit for bit.
template <typename LutT>
__device__ __forceinline__ float adc_warp_sum(const LutT *lut, int K, int M, const uint8_t *row, int lane) {
    const uint32_t *row32 = reinterpret_cast<const uint32_t *>(row);
    const int nwords = (M + 3) >> 2;
    float s = 0.f;
    for (int w = lane; w < nwords; w += 32) {
        const uint32_t cw = __ldg(row32 + w);
        const int m0 = w * 4;
#pragma unroll
        for (int b = 0; b < 4; b++)
            if (m0 + b < M) s = __fadd_rn(s, lut_get(lut, (m0 + b) * K + (int)((cw >> (8 * b)) & 0xffu)));
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s = __fadd_rn(s, __shfl_xor_sync(JV_FULL_MASK, s, off));
    return s;
}

__device__ __forceinline__ int lower_bound_u64(const uint64_t *a, int n, uint64_t key) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] < key)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

static inline void fill_params(const jv_index *ix, SearchParams &p) {
    p.adjacency = ix->adjacency.as<int32_t>();
    p.vectors = ix->vectors_dev;
    p.vec_norm = ix->vec_norm.as<float>();
    p.ord_to_doc = ix->ord_to_doc.as<int32_t>();
    p.codes = ix->codes.as<uint8_t>();
    p.codebooks = ix->codebooks.as<float>();
    p.codebooks_h = nullptr; // only the fast kernel opts in (launch_search_fast)
    p.gcent = ix->gcent.as<float>();
    p.pq_size = ix->pq_size.as<int32_t>();
    p.pq_off = ix->pq_off.as<int32_t>();
    p.pq_cboff = ix->pq_cboff.as<int32_t>();
    p.node_norm = ix->node_norm.as<float>();
    p.n = ix->n;
    p.R = ix->R;
    p.entry = ix->entry;
    p.sim = ix->sim;
    p.dim = ix->dim;
    p.M = ix->pq.M;
    p.K = ix->pq.K;
    p.code_stride = ix->code_stride;
    p.sub_uniform = ix->pq.uniform ? 1 : 0;
    // wrapExactScoreFunction, JVectorReader.java:220-239 — only the plain un-quantised branch (:359-363); the NVQ-only branch (:357-358) is not wrapped
    p.mip_mul = (ix->sim == JV_SIM_MIP && !ix->has_pq && !ix->nvq_only) ? 2.0f : 1.0f;
}

// the fast kernel's launcher (jv_search_fast.cu)
int32_t launch_search_fast(jv_index *ix, SearchCtx *ctx, SearchParams &p, int expand_width, bool f16);

}  // namespace jv
