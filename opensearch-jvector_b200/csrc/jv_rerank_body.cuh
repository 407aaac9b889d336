// jv_rerank_body.cuh — K3 for ONE query by one CTA of THREADS threads: gathers <= cnt inline fp32 vectors (dim*4 B each,
// coalesced 16-byte loads), scores them exactly (canonical reduction, bit-identical to the oracle), selects the top k by
// rank counting, maps ordinals to docs.  Shared by rerank_kernel (jv_rerank.cu) and the fused epilogue of the production
// traversal (jv_q8.cu).  Replaces the rerank step inside GraphSearcher.search fed by view.rerankerFor(q, sim)
// (JVectorReader.java:355) plus the ordinal->doc mapping and collector hand-off (JVectorReader.java:175-177).
#pragma once
#include "jv_internal.h"

namespace jv {

// NVQ-inline vectors ("nvq+pq" segments): decoder = JVectorIndexQuantization.java:316-361, restated operation by operation
// (oracle: nvq_sub / logistic_nqt / logit_nqt).  bytes == nullptr: fp32 inline vectors.
struct NvqView {
    const uint8_t *bytes;   // [n][dim]
    const float *params;    // [n][m][4] growthRate, midpoint, minValue, maxValue
    const float *gmean;     // [dim]
    const int32_t *off;     // [m + 1] sub-vector offsets
    int m;
    float *xbuf;            // shared memory: one decoded vector per warp, [warps][dim]
    float *consts;          // shared memory: [warps][m][4] scale, bias, 1/alpha, mid
};

__device__ __forceinline__ float nvq_logistic(float value, float alpha, float x0) { // logisticNQT, :344-351
    float temp = __fmaf_rn(value, alpha, -__fmul_rn(alpha, x0));
    const int p = __double2int_rd((double)__fadd_rn(temp, 0.5f) + 0.5); // Math.round(temp + 0.5f)
    const int m = __float_as_int(__fmaf_rn(__fsub_rn(temp, (float)p), 0.5f, 1.0f));
    temp = __int_as_float(m + (int)((unsigned)p << 23));
    return __fdiv_rn(temp, __fadd_rn(temp, 1.0f));
}
__device__ __forceinline__ float nvq_logit(float scaled, float inv_alpha, float x0) { // logitNQT, :354-361
    const float z = __fdiv_rn(scaled, __fsub_rn(1.0f, scaled));
    const int temp = __float_as_int(z);
    const int e = temp & 0x7f800000;
    const float p = (float)((e >> 23) - 128);
    const float m = __int_as_float((temp & 0x007fffff) + 0x3f800000);
    return __fadd_rn(__fmul_rn(__fadd_rn(m, p), inv_alpha), x0);
}
// one warp decodes the inline vector of `node` into its xbuf slice (nvqDequantize, :316-341)
__device__ __forceinline__ const float *nvq_decode_warp(const NvqView &v, int32_t node, int dim, int warp, int lane) {
    float *c = v.consts + (size_t)warp * v.m * 4;
    float *x = v.xbuf + (size_t)warp * ((dim + 3) & ~3); // 16-B aligned slices
    __syncwarp(); // the previous vector of this warp has been consumed
    if (lane < v.m) {
        const float *prm = v.params + ((int64_t)node * v.m + lane) * 4;
        const float growth = __ldg(prm), midpoint = __ldg(prm + 1), lo = __ldg(prm + 2), hi = __ldg(prm + 3);
        const float delta = __fsub_rn(hi, lo);
        const float sgr = __fdiv_rn(growth, delta);
        const float mid = __fmul_rn(midpoint, delta);
        const float bias = nvq_logistic(lo, sgr, mid);
        c[lane * 4 + 0] = __fdiv_rn(__fsub_rn(nvq_logistic(hi, sgr, mid), bias), 255.0f);
        c[lane * 4 + 1] = bias;
        c[lane * 4 + 2] = __fdiv_rn(1.0f, sgr);
        c[lane * 4 + 3] = mid;
    }
    __syncwarp();
    const uint8_t *b = v.bytes + (int64_t)node * dim;
    for (int i = lane; i < dim; i += 32) {
        int sub = 0;
        while (sub + 1 < v.m && i >= __ldg(v.off + sub + 1)) sub++;
        const float y = nvq_logit(__fmaf_rn((float)__ldg(b + i), c[sub * 4], c[sub * 4 + 1]), c[sub * 4 + 2], c[sub * 4 + 3]);
        x[i] = __fadd_rn(y, __ldg(v.gmean + i));
    }
    __syncwarp();
    return x;
}

// sq: shared memory for the query (dim floats, 16-B aligned); keys: shared memory for cnt keys; akeys: the approximate
// list, best first (shared or global memory).  All THREADS threads of the CTA must call this (it synchronises).
// Returns (in every thread) the number of reranked candidates.
// Row map of a de-duplicated batch (rerank vectors in host memory, jv_rerank.cu): the rows every query of the batch needs were
// gathered once into a dense staging array in HBM; row(node) = base[node / 32] + popc(bits of the word below the node's).
// bitmap == nullptr: the identity (vectors indexed by ordinal).
struct RowMap {
    const uint32_t *bitmap;
    const int32_t *base;
};
__device__ __forceinline__ int64_t row_of(const RowMap &m, int32_t node) {
    if (m.bitmap == nullptr) return node;
    const uint32_t w = __ldg(m.bitmap + (node >> 5));
    return (int64_t)__ldg(m.base + (node >> 5)) + __popc(w & ((1u << (node & 31)) - 1u));
}

template <int THREADS>
__device__ __forceinline__ int rerank_query(const float *__restrict__ vectors, const float *__restrict__ vec_norm,
                                            const int32_t *__restrict__ ord_to_doc, int dim, int sim, int has_pq,
                                            const float *__restrict__ gq, bool vec4, int k, int cnt, float rerank_floor,
                                            const uint64_t *akeys, float *sq, uint64_t *keys, int32_t *out_doc, float *out_score,
                                            int32_t *out_count, const NvqView nvq = NvqView{nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr},
                                            const RowMap rows = RowMap{nullptr, nullptr}) {
    __shared__ float s_qnorm;
    __shared__ int s_valid, s_reranked;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_valid = 0, s_reranked = 0;
    if (has_pq) {
        for (int i = tid; i < dim; i += THREADS) sq[i] = __ldg(gq + i);
        __syncthreads();
        if (warp == 0) {
            float qn = jv_warp_reduce_pair<false, false>(sq, sq, dim, lane, (dim & 3) == 0);
            if (lane == 0) s_qnorm = qn;
        }
    }
    __syncthreads();
    int reranked = 0;
    for (int j = warp; j < cnt; j += THREADS / 32) {
        const uint64_t ak = akeys[j];
        const int32_t node = jv_key_id(ak);
        float s = jv_key_score(ak);
        uint64_t key = 0ull;
        const int32_t doc = ord_to_doc ? __ldg(ord_to_doc + node) : node;
        if (has_pq) {
            if (s >= rerank_floor) {
                float raw, xn = 0.f;
                if (nvq.bytes != nullptr) { // score the dequantised inline vector (shared memory), same canonical reduction
                    const float *x = nvq_decode_warp(nvq, node, dim, warp, lane);
                    const bool v4 = (dim & 3) == 0;
                    raw = sim == JV_SIM_EUCLIDEAN ? jv_warp_reduce_pair<true, false>(sq, x, dim, lane, v4)
                                                  : jv_warp_reduce_pair<false, false>(sq, x, dim, lane, v4);
                    if (sim == JV_SIM_COSINE) xn = jv_warp_reduce_pair<false, false>(x, x, dim, lane, v4);
                } else {
                    const float *x = vectors + row_of(rows, node) * dim;
                    raw = sim == JV_SIM_EUCLIDEAN ? jv_warp_reduce_pair<true>(sq, x, dim, lane, vec4)
                                                  : jv_warp_reduce_pair<false>(sq, x, dim, lane, vec4);
                    if (sim == JV_SIM_COSINE) xn = __ldg(vec_norm + node);
                }
                s = jv_finish_score(sim, raw, s_qnorm, xn); // the PQ reranker is NOT x2-wrapped (JVectorReader.java:352-356)
                key = jv_mk_key(s, doc);
                reranked++;
            }
        } else {
            key = jv_mk_key(s, doc); // traversal scores are already exact (and MIP-doubled)
        }
        if (lane == 0) keys[j] = key;
    }
    __syncthreads();
    // rank selection: rank = number of strictly better keys (keys are unique: doc ids differ)
    int valid_local = 0;
    for (int j = tid; j < cnt; j += THREADS) {
        const uint64_t my = keys[j];
        if (my == 0ull) continue;
        valid_local++;
        int rank = 0;
        for (int t = 0; t < cnt; t++) rank += keys[t] > my ? 1 : 0;
        if (rank < k) {
            out_doc[rank] = jv_key_id(my);
            out_score[rank] = jv_key_score(my);
        }
    }
    if (valid_local) atomicAdd(&s_valid, valid_local);
    if (lane == 0 && reranked) atomicAdd(&s_reranked, reranked); // `reranked` is warp-uniform
    __syncthreads();
    const int nout = s_valid < k ? s_valid : k;
    for (int j = nout + tid; j < k; j += THREADS) {
        out_doc[j] = -1;
        out_score[j] = 0.f;
    }
    if (tid == 0) *out_count = nout;
    return s_reranked;
}

}  // namespace jv
