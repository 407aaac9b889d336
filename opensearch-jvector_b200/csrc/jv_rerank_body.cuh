// jv_rerank_body.cuh — K3 for ONE query by one CTA of THREADS threads: gathers <= cnt inline fp32 vectors (dim*4 B each,
// coalesced 16-byte loads), scores them exactly (canonical reduction, bit-identical to the oracle), selects the top k by
// rank counting, maps ordinals to docs.  Shared by rerank_kernel (jv_rerank.cu) and the fused epilogue of the production
// traversal (jv_q8.cu).  Replaces the rerank step inside GraphSearcher.search fed by view.rerankerFor(q, sim)
// (JVectorReader.java:355) plus the ordinal->doc mapping and collector hand-off (JVectorReader.java:175-177).
#pragma once
#include "jv_internal.h"

namespace jv {

// sq: shared memory for the query (dim floats, 16-B aligned); keys: shared memory for cnt keys; akeys: the approximate
// list, best first (shared or global memory).  All THREADS threads of the CTA must call this (it synchronises).
// Returns (in every thread) the number of reranked candidates.
template <int THREADS>
__device__ __forceinline__ int rerank_query(const float *__restrict__ vectors, const float *__restrict__ vec_norm,
                                            const int32_t *__restrict__ ord_to_doc, int dim, int sim, int has_pq,
                                            const float *__restrict__ gq, bool vec4, int k, int cnt, float rerank_floor,
                                            const uint64_t *akeys, float *sq, uint64_t *keys, int32_t *out_doc, float *out_score,
                                            int32_t *out_count) {
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
                const float *x = vectors + (int64_t)node * dim;
                float raw = sim == JV_SIM_EUCLIDEAN ? jv_warp_reduce_pair<true>(sq, x, dim, lane, vec4)
                                                    : jv_warp_reduce_pair<false>(sq, x, dim, lane, vec4);
                float xn = sim == JV_SIM_COSINE ? __ldg(vec_norm + node) : 0.f;
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
