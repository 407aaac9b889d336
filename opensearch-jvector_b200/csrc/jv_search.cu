// jv_search.cu — K1 (ADC table build) + K2 (graph traversal with PQ ADC scoring) / K4 (exact traversal).
//
// Replaces GraphSearcher.search(ssp, topK, rerankK, threshold, rerankFloor, bits) as called from
// JVectorReader.java:165-173 with the score provider of JVectorReader.java:352-365, for a BATCH of
// queries.  One CTA owns one query at a time (persistent grid, atomic work counter); the traversal
// is the reference's best-first order exactly:
//     pop best candidate -> stop if results are full and it is worse than the worst result
//     -> accepted && >= threshold: push to the bounded result list (rerankK)
//     -> every not-yet-visited neighbour is scored and pushed as a candidate
// Data layout in shared memory per CTA:
//     lut[M*K]   the query's ADC table (fp32, or fp16 with JV_INDEX_FLAG_LUT_F16)        (PQ only)
//     q[dim]     the query vector
//     cand[2][C] unexpanded candidates, sorted ascending by (score, ~node) key, double buffered
//     res[2][L]  expanded+accepted results (the reference's approxResults), ascending, L = rerankK
//     hash[H]    open-addressing visited set (ordinals)
// HBM reads per expansion: one adjacency row (R*4 B, one coalesced warp load) + one code row
// (M bytes, coalesced 4-byte-per-lane loads) per NEW neighbour — the algorithmic bytes of SURVEY 8(d).
#include "jv_internal.h"

#include "jv_search_common.cuh"

namespace jv {


template <bool PQ, typename LutT> __global__ void __launch_bounds__(kThreads) search_kernel(const SearchParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = p.cand_cap, L = p.L, H = 1 << p.hash_log2;

    // carve shared memory (all segments 16-byte aligned)
    unsigned char *sp = smem_raw;
    LutT *lut = reinterpret_cast<LutT *>(sp);
    if (PQ) sp += (((size_t)p.M * p.K * sizeof(LutT)) + 15) & ~(size_t)15;
    float *sq = reinterpret_cast<float *>(sp);
    sp += (((size_t)p.dim * 4) + 15) & ~(size_t)15;
    uint64_t *cand0 = reinterpret_cast<uint64_t *>(sp);
    sp += (size_t)C * 8;
    uint64_t *cand1 = reinterpret_cast<uint64_t *>(sp);
    sp += (size_t)C * 8;
    uint64_t *res0 = reinterpret_cast<uint64_t *>(sp);
    sp += (size_t)L * 8;
    uint64_t *res1 = reinterpret_cast<uint64_t *>(sp);
    sp += (size_t)L * 8;
    uint64_t *new_keys = reinterpret_cast<uint64_t *>(sp);
    sp += (size_t)kMaxR * 8;
    uint64_t *sorted_keys = reinterpret_cast<uint64_t *>(sp);
    sp += (size_t)kMaxR * 8;
    int32_t *new_ids = reinterpret_cast<int32_t *>(sp);
    sp += (size_t)kMaxR * 4;
    uint32_t *hash = reinterpret_cast<uint32_t *>(sp);

    __shared__ int s_query, s_nn, s_accept, s_hash_cnt, s_overflow;
    __shared__ float s_qnorm;

    const bool vec4 = (p.dim & 3) == 0 && (reinterpret_cast<uintptr_t>(p.queries) & 15) == 0;
    const uint32_t hmask = (uint32_t)H - 1u;
    const int hash_limit = H - (H >> 3) - 1; // stop inserting beyond 87.5 % load

    for (;;) {
        __syncthreads(); // previous query fully retired before shared state is reused
        if (tid == 0) s_query = atomicAdd(p.work_counter, 1);
        __syncthreads();
        const int qi = s_query;
        if (qi >= p.nq) break;

        // builder mode: the query is a stored vector addressed by ordinal
        const float *gq = p.query_ids ? p.vectors + (int64_t)__ldg(p.query_ids + qi) * p.dim : p.queries + (int64_t)qi * p.dim;
        for (int i = tid; i < p.dim; i += kThreads) sq[i] = __ldg(gq + i);
        for (int i = tid; i < H; i += kThreads) hash[i] = kEmpty;
        if (tid == 0) {
            s_hash_cnt = 0;
            s_overflow = 0;
        }
        __syncthreads();
        if (PQ) build_lut<LutT>(p, sq, lut, tid, kThreads);
        if (warp == 0) {
            float qn = jv_warp_reduce_pair<false, false>(sq, sq, p.dim, lane, (p.dim & 3) == 0);
            if (lane == 0) s_qnorm = qn;
        }
        __syncthreads();
        const float qnorm = s_qnorm;
        const uint64_t *accept = p.accept ? p.accept + (int64_t)qi * p.accept_stride : nullptr;

        // node scoring by one warp; result valid in all lanes
        auto score_node = [&](int32_t node) -> float {
            if (PQ) {
                float s = adc_warp_sum<LutT>(lut, p.K, p.M, p.codes + (int64_t)node * p.code_stride, lane);
                float nn = p.sim == JV_SIM_COSINE ? __ldg(p.node_norm + node) : 0.f;
                return adc_finish(p.sim, s, nn, qnorm);
            } else {
                const float *x = p.vectors + (int64_t)node * p.dim;
                float raw = p.sim == JV_SIM_EUCLIDEAN ? jv_warp_reduce_pair<true>(sq, x, p.dim, lane, vec4)
                                                      : jv_warp_reduce_pair<false>(sq, x, p.dim, lane, vec4);
                float xn = p.sim == JV_SIM_COSINE ? __ldg(p.vec_norm + node) : 0.f;
                return jv_finish_score(p.sim, raw, qnorm, xn) * p.mip_mul;
            }
        };

        // seed with the entry node (SURVEY A.1 "seed")
        int cand_n = 0, res_n = 0, cur_c = 0, cur_r = 0;
        int visited = 0, expanded = 0;
        const int32_t entry = p.entry;
        if (entry >= 0 && entry < p.n_limit) {
            if (warp == 0) {
                float s = score_node(entry);
                if (lane == 0) {
                    cand0[0] = jv_mk_key(s, entry);
                    uint32_t h = ((uint32_t)entry * 2654435761u) >> (32 - p.hash_log2);
                    hash[h] = (uint32_t)entry;
                    s_hash_cnt = 1;
                }
            }
            cand_n = 1;
            visited = 1;
        }
        __syncthreads();

        while (cand_n > 0) {
            uint64_t *cand = cur_c ? cand1 : cand0, *cand_next = cur_c ? cand0 : cand1;
            uint64_t *res = cur_r ? res1 : res0, *res_next = cur_r ? res0 : res1;
            const uint64_t top = cand[cand_n - 1];
            const uint32_t top_ord = (uint32_t)(top >> 32);
            if (res_n >= L && top_ord < (uint32_t)(res[0] >> 32)) break; // strictly worse than the worst result
            const int32_t node = jv_key_id(top);
            cand_n -= 1;
            expanded += 1;

            // ---- phase 1: warp 0 reads the adjacency row and filters it through the visited set;
            //               warp 1 evaluates the accept predicate of the popped node
            if (warp == 0) {
                int base_out = 0;
                const int32_t *row = p.adjacency + (int64_t)node * p.R;
                for (int j0 = 0; j0 < p.R; j0 += 32) {
                    const int j = j0 + lane;
                    int32_t nb = j < p.R ? __ldg(row + j) : -1;
                    bool fresh = false;
                    if (nb >= 0 && nb < p.n_limit) {
                        if (s_hash_cnt + base_out + 32 > hash_limit) {
                            s_overflow = 1; // table (nearly) full: treat as visited, search degrades gracefully
                        } else {
                            uint32_t h = ((uint32_t)nb * 2654435761u) >> (32 - p.hash_log2);
                            for (;;) {
                                uint32_t old = atomicCAS(&hash[h], kEmpty, (uint32_t)nb);
                                if (old == kEmpty) {
                                    fresh = true;
                                    break;
                                }
                                if (old == (uint32_t)nb) break;
                                h = (h + 1) & hmask;
                            }
                        }
                    }
                    const uint32_t ballot = __ballot_sync(JV_FULL_MASK, fresh);
                    if (fresh) new_ids[base_out + __popc(ballot & ((1u << lane) - 1u))] = nb;
                    base_out += __popc(ballot);
                }
                if (lane == 0) {
                    s_nn = base_out;
                    s_hash_cnt += base_out;
                }
            } else if (warp == 1 && lane == 0) {
                int32_t doc = p.ord_to_doc ? __ldg(p.ord_to_doc + node) : node;
                bool ok = jv_doc_accepted(accept, doc) && jv_key_score(top) >= p.threshold;
                s_accept = ok ? 1 : 0;
            }
            __syncthreads();
            const int nn = s_nn;
            const bool acc = s_accept != 0;

            // ---- phase 2a: bounded result insertion (approxResults.push)
            bool res_swapped = false;
            int res_n_next = res_n;
            if (acc) {
                if (res_n < L) {
                    const int pos = lower_bound_u64(res, res_n, top);
                    for (int i = tid; i < res_n; i += kThreads) {
                        uint64_t v = res[i];
                        res_next[i < pos ? i : i + 1] = v;
                    }
                    if (tid == 0) res_next[pos] = top;
                    res_n_next = res_n + 1;
                    res_swapped = true;
                } else if (top > res[0]) {
                    const int pos = lower_bound_u64(res, res_n, top); // >= 1
                    for (int i = tid + 1; i < res_n; i += kThreads) {
                        uint64_t v = res[i];
                        res_next[i < pos ? i - 1 : i] = v;
                    }
                    if (tid == 0) res_next[pos - 1] = top;
                    res_swapped = true;
                }
            }
            // ---- phase 2b: score the new neighbours, one warp per neighbour
            for (int i = warp; i < nn; i += kWarps) {
                const int32_t nb = new_ids[i];
                float s = score_node(nb);
                if (lane == 0) new_keys[i] = jv_mk_key(s, nb);
            }
            __syncthreads();
            if (res_swapped) cur_r ^= 1;
            res_n = res_n_next;

            // ---- phase 3: rank-sort the new keys (ascending)
            if (tid < nn) {
                const uint64_t k = new_keys[tid];
                int r = 0;
                for (int j = 0; j < nn; j++) r += new_keys[j] < k ? 1 : 0;
                sorted_keys[r] = k;
            }
            __syncthreads();

            // ---- phase 4: merge into the candidate array, keep the best C
            const int total = cand_n + nn;
            const int drop = total > C ? total - C : 0;
            for (int i = tid; i < cand_n; i += kThreads) {
                const uint64_t a = cand[i];
                const int pos = i + lower_bound_u64(sorted_keys, nn, a) - drop;
                if (pos >= 0) cand_next[pos] = a;
            }
            if (tid < nn) {
                const uint64_t b = sorted_keys[tid];
                const int pos = tid + lower_bound_u64(cand, cand_n, b) - drop;
                if (pos >= 0) cand_next[pos] = b;
            }
            cand_n = total - drop;
            cur_c ^= 1;
            visited += nn;
            __syncthreads();
        }

        // ---- emit the approximate result list, best first
        {
            const uint64_t *res = cur_r ? res1 : res0;
            uint64_t *out = p.approx_keys + (int64_t)qi * L;
            for (int i = tid; i < L; i += kThreads) out[i] = i < res_n ? res[res_n - 1 - i] : 0ull;
            if (tid == 0) {
                p.approx_count[qi] = res_n;
                if (p.stats) {
                    jv_query_stats st;
                    st.visited = visited;
                    st.expanded = expanded;
                    st.expanded_base = expanded;
                    st.reranked = 0; // filled by the rerank kernel
                    p.stats[qi] = st;
                }
                if (s_overflow) atomicAdd(p.dbg, 1);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------

template <bool PQ, typename LutT>
static int32_t launch_typed(jv_index *ix, SearchCtx *ctx, SearchParams &p, size_t fixed_bytes, size_t smem_limit) {
    // visited table: as large as shared memory allows (power of two), capped
    int hash_log2 = 16;
    while (hash_log2 > 8 && fixed_bytes + ((size_t)4 << hash_log2) > smem_limit) hash_log2--;
    if (fixed_bytes + ((size_t)4 << hash_log2) > smem_limit) {
        set_error("search: shared memory budget exceeded (%zu fixed bytes, limit %zu); use JV_INDEX_FLAG_LUT_F16 or a smaller rerank_k",
                  fixed_bytes, smem_limit);
        return JV_ERR_UNSUPPORTED;
    }
    // do not take more than needed: visits ~ expansions * new-neighbours ~ 0.7 * L * R without a filter;
    // a filter multiplies the expansions by ~1/selectivity, so keep everything shared memory offers then
    const int64_t want = (p.accept || p.threshold > 0.f) ? ((int64_t)1 << 30) : (int64_t)3 * p.L * (p.R > 0 ? p.R : 1);
    while (hash_log2 > 12 && ((int64_t)1 << (hash_log2 - 1)) >= want) hash_log2--;
    p.hash_log2 = hash_log2;
    const size_t smem = fixed_bytes + ((size_t)4 << hash_log2);
    auto kern = search_kernel<PQ, LutT>;
    JV_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    JV_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem));
    if (occ < 1) {
        set_error("search: kernel does not fit on an SM (smem %zu)", smem);
        return JV_ERR_UNSUPPORTED;
    }
    int grid = ix->sm_count * occ;
    if (grid > p.nq) grid = p.nq;
    if (grid < 1) grid = 1;
    JV_CUDA_TRY(cudaMemsetAsync(p.work_counter, 0, sizeof(int), ctx->stream));
    kern<<<grid, kThreads, smem, ctx->stream>>>(p);
    JV_CUDA_TRY(cudaGetLastError());
    return JV_OK;
}

int32_t launch_search(jv_index *ix, SearchCtx *ctx, const SearchLaunch &a, int *launches, bool *reranked) {
    if (reranked) *reranked = false;
    ctx->last_width = 0;
    ctx->last_kernel = JV_KERNEL_STRICT;
    if (a.nq <= 0) return JV_OK;
    JV_REQUIRE(ix->R <= kMaxR, "max_degree %d exceeds the supported %d", ix->R, kMaxR);
    SearchParams p;
    memset(&p, 0, sizeof(p));
    fill_params(ix, p);
    p.n_limit = a.n_limit > 0 ? a.n_limit : ix->n;
    if (a.entry_override >= 0) p.entry = a.entry_override;
    p.queries = a.d_queries;
    p.query_ids = a.d_query_ids;
    p.accept = a.d_accept;
    p.accept_stride = a.accept_stride_words;
    p.approx_keys = a.d_approx_keys;
    p.approx_count = a.d_approx_count;
    p.stats = a.d_stats;
    p.nq = a.nq;
    p.L = a.rerank_k;
    p.threshold = a.threshold;
    JV_TRY(ctx->counter.ensure(sizeof(int)));
    p.work_counter = ctx->counter.as<int>();
    p.dbg = ix->dbg.as<int>();
    // candidate capacity: without a filter only the best L unexpanded candidates can ever be popped
    // (DESIGN.md "bounded candidate list"); with a filter rejected nodes do not fill the results, keep more.
    int C = a.rerank_k < 64 ? 64 : a.rerank_k;
    if (a.d_accept || a.threshold > 0.f) C = C * 8 > 4096 ? (C > 4096 ? C : 4096) : C * 8; // results may never fill up
    p.cand_cap = C;
    const bool f16 = (ix->flags & JV_INDEX_FLAG_LUT_F16) != 0;
    // production path: unfiltered, threshold-free queries go to the fast kernel (jv_search_fast.cu); filters and
    // range thresholds need the reference's two-queue semantics and stay on the strict kernel below.
    const bool plain = a.expand_width >= 0 && a.d_accept == nullptr && !(a.threshold > 0.f);
    // filtered queries also take the 8-bit table path (accepted + rejected nodes in one list of 8 * rerankK entries); range
    // thresholds keep the strict kernel
    const bool filtered = a.expand_width >= 0 && a.d_accept != nullptr && !(a.threshold > 0.f);
    if ((plain || filtered) && (ix->flags & JV_INDEX_FLAG_LUT_U8) && a.d_query_ids == nullptr &&
        q8_search_supported(ix, a.rerank_k, ix->R, filtered))
        return launch_search_q8(ix, ctx, a, launches, reranked);
    if (plain && (!ix->has_pq || (p.M & 3) == 0)) {
        const int32_t fs = launch_search_fast(ix, ctx, p, a.expand_width == 0 ? 4 : a.expand_width, f16);
        if (fs == JV_OK && launches) *launches += 1;
        ctx->last_width = p.expand_width;
        ctx->last_kernel = JV_KERNEL_FAST;
        return fs;
    }
    size_t fixed = 0;
    if (ix->has_pq) fixed += (((size_t)p.M * p.K * (f16 ? 2 : 4)) + 15) & ~(size_t)15;
    fixed += (((size_t)p.dim * 4) + 15) & ~(size_t)15;
    fixed += (size_t)C * 16 + (size_t)p.L * 16 + (size_t)kMaxR * 20;
    const size_t limit = ix->smem_optin - 1024; // static __shared__ + reserve
    int32_t st;
    if (ix->has_pq)
        st = f16 ? launch_typed<true, __half>(ix, ctx, p, fixed, limit) : launch_typed<true, float>(ix, ctx, p, fixed, limit);
    else
        st = launch_typed<false, float>(ix, ctx, p, fixed, limit);
    if (st == JV_OK && launches) *launches += 1;
    return st;
}

// ---- test hooks: the ADC table and ADC scores through the same device functions ----------------
__global__ void lut_kernel(const SearchParams p, float *out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sq = reinterpret_cast<float *>(smem_raw);
    const int qi = blockIdx.x;
    for (int i = threadIdx.x; i < p.dim; i += blockDim.x) sq[i] = p.queries[(int64_t)qi * p.dim + i];
    __syncthreads();
    build_lut<float>(p, sq, out + (int64_t)qi * p.M * p.K, threadIdx.x, blockDim.x);
}

int32_t launch_pq_lut(jv_index *ix, cudaStream_t stream, const float *d_queries, int nq, float *d_lut) {
    JV_REQUIRE(ix->has_pq, "index has no PQ");
    SearchParams p;
    memset(&p, 0, sizeof(p));
    fill_params(ix, p);
    p.queries = d_queries;
    p.nq = nq;
    lut_kernel<<<nq, 256, (size_t)ix->dim * 4, stream>>>(p, d_lut);
    JV_CUDA_TRY(cudaGetLastError());
    return JV_OK;
}

// one CTA per query: table in GLOBAL memory (test hook, any M*K), warps score the requested nodes
__global__ void adc_pairs_kernel(const SearchParams p, const float *lut_all, const int32_t *nodes, int per_query, float *out) {
    const int qi = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const float *lut = lut_all + (int64_t)qi * p.M * p.K;
    const float *gq = p.queries + (int64_t)qi * p.dim;
    const float qnorm = jv_warp_reduce_pair<false>(gq, gq, p.dim, lane, (p.dim & 3) == 0);
    for (int j = warp; j < per_query; j += nwarps) {
        const int32_t node = nodes[(int64_t)qi * per_query + j];
        float s = adc_warp_sum<float>(lut, p.K, p.M, p.codes + (int64_t)node * p.code_stride, lane);
        float nn = p.sim == JV_SIM_COSINE ? p.node_norm[node] : 0.f;
        if (lane == 0) out[(int64_t)qi * per_query + j] = adc_finish(p.sim, s, nn, qnorm);
    }
}

int32_t launch_adc_pairs(jv_index *ix, cudaStream_t stream, const float *d_queries, int nq, const int32_t *d_nodes,
                         int per_query, float *d_out) {
    JV_REQUIRE(ix->has_pq, "index has no PQ");
    DevBuf lut;
    JV_TRY(lut.alloc((size_t)nq * ix->pq.M * ix->pq.K * 4));
    JV_TRY(launch_pq_lut(ix, stream, d_queries, nq, lut.as<float>()));
    SearchParams p;
    memset(&p, 0, sizeof(p));
    fill_params(ix, p);
    p.queries = d_queries;
    adc_pairs_kernel<<<nq, 128, 0, stream>>>(p, lut.as<float>(), d_nodes, per_query, d_out);
    JV_CUDA_TRY(cudaGetLastError());
    JV_CUDA_TRY(cudaStreamSynchronize(stream));
    return JV_OK;
}

}  // namespace jv
