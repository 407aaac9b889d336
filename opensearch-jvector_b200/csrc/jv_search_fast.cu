// jv_search_fast.cu — the production traversal kernel (K1 + K2 / K4) for unfiltered, threshold-free queries.
//
// Same search as jv_search.cu (GraphSearcher.search, JVectorReader.java:165-173) restructured for the GPU:
//   * ONE sorted list of the best L = rerankK visited nodes, each flagged expanded / unexpanded, replaces the
//     candidate heap + result heap.  Without a filter a candidate ranked below L visited nodes can never be
//     popped by the reference loop (everything better is popped first and fills the result heap), so the bounded
//     list explores exactly the same nodes for expand_width = 1 (up to exact score ties).
//   * expand_width E > 1 expands the E best unexpanded entries per step (a superset exploration: recall >= the
//     strict order's, fewer dependent round trips to HBM, E*R neighbour rows in flight at once).
//   * the visited set is a small direct-mapped filter in shared memory: it only avoids re-scoring.  Correctness does
//     not depend on it — a re-scored node is either already in the list (dropped as a duplicate at merge time) or
//     scores below the list's worst entry (dropped again) — so it never overflows and needs no atomics.
//   * the ADC table may be held in fp16 (JV_INDEX_FLAG_LUT_F16): 96 KB instead of 192 KB at M=192,K=256, which lets two
//     CTAs share an SM.  ADC scores only steer the traversal; returned scores come from the exact rerank (K3).
// Shared memory per CTA: lut[M*K] | q[dim] | list[2][L] | surv[256] | flags | ids[256] | sel[8] | filter[H].
#include "jv_search_common.cuh"

namespace jv {

constexpr int kFThreads = 256;
constexpr int kFWarps = kFThreads / 32;
constexpr int kMaxE = 8;
constexpr int kMaxNew = 256; // E * R neighbour slots per step

// list key: score ordinal | (0x7fffffff - node) << 1 | unexpanded
__device__ __forceinline__ uint64_t fkey_make(float score, int32_t node) {
    return ((uint64_t)jv_f2ord(score) << 32) | ((uint64_t)(uint32_t)(0x7fffffff - node) << 1) | 1ull;
}
__device__ __forceinline__ int32_t fkey_node(uint64_t k) { return 0x7fffffff - (int32_t)((k >> 1) & 0x7fffffffu); }
__device__ __forceinline__ float fkey_score(uint64_t k) { return jv_ord2f((uint32_t)(k >> 32)); }

// number of list entries strictly better than `a` (= key >> 1); list sorted descending
__device__ __forceinline__ int count_better(const uint64_t *list, int n, uint64_t a) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((list[mid] >> 1) > a)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

template <bool PQ, typename LutT>
__global__ void __launch_bounds__(kFThreads, 2) fast_search_kernel(const SearchParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = p.L, E = p.expand_width, H = 1 << p.hash_log2;

    unsigned char *sp = smem_raw;
    LutT *lut = reinterpret_cast<LutT *>(sp);
    if (PQ) sp += (((size_t)p.M * p.K * sizeof(LutT)) + 15) & ~(size_t)15;
    float *sq = reinterpret_cast<float *>(sp);
    sp += (((size_t)p.dim * 4) + 15) & ~(size_t)15;
    uint64_t *list0 = reinterpret_cast<uint64_t *>(sp);
    sp += (size_t)L * 8;
    uint64_t *list1 = reinterpret_cast<uint64_t *>(sp);
    sp += (size_t)L * 8;
    uint64_t *surv = reinterpret_cast<uint64_t *>(sp);
    sp += (size_t)kMaxNew * 8;
    int32_t *nb_ids = reinterpret_cast<int32_t *>(sp);
    sp += (size_t)kMaxNew * 4;
    uint32_t *filter = reinterpret_cast<uint32_t *>(sp);
    sp += (size_t)H * 4;
    uint8_t *sflag = reinterpret_cast<uint8_t *>(sp);

    __shared__ int s_query, s_nn, s_ns, s_warp_cnt[kFWarps], s_sel[kMaxE];
    __shared__ float s_qnorm;

    const bool vec4 = (p.dim & 3) == 0 && (p.query_ids != nullptr || (reinterpret_cast<uintptr_t>(p.queries) & 15) == 0);
    const int hshift = 32 - p.hash_log2;

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            s_query = atomicAdd(p.work_counter, 1);
            s_nn = 0;
            s_ns = 0;
        }
        __syncthreads();
        const int qi = s_query;
        if (qi >= p.nq) break;
        const float *gq = p.query_ids ? p.vectors + (int64_t)__ldg(p.query_ids + qi) * p.dim : p.queries + (int64_t)qi * p.dim;
        for (int i = tid; i < p.dim; i += kFThreads) sq[i] = __ldg(gq + i);
        for (int i = tid; i < H; i += kFThreads) filter[i] = kEmpty;
        __syncthreads();
        if (PQ) build_lut<LutT>(p, sq, lut, tid, kFThreads);
        if (warp == 0) {
            const float qn = jv_warp_reduce_pair<false>(sq, gq, p.dim, lane, vec4);
            if (lane == 0) s_qnorm = qn;
        }
        __syncthreads();
        const float qnorm = s_qnorm;

        auto score_node = [&](int32_t node) -> float { // one warp; result valid in all lanes
            if (PQ) {
                const float s = adc_warp_sum<LutT>(lut, p.K, p.M, p.codes + (int64_t)node * p.code_stride, lane);
                const float nn = p.sim == JV_SIM_COSINE ? __ldg(p.node_norm + node) : 0.f;
                return adc_finish(p.sim, s, nn, qnorm);
            } else {
                const float *x = p.vectors + (int64_t)node * p.dim;
                const float raw = p.sim == JV_SIM_EUCLIDEAN ? jv_warp_reduce_pair<true>(sq, x, p.dim, lane, vec4)
                                                            : jv_warp_reduce_pair<false>(sq, x, p.dim, lane, vec4);
                const float xn = p.sim == JV_SIM_COSINE ? __ldg(p.vec_norm + node) : 0.f;
                return jv_finish_score(p.sim, raw, qnorm, xn) * p.mip_mul;
            }
        };

        int n = 0, cur = 0, visited = 0, expanded = 0;
        if (p.entry >= 0 && p.entry < p.n_limit) {
            if (warp == 0) {
                const float s = score_node(p.entry);
                if (lane == 0) {
                    list0[0] = fkey_make(s, p.entry);
                    filter[((uint32_t)p.entry * 2654435761u) >> hshift] = (uint32_t)p.entry;
                }
            }
            n = 1;
            visited = 1;
        }
        __syncthreads();

        while (n > 0) {
            uint64_t *list = cur ? list1 : list0, *out = cur ? list0 : list1;
            // ---- (a) pick the E best unexpanded entries (list order = best first) and mark them expanded
            int nsel = 0;
            for (int c0 = 0; c0 < n && nsel < E; c0 += kFThreads) {
                const int i = c0 + tid;
                const bool un = i < n && (list[i] & 1ull);
                const uint32_t ballot = __ballot_sync(JV_FULL_MASK, un);
                if (lane == 0) s_warp_cnt[warp] = __popc(ballot);
                __syncthreads();
                int before = nsel, total = nsel;
                for (int w = 0; w < kFWarps; w++) {
                    const int c = s_warp_cnt[w];
                    if (w < warp) before += c;
                    total += c;
                }
                const int rank = before + __popc(ballot & ((1u << lane) - 1u));
                if (un && rank < E) {
                    s_sel[rank] = fkey_node(list[i]);
                    list[i] &= ~1ull;
                }
                nsel = total < E ? total : E;
                __syncthreads();
            }
            if (nsel == 0) break;

            // ---- (b) neighbour rows of the selected nodes -> visited filter -> compacted id list
            {
                const int R = p.R;
                const int t = tid;
                int32_t nb = -1;
                if (t < nsel * R) {
                    const int ci = t / R, j = t - ci * R;
                    nb = __ldg(p.adjacency + (int64_t)s_sel[ci] * R + j);
                }
                bool fresh = false;
                if (nb >= 0 && nb < p.n_limit) {
                    const uint32_t h = ((uint32_t)nb * 2654435761u) >> hshift;
                    if (filter[h] != (uint32_t)nb) {
                        filter[h] = (uint32_t)nb;
                        fresh = true;
                    }
                }
                const uint32_t ballot = __ballot_sync(JV_FULL_MASK, fresh);
                int base = 0;
                if (lane == 0 && ballot) base = atomicAdd(&s_nn, __popc(ballot));
                base = __shfl_sync(JV_FULL_MASK, base, 0);
                if (fresh) nb_ids[base + __popc(ballot & ((1u << lane) - 1u))] = nb;
            }
            __syncthreads();
            const int nn = s_nn;
            expanded += nsel;
            visited += nn;

            // ---- (c) score the gathered neighbours, one warp each; keep those that can enter the list
            {
                const uint64_t worst = n >= L ? (list[L - 1] >> 1) : 0ull;
                if (PQ) {
                    // 4 code rows per warp in flight: all loads are issued before the first table lookup
                    const int nwords = (p.M + 3) >> 2;
                    for (int i0 = warp; i0 < nn; i0 += kFWarps * 4) {
                        uint32_t cw[4][2];
                        int32_t nbs[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int i = i0 + u * kFWarps;
                            nbs[u] = i < nn ? nb_ids[i] : -1;
                            cw[u][0] = cw[u][1] = 0u;
                            if (nbs[u] >= 0) {
                                const uint32_t *row32 = reinterpret_cast<const uint32_t *>(p.codes + (int64_t)nbs[u] * p.code_stride);
                                if (lane < nwords) cw[u][0] = __ldg(row32 + lane);
                                if (lane + 32 < nwords) cw[u][1] = __ldg(row32 + lane + 32);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            if (nbs[u] < 0) continue; // warp-uniform
                            float s = 0.f;
#pragma unroll
                            for (int h = 0; h < 2; h++) {
                                const int m0 = (lane + 32 * h) * 4;
#pragma unroll
                                for (int b = 0; b < 4; b++)
                                    if (m0 + b < p.M)
                                        s = __fadd_rn(s, lut_get(lut, (m0 + b) * p.K + (int)((cw[u][h] >> (8 * b)) & 0xffu)));
                            }
                            if (nwords > 64) { // very wide codes (M > 256): remaining words straight from global
                                const uint32_t *row32 = reinterpret_cast<const uint32_t *>(p.codes + (int64_t)nbs[u] * p.code_stride);
                                for (int w = lane + 64; w < nwords; w += 32) {
                                    const uint32_t c = __ldg(row32 + w);
                                    for (int b = 0; b < 4; b++)
                                        if (w * 4 + b < p.M) s = __fadd_rn(s, lut_get(lut, (w * 4 + b) * p.K + (int)((c >> (8 * b)) & 0xffu)));
                                }
                            }
#pragma unroll
                            for (int off = 16; off >= 1; off >>= 1) s = __fadd_rn(s, __shfl_xor_sync(JV_FULL_MASK, s, off));
                            if (lane == 0) {
                                const float nnorm = p.sim == JV_SIM_COSINE ? __ldg(p.node_norm + nbs[u]) : 0.f;
                                const uint64_t k = fkey_make(adc_finish(p.sim, s, nnorm, qnorm), nbs[u]);
                                if ((k >> 1) > worst) surv[atomicAdd(&s_ns, 1)] = k;
                            }
                        }
                    }
                } else {
                    for (int i = warp; i < nn; i += kFWarps) {
                        const int32_t nb = nb_ids[i];
                        const float s = score_node(nb);
                        if (lane == 0) {
                            const uint64_t k = fkey_make(s, nb);
                            if ((k >> 1) > worst) surv[atomicAdd(&s_ns, 1)] = k;
                        }
                    }
                }
            }
            __syncthreads();
            const int ns = s_ns;

            // ---- (d) merge survivors into the list (dedupe against the list and among themselves)
            bool valid = false;
            uint64_t mine = 0ull;
            if (tid < ns) {
                mine = surv[tid];
                const uint64_t a = mine >> 1;
                const int pos = count_better(list, n, a);
                valid = !(pos < n && (list[pos] >> 1) == a);
                for (int j = 0; j < tid && valid; j++) valid = (surv[j] >> 1) != a;
                sflag[tid] = valid ? 1 : 0;
            }
            __syncthreads();
            if (tid == 0) {
                s_nn = 0;
                s_ns = 0;
            }
            if (valid) {
                const uint64_t a = mine >> 1;
                int pos = count_better(list, n, a);
                for (int j = 0; j < ns; j++) pos += (sflag[j] && (surv[j] >> 1) > a) ? 1 : 0;
                if (pos < L) out[pos] = mine;
            }
            for (int t = tid; t < n; t += kFThreads) {
                const uint64_t k = list[t];
                const uint64_t a = k >> 1;
                int pos = t;
                for (int j = 0; j < ns; j++) pos += (sflag[j] && (surv[j] >> 1) > a) ? 1 : 0;
                if (pos < L) out[pos] = k;
            }
            const int nvalid = __syncthreads_count(valid ? 1 : 0);
            n = n + nvalid < L ? n + nvalid : L;
            cur ^= 1;
        }

        // ---- emit the approximate result list, best first, in the (score, ~node) key format of the rerank kernel
        {
            const uint64_t *list = cur ? list1 : list0;
            uint64_t *o = p.approx_keys + (int64_t)qi * L;
            for (int i = tid; i < L; i += kFThreads) o[i] = i < n ? jv_mk_key(fkey_score(list[i]), fkey_node(list[i])) : 0ull;
            if (tid == 0) {
                p.approx_count[qi] = n;
                if (p.stats) {
                    jv_query_stats st;
                    st.visited = visited;
                    st.expanded = expanded;
                    st.expanded_base = expanded;
                    st.reranked = 0;
                    p.stats[qi] = st;
                }
            }
        }
    }
}

template <bool PQ, typename LutT>
static int32_t launch_fast_typed(jv_index *ix, SearchCtx *ctx, SearchParams &p, size_t fixed) {
    auto kern = fast_search_kernel<PQ, LutT>;
    // shared memory per SM is 228 KB; every resident CTA also reserves 1 KB.  Take the highest occupancy that still
    // leaves a >= 2048-slot visited filter, then give the filter what is left (capped: it only has to cover ~2x the
    // nodes a query touches).
    const size_t sm_total = 228 * 1024;
    int64_t want = (int64_t)2 * p.L * p.R;
    if (want < 2048) want = 2048;
    int best_occ = 0, best_log2 = 0;
    for (int occ = 8; occ >= 1; occ--) {
        const int64_t per = (int64_t)(sm_total / occ) - 1024 - 256 - (int64_t)fixed;
        if (per < 2048 * 4) continue;
        int lg = 11;
        while (lg < 16 && ((int64_t)4 << (lg + 1)) <= per && ((int64_t)1 << lg) < want) lg++;
        best_occ = occ;
        best_log2 = lg;
        break;
    }
    if (!best_occ) {
        set_error("search: shared memory budget exceeded (%zu fixed bytes); use JV_INDEX_FLAG_LUT_F16 or a smaller rerank_k", fixed);
        return JV_ERR_UNSUPPORTED;
    }
    p.hash_log2 = best_log2;
    const size_t smem = fixed + ((size_t)4 << best_log2);
    JV_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    JV_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kFThreads, smem));
    if (occ < 1) {
        set_error("search: kernel does not fit on an SM (smem %zu)", smem);
        return JV_ERR_UNSUPPORTED;
    }
    int grid = ix->sm_count * occ;
    if (grid > p.nq) grid = p.nq;
    JV_CUDA_TRY(cudaMemsetAsync(p.work_counter, 0, sizeof(int), ctx->stream));
    kern<<<grid, kFThreads, smem, ctx->stream>>>(p);
    JV_CUDA_TRY(cudaGetLastError());
    return JV_OK;
}

int32_t launch_search_fast(jv_index *ix, SearchCtx *ctx, SearchParams &p, int expand_width, bool f16) {
    int E = expand_width < 1 ? 1 : (expand_width > kMaxE ? kMaxE : expand_width);
    if (E * p.R > kMaxNew) E = kMaxNew / p.R;
    if (E < 1) E = 1;
    p.expand_width = E;
    p.list_cap = p.L;
    size_t fixed = 0;
    if (ix->has_pq) fixed += (((size_t)p.M * p.K * (f16 ? 2 : 4)) + 15) & ~(size_t)15;
    fixed += (((size_t)p.dim * 4) + 15) & ~(size_t)15;
    fixed += (size_t)p.L * 16 + (size_t)kMaxNew * (8 + 4 + 1);
    if (ix->has_pq) return f16 ? launch_fast_typed<true, __half>(ix, ctx, p, fixed) : launch_fast_typed<true, float>(ix, ctx, p, fixed);
    return launch_fast_typed<false, float>(ix, ctx, p, fixed);
}

}  // namespace jv
