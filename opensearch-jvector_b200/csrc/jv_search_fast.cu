// jv_search_fast.cu — the production traversal kernel (K1 + K2 / K4) for unfiltered, threshold-free queries.
//
// Same search as jv_search.cu (GraphSearcher.search, JVectorReader.java:165-173) restructured for the GPU:
//   * ONE sorted list of the best L = rerankK visited nodes, each flagged expanded / unexpanded, replaces the
//     candidate heap + result heap.  Without a filter a candidate ranked below L visited nodes can never be
//     popped by the reference loop (everything better is popped first and fills the result heap), so the bounded
//     list explores exactly the same nodes for expand_width = 1 (up to exact score ties).
//   * expand_width E > 1 expands the E best unexpanded entries per step (a superset exploration: recall >= the
//     strict order's, fewer dependent round trips to HBM, E*R neighbour rows in flight at once).
//   * the visited set is a small 2-way set-associative filter of 15-bit tags in shared memory.  It only avoids
//     re-scoring: a re-scored node is either already in the list (dropped as a duplicate at merge time) or scores
//     below the list's worst entry (dropped again), so evictions cost work, never correctness, and it cannot
//     overflow.  (set, tag) is a bijection of the ordinal, so there are no false positives.
//   * the ADC table may be held in fp16 (JV_INDEX_FLAG_LUT_F16): 96 KB instead of 192 KB at M=192,K=256, which lets two
//     CTAs share an SM.  ADC scores only steer the traversal; returned scores come from the exact rerank (K3).
//   * ADC sums use `adc_lanes` lanes per code row (<= 4 code words per lane) so small-M configurations keep all 32
//     lanes busy; the summation order is "order = adc_lanes" of oracle adc_sum.
// Shared memory per CTA: lut[M*K] | q[dim] | list[2][L] | surv[256] | ids[256] | filter[H] | flags[256].
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

// 4-byte read-only global load the compiler may not sink towards its use: the scoring loop issues the code words of
// several rows back to back and only then starts the table lookups (one DRAM round trip instead of one per row).
__device__ __forceinline__ uint32_t ldg_u32_pinned(const uint32_t *ptr) {
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(ptr));
    return v;
}

// visited filter: returns true when `nb` was NOT present (and records it).  2 tags of 15 bits + valid bit per word.
__device__ __forceinline__ bool filter_insert(uint32_t *filter, int set_bits, bool tagged, int32_t nb) {
    if (tagged) {
        const uint32_t x = ((uint32_t)nb * 0x9E3779B1u) & ((1u << (set_bits + 15)) - 1u); // bijection on set_bits+15 bit ordinals
        const uint32_t set = x >> 15, tag = (x & 0x7fffu) | 0x8000u;
        uint32_t old = filter[set];
        for (;;) {
            if ((old & 0xffffu) == tag || (old >> 16) == tag) return false;
            const uint32_t seen = atomicCAS(&filter[set], old, (old << 16) | tag);
            if (seen == old) return true;
            old = seen;
        }
    } else { // huge shards (ordinal does not fit set+tag): direct-mapped full ordinals
        const uint32_t h = ((uint32_t)nb * 2654435761u) >> (32 - set_bits);
        return atomicExch(&filter[h], (uint32_t)nb) != (uint32_t)nb;
    }
}

// WPL > 0: every lane of a code-row group owns exactly WPL code words (M/4 == WPL * adc_lanes): no per-word predicates
template <bool PQ, typename LutT, bool K256, int WPL>
__global__ void __launch_bounds__(kFThreads, 2) fast_search_kernel(const SearchParams p) {
    constexpr bool EX = WPL > 0;
    constexpr int NT = EX ? WPL : 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = p.L, E = p.expand_width, H = 1 << p.hash_log2;
    const int K = K256 ? 256 : p.K;

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

    __shared__ int s_query, s_nn[2], s_ns[2], w_sel[kFWarps][2 * kMaxE], w_pos[kFWarps][2 * kMaxE];
    __shared__ float s_qnorm;

    const bool vec4 = (p.dim & 3) == 0 && (p.query_ids != nullptr || (reinterpret_cast<uintptr_t>(p.queries) & 15) == 0);
    const bool tagged = p.n_limit <= ((int64_t)1 << (p.hash_log2 + 15));
    // ADC lane geometry
    const int lpn_log2 = p.adc_lanes_log2, LPN = 1 << lpn_log2, G = 32 >> lpn_log2;
    const int sub = lane >> lpn_log2, sl = lane & (LPN - 1);
    const int nwords = p.M >> 2;

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            s_query = atomicAdd(p.work_counter, 1);
            s_nn[0] = 0;
            s_ns[0] = 0;
        }
        __syncthreads();
        const int qi = s_query;
        if (qi >= p.nq) break;
        long long ck[8] = {0, 0, 0, 0, 0, 0, 0, 0}; // per-phase cycles (thread 0), see jv_index_debug_counter
        long long t_prev = clock64();
#define JV_PHASE(i)                      \
    {                                    \
        const long long t_now = clock64(); \
        ck[i] += t_now - t_prev;         \
        t_prev = t_now;                  \
    }
        const float *gq = p.query_ids ? p.vectors + (int64_t)__ldg(p.query_ids + qi) * p.dim : p.queries + (int64_t)qi * p.dim;
        for (int i = tid; i < p.dim; i += kFThreads) sq[i] = __ldg(gq + i);
        for (int i = tid; i < H; i += kFThreads) filter[i] = tagged ? 0u : kEmpty;
        __syncthreads();
        JV_PHASE(0)
        if (PQ) build_lut<LutT>(p, sq, lut, tid, kFThreads);
        if (warp == 0) {
            const float qn = jv_warp_reduce_pair<false, false>(sq, sq, p.dim, lane, (p.dim & 3) == 0);
            if (lane == 0) s_qnorm = qn;
        }
        __syncthreads();
        JV_PHASE(1)
        const float qnorm = s_qnorm;

        // ADC sum of one code row by a group of LPN lanes (valid in the group's lane 0 after the reduction)
        auto adc_group = [&](int32_t nb) -> float {
            float s = 0.f;
            uint32_t cw[4] = {0u, 0u, 0u, 0u};
            if (nb >= 0) {
                const uint32_t *row32 = reinterpret_cast<const uint32_t *>(p.codes + (int64_t)nb * p.code_stride);
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const int w = sl + (t << lpn_log2);
                    if (w < nwords) cw[t] = __ldg(row32 + w);
                }
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const int w = sl + (t << lpn_log2);
                    if (w < nwords) {
                        const LutT *lw = lut + (size_t)w * 4 * K;
                        s = __fadd_rn(s, lut_get(lw, (int)(cw[t] & 0xffu)));
                        s = __fadd_rn(s, lut_get(lw, K + (int)((cw[t] >> 8) & 0xffu)));
                        s = __fadd_rn(s, lut_get(lw, 2 * K + (int)((cw[t] >> 16) & 0xffu)));
                        s = __fadd_rn(s, lut_get(lw, 3 * K + (int)(cw[t] >> 24)));
                    }
                }
                for (int w = sl + (4 << lpn_log2); w < nwords; w += LPN) { // M > 512 only
                    const uint32_t c = __ldg(row32 + w);
                    const LutT *lw = lut + (size_t)w * 4 * K;
                    s = __fadd_rn(s, lut_get(lw, (int)(c & 0xffu)));
                    s = __fadd_rn(s, lut_get(lw, K + (int)((c >> 8) & 0xffu)));
                    s = __fadd_rn(s, lut_get(lw, 2 * K + (int)((c >> 16) & 0xffu)));
                    s = __fadd_rn(s, lut_get(lw, 3 * K + (int)(c >> 24)));
                }
            }
            for (int off = LPN >> 1; off >= 1; off >>= 1) s = __fadd_rn(s, __shfl_xor_sync(JV_FULL_MASK, s, off));
            return s;
        };
        auto exact_warp = [&](int32_t node) -> float { // one warp; result valid in all lanes
            const float *x = p.vectors + (int64_t)node * p.dim;
            const float raw = p.sim == JV_SIM_EUCLIDEAN ? jv_warp_reduce_pair<true>(sq, x, p.dim, lane, vec4)
                                                        : jv_warp_reduce_pair<false>(sq, x, p.dim, lane, vec4);
            const float xn = p.sim == JV_SIM_COSINE ? __ldg(p.vec_norm + node) : 0.f;
            return jv_finish_score(p.sim, raw, qnorm, xn) * p.mip_mul;
        };

        int n = 0, cur = 0, visited = 0, expanded = 0, step = 0;
        if (p.entry >= 0 && p.entry < p.n_limit) {
            if (warp == 0) {
                float s;
                if (PQ) {
                    s = adc_group(sub == 0 ? p.entry : -1);
                    s = adc_finish(p.sim, s, p.sim == JV_SIM_COSINE ? __ldg(p.node_norm + p.entry) : 0.f, qnorm);
                } else {
                    s = exact_warp(p.entry);
                }
                if (lane == 0) {
                    list0[0] = fkey_make(s, p.entry);
                    filter_insert(filter, p.hash_log2, tagged, p.entry);
                }
            }
            n = 1;
            visited = 1;
        }
        __syncthreads();

        while (n > 0) {
            uint64_t *list = cur ? list1 : list0, *out = cur ? list0 : list1;
            const int par = step & 1; // the two step counters are double buffered: no reset race
            // ---- (a) every warp scans the list flags itself (two ballots at L = 50; the list is not written until after
            //          the barrier below), so the selection needs neither a serial section nor a barrier of its own
            int found = 0;
            for (int c0 = 0; c0 < n; c0 += 32) {
                const int i = c0 + lane;
                const bool un = i < n && (list[i] & 1ull);
                const uint32_t ballot = __ballot_sync(JV_FULL_MASK, un);
                const int rank = found + __popc(ballot & ((1u << lane) - 1u));
                if (un && rank < 2 * E) { // ranks E..2E-1 are runners-up: their rows are prefetched into L2 below
                    w_sel[warp][rank] = fkey_node(list[i]);
                    w_pos[warp][rank] = i;
                }
                found += __popc(ballot);
                if (found >= 2 * E) break;
            }
            __syncwarp();
            const int nsel = found < E ? found : E;
            const int npref = found < 2 * E ? (found > E ? found - E : 0) : E;
            if (nsel == 0) break;
            JV_PHASE(2)

            // ---- (b) neighbour rows of the selected nodes -> visited filter -> compacted id list
            {
                const int R = p.R;
                int32_t nb = -1;
                if (tid < nsel * R) {
                    const int ci = tid / R, j = tid - ci * R;
                    nb = __ldg(p.adjacency + (int64_t)w_sel[warp][ci] * R + j);
                } else if (tid >= kFThreads - 32) {
                    // speculative: the next step most likely expands the runners-up; pull their rows into L2 now so that
                    // step's dependent row read is an L2 hit instead of a DRAM round trip
                    const int lines = (R * 4 + 127) >> 7, t = tid - (kFThreads - 32);
                    if (t < npref * lines) {
                        const char *row = reinterpret_cast<const char *>(p.adjacency + (int64_t)w_sel[warp][nsel + t / lines] * R) + (t % lines) * 128;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
                    }
                }
                const bool fresh = nb >= 0 && nb < p.n_limit && filter_insert(filter, p.hash_log2, tagged, nb);
                const uint32_t ballot = __ballot_sync(JV_FULL_MASK, fresh);
                int base = 0;
                if (lane == 0 && ballot) base = atomicAdd(&s_nn[par], __popc(ballot));
                base = __shfl_sync(JV_FULL_MASK, base, 0);
                if (fresh) nb_ids[base + __popc(ballot & ((1u << lane) - 1u))] = nb;
            }
            __syncthreads();
            if (warp == 0 && lane < nsel) list[w_pos[0][lane]] &= ~1ull; // mark expanded: flags are next read in (d), two barriers away
            JV_PHASE(3)
            const int nn = s_nn[par];
            if (tid == 0) { // next step's counters
                s_nn[par ^ 1] = 0;
                s_ns[par ^ 1] = 0;
            }
            expanded += nsel;
            visited += nn;

            // ---- (c) score the gathered neighbours; keep those that can enter the list and are not already in it
            {
                const uint64_t worst = n >= L ? (list[L - 1] >> 1) : 0ull;
                auto offer = [&](float score, int32_t nb) {
                    const uint64_t k = fkey_make(score, nb);
                    const uint64_t a = k >> 1;
                    if (a > worst) {
                        const int pos = count_better(list, n, a);
                        if (!(pos < n && (list[pos] >> 1) == a)) surv[atomicAdd(&s_ns[par], 1)] = k; // drop re-scored list members
                    }
                };
                if (PQ) {
                    // 4 groups of code rows per warp in flight: every load is issued before the first table lookup
                    constexpr int U = 4;
                    for (int i0 = warp * G; i0 < nn; i0 += kFWarps * G * U) {
                        uint32_t cw[U][NT];
                        int32_t nbv[U];
#pragma unroll
                        for (int u = 0; u < U; u++) {
                            const int i = i0 + u * kFWarps * G + sub;
                            nbv[u] = i < nn ? nb_ids[i] : -1;
#pragma unroll
                            for (int t = 0; t < NT; t++) cw[u][t] = 0u;
                            if (nbv[u] >= 0) {
                                const uint32_t *row32 = reinterpret_cast<const uint32_t *>(p.codes + (int64_t)nbv[u] * p.code_stride);
#pragma unroll
                                for (int t = 0; t < NT; t++) {
                                    const int w = sl + (t << lpn_log2);
                                    if (EX || w < nwords) cw[u][t] = ldg_u32_pinned(row32 + w);
                                }
                            }
                        }
#pragma unroll
                        for (int u = 0; u < U; u++) {
                            if (i0 + u * kFWarps * G >= nn) break; // warp-uniform
                            float s = 0.f;
                            if (nbv[u] >= 0) {
#pragma unroll
                                for (int t = 0; t < NT; t++) {
                                    const int w = sl + (t << lpn_log2);
                                    if (EX || w < nwords) {
                                        const LutT *lw = lut + (size_t)w * 4 * K;
                                        s = __fadd_rn(s, lut_get(lw, (int)(cw[u][t] & 0xffu)));
                                        s = __fadd_rn(s, lut_get(lw, K + (int)((cw[u][t] >> 8) & 0xffu)));
                                        s = __fadd_rn(s, lut_get(lw, 2 * K + (int)((cw[u][t] >> 16) & 0xffu)));
                                        s = __fadd_rn(s, lut_get(lw, 3 * K + (int)(cw[u][t] >> 24)));
                                    }
                                }
                                if (!EX && nwords > (4 << lpn_log2)) { // M > 512 only
                                    const uint32_t *row32 = reinterpret_cast<const uint32_t *>(p.codes + (int64_t)nbv[u] * p.code_stride);
                                    for (int w = sl + (4 << lpn_log2); w < nwords; w += LPN) {
                                        const uint32_t c = __ldg(row32 + w);
                                        const LutT *lw = lut + (size_t)w * 4 * K;
                                        s = __fadd_rn(s, lut_get(lw, (int)(c & 0xffu)));
                                        s = __fadd_rn(s, lut_get(lw, K + (int)((c >> 8) & 0xffu)));
                                        s = __fadd_rn(s, lut_get(lw, 2 * K + (int)((c >> 16) & 0xffu)));
                                        s = __fadd_rn(s, lut_get(lw, 3 * K + (int)(c >> 24)));
                                    }
                                }
                            }
                            for (int off = LPN >> 1; off >= 1; off >>= 1) s = __fadd_rn(s, __shfl_xor_sync(JV_FULL_MASK, s, off));
                            if (sl == 0 && nbv[u] >= 0) {
                                const float nnorm = p.sim == JV_SIM_COSINE ? __ldg(p.node_norm + nbv[u]) : 0.f;
                                offer(adc_finish(p.sim, s, nnorm, qnorm), nbv[u]);
                            }
                        }
                    }
                } else {
                    for (int i = warp; i < nn; i += kFWarps) {
                        const int32_t nb = nb_ids[i];
                        const float s = exact_warp(nb);
                        if (lane == 0) offer(s, nb);
                    }
                }
            }
            __syncthreads();
            JV_PHASE(4)
            const int ns = s_ns[par];

            // ---- (d) single-pass merge: survivors are distinct (atomic filter insertion) and not in the list
            //          (checked in (c)); position = own rank + number of better keys on the other side
            auto count_surv_better = [&](uint64_t a) -> int {
                int c0 = 0, c1 = 0, c2 = 0, c3 = 0, j = 0;
                for (; j + 4 <= ns; j += 4) { // 4 independent shared-memory reads in flight
                    c0 += ((surv[j] >> 1) > a) ? 1 : 0;
                    c1 += ((surv[j + 1] >> 1) > a) ? 1 : 0;
                    c2 += ((surv[j + 2] >> 1) > a) ? 1 : 0;
                    c3 += ((surv[j + 3] >> 1) > a) ? 1 : 0;
                }
                for (; j < ns; j++) c0 += ((surv[j] >> 1) > a) ? 1 : 0;
                return (c0 + c1) + (c2 + c3);
            };
            if (tid < ns) {
                const uint64_t mine = surv[tid];
                const uint64_t a = mine >> 1;
                int pos = count_better(list, n, a) + count_surv_better(a);
                // a node can be queued twice in one step (the visited filter evicted an entry that was inserted in this
                // very step): equal keys are ordered by queue index so that every element keeps its own slot (no holes);
                // the copy is removed when the list is emitted
                for (int j = 0; j < tid; j++) pos += ((surv[j] >> 1) == a) ? 1 : 0;
                if (pos < L) out[pos] = mine;
            }
            for (int t = kFThreads - 1 - tid; t < n; t += kFThreads) { // list entries on the high warps: survivors use the low ones
                const uint64_t k = list[t];
                const int pos = t + count_surv_better(k >> 1);
                if (pos < L) out[pos] = k;
            }
            __syncthreads();
            JV_PHASE(5)
            n = n + ns < L ? n + ns : L;
            cur ^= 1;
            step++;
        }

        // ---- emit the approximate result list, best first, in the (score, ~node) key format of the rerank kernel
        {
            const uint64_t *list = cur ? list1 : list0;
            uint64_t *o = p.approx_keys + (int64_t)qi * L;
            int dup = 0; // a node scored twice in one step sits in two adjacent slots (see the merge)
            for (int i = tid + 1; i < n; i += kFThreads) dup |= ((list[i] >> 1) == (list[i - 1] >> 1)) ? 1 : 0;
            if (!__syncthreads_or(dup)) {
                for (int i = tid; i < L; i += kFThreads) o[i] = i < n ? jv_mk_key(fkey_score(list[i]), fkey_node(list[i])) : 0ull;
                if (tid == 0) p.approx_count[qi] = n;
            } else if (tid == 0) { // rare: serial compaction
                int w = 0;
                for (int i = 0; i < n; i++)
                    if (i == 0 || (list[i] >> 1) != (list[i - 1] >> 1)) o[w++] = jv_mk_key(fkey_score(list[i]), fkey_node(list[i]));
                p.approx_count[qi] = w;
                for (; w < L; w++) o[w] = 0ull;
            }
            if (tid == 0) {
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
        JV_PHASE(6)
        if (tid == 0 && p.dbg) {
            unsigned long long *ph = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(p.dbg) + 64);
            for (int i = 0; i < 7; i++) atomicAdd(ph + i, (unsigned long long)ck[i]);
            atomicAdd(ph + 7, (unsigned long long)step);
        }
#undef JV_PHASE
    }
}

template <bool PQ, typename LutT, bool K256, int WPL>
static int32_t launch_fast_typed(jv_index *ix, SearchCtx *ctx, SearchParams &p, size_t fixed) {
    auto kern = fast_search_kernel<PQ, LutT, K256, WPL>;
    // shared memory per SM is 228 KB; every resident CTA also reserves 1 KB.  Take the highest occupancy that still
    // leaves a >= 2048-word visited filter (4096 tags), then give the filter what is left, capped at ~2x the nodes a
    // query touches.
    const size_t sm_total = 228 * 1024;
    int64_t want = (int64_t)p.L * p.R; // words; 2 tags each
    if (want < 2048) want = 2048;
    int best_occ = 0, best_log2 = 0;
    for (int occ = 8; occ >= 1; occ--) {
        const int64_t per = (int64_t)(sm_total / occ) - 1024 - 1152 - (int64_t)fixed; // 1 KB system + static __shared__
        if (per < 2048 * 4) continue;
        int lg = 11;
        while (lg < 15 && ((int64_t)4 << (lg + 1)) <= per && ((int64_t)1 << lg) < want) lg++;
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

// lanes per code row: the smallest power of two that leaves <= 4 code words (16 subspaces) per lane
int adc_lanes_for(int M) {
    const int nwords = (M + 3) / 4;
    int lanes = 1;
    while (lanes < 32 && lanes * 4 < nwords) lanes <<= 1;
    return lanes;
}

int32_t launch_search_fast(jv_index *ix, SearchCtx *ctx, SearchParams &p, int expand_width, bool f16) {
    int E = expand_width < 1 ? 1 : (expand_width > kMaxE ? kMaxE : expand_width);
    if (E * p.R > kMaxNew) E = kMaxNew / p.R;
    if (E < 1) E = 1;
    p.expand_width = E;
    p.list_cap = p.L;
    p.codebooks_h = (f16 && ix->has_pq) ? ix->codebooks_h.as<__half>() : nullptr;
    int lanes = ix->has_pq ? adc_lanes_for(p.M) : 32, lg = 0;
    while ((1 << lg) < lanes) lg++;
    p.adc_lanes_log2 = lg;
    size_t fixed = 0;
    if (ix->has_pq) fixed += (((size_t)p.M * p.K * (f16 ? 2 : 4)) + 15) & ~(size_t)15;
    fixed += (((size_t)p.dim * 4) + 15) & ~(size_t)15;
    fixed += (size_t)p.L * 16 + (size_t)kMaxNew * (8 + 4 + 1);
    if (ix->has_pq) {
        if ((p.M & 3) != 0) {
            set_error("fast search kernel needs M %% 4 == 0 (M=%d)", p.M);
            return JV_ERR_UNSUPPORTED;
        }
        const bool k256 = p.K == 256;
        const int nwords = p.M >> 2;
        const int wpl = (nwords % lanes == 0 && (nwords / lanes == 3 || nwords / lanes == 4)) ? nwords / lanes : 0;
        if (k256 && wpl == 3)
            return f16 ? launch_fast_typed<true, __half, true, 3>(ix, ctx, p, fixed) : launch_fast_typed<true, float, true, 3>(ix, ctx, p, fixed);
        if (k256 && wpl == 4)
            return f16 ? launch_fast_typed<true, __half, true, 4>(ix, ctx, p, fixed) : launch_fast_typed<true, float, true, 4>(ix, ctx, p, fixed);
        if (f16) return k256 ? launch_fast_typed<true, __half, true, 0>(ix, ctx, p, fixed) : launch_fast_typed<true, __half, false, 0>(ix, ctx, p, fixed);
        return k256 ? launch_fast_typed<true, float, true, 0>(ix, ctx, p, fixed) : launch_fast_typed<true, float, false, 0>(ix, ctx, p, fixed);
    }
    return launch_fast_typed<false, float, false, 0>(ix, ctx, p, fixed);
}

}  // namespace jv
