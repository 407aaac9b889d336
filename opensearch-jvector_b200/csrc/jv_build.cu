// jv_build.cu — device-side fixtures for the "next" rows of the scope table:
//   * PQ codebook training  = ProductQuantization.compute(...)      JVectorIndexQuantization.java:123-131
//   * Vamana construction   = GraphIndexBuilder.addGraphNode/cleanup JVectorWriter.java:1383-1422
// Both follow the SAME deterministic definitions as the oracle's fixture builders (oracle/jv_oracle.c
// jvo_pq_train / jvo_graph_build) so that codebooks and adjacency can be compared exactly.
#include "jv_internal.h"

namespace jv {

// ================================================================================================
// mean vector (global centroid / medoid seed): sequential double sum per dimension, thread = dimension
// ================================================================================================
__global__ void mean_kernel(const float *__restrict__ x, int64_t n, int dim, float *out) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= dim) return;
    double acc = 0.0;
    for (int64_t i = 0; i < n; i++) acc += (double)__ldg(x + i * dim + d);
    out[d] = (float)(acc / (double)n);
}

// ================================================================================================
// PQ training: one CTA per subspace; k-means++ (D^2 sampling over 256-element block sums) + Lloyd.
// ================================================================================================
constexpr int kTrainThreads = 256;
constexpr int kKmppBlock = 256;

__device__ __forceinline__ uint64_t splitmix64(uint64_t &s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__device__ __forceinline__ float train_l2sq(const float *__restrict__ xrow, const float *__restrict__ g, const float *c, int len) {
    float acc = 0.f;
    for (int j = 0; j < len; j++) {
        float xv = __ldg(xrow + j);
        if (g) xv = xv - __ldg(g + j);
        const float d = xv - c[j];
        acc = __fmaf_rn(d, d, acc);
    }
    return acc;
}

__global__ void __launch_bounds__(kTrainThreads)
pq_train_kernel(const float *__restrict__ x, int64_t n, int dim, int K, int iters, uint64_t seed, const float *__restrict__ gcent,
                const int32_t *__restrict__ size, const int32_t *__restrict__ off, const int32_t *__restrict__ cboff,
                float *codebooks, float *d2_all, uint8_t *assign_all, int64_t nblk) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *bsum = reinterpret_cast<double *>(smem_raw);
    __shared__ long long s_pick;
    const int m = blockIdx.x, tid = threadIdx.x;
    const int len = size[m], o = off[m];
    float *cb = codebooks + cboff[m];
    float *d2 = d2_all + (int64_t)m * n;
    uint8_t *assign = assign_all + (int64_t)m * n;
    const float *g = gcent ? gcent + o : nullptr;
    uint64_t rng = seed * 0x9E3779B97F4A7C15ULL + (uint64_t)m * 0xD1B54A32D192ED03ULL + 1; // used by thread 0 only

    // ---- k-means++ seeding
    if (tid == 0) s_pick = (long long)(splitmix64(rng) % (uint64_t)n);
    __syncthreads();
    for (int c = 0; c < K; c++) {
        const int64_t pick = s_pick;
        for (int j = tid; j < len; j += kTrainThreads) {
            float xv = x[pick * dim + o + j];
            if (g) xv = xv - g[j];
            cb[(int64_t)c * len + j] = xv;
        }
        __syncthreads();
        if (c == K - 1) break;
        const float *cc = cb + (int64_t)c * len;
        for (int64_t i = tid; i < n; i += kTrainThreads) {
            const float dd = train_l2sq(x + i * dim + o, g, cc, len);
            d2[i] = (c == 0 || dd < d2[i]) ? dd : d2[i];
        }
        __syncthreads();
        for (int64_t b = tid; b < nblk; b += kTrainThreads) {
            double acc = 0.0;
            const int64_t e = min((b + 1) * (int64_t)kKmppBlock, n);
            for (int64_t i = b * kKmppBlock; i < e; i++) acc += (double)d2[i];
            bsum[b] = acc;
        }
        __syncthreads();
        if (tid == 0) {
            double total = 0.0;
            for (int64_t b = 0; b < nblk; b++) total += bsum[b];
            const double r = ((double)(splitmix64(rng) >> 11) * (1.0 / 9007199254740992.0)) * total;
            double run = 0.0;
            int64_t pk = n - 1;
            for (int64_t b = 0; b < nblk; b++) {
                if (run + bsum[b] > r || b == nblk - 1) {
                    const int64_t e = min((b + 1) * (int64_t)kKmppBlock, n);
                    pk = e - 1;
                    for (int64_t i = b * kKmppBlock; i < e; i++) {
                        run += (double)d2[i];
                        if (run > r) {
                            pk = i;
                            break;
                        }
                    }
                    break;
                }
                run += bsum[b];
            }
            s_pick = pk;
        }
        __syncthreads();
    }

    // ---- Lloyd iterations: assign (first minimum wins), then ordinal-order double sums per centroid
    for (int it = 0; it < iters; it++) {
        for (int64_t i = tid; i < n; i += kTrainThreads) {
            float best = INFINITY;
            int idx = 0;
            for (int c = 0; c < K; c++) {
                const float dd = train_l2sq(x + i * dim + o, g, cb + (int64_t)c * len, len);
                if (dd < best) {
                    best = dd;
                    idx = c;
                }
            }
            assign[i] = (uint8_t)idx;
        }
        __syncthreads();
        if (tid < K) {
            for (int j0 = 0; j0 < len; j0 += 8) { // sub-vector handled 8 components at a time
                double sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                int64_t cnt = 0;
                const int jn = min(8, len - j0);
                for (int64_t i = 0; i < n; i++) {
                    if (assign[i] == (uint8_t)tid) {
                        cnt++;
                        for (int j = 0; j < jn; j++) {
                            float xv = x[i * dim + o + j0 + j];
                            if (g) xv = xv - g[j0 + j];
                            sum[j] += (double)xv;
                        }
                    }
                }
                if (cnt > 0)
                    for (int j = 0; j < jn; j++) cb[(int64_t)tid * len + j0 + j] = (float)(sum[j] / (double)cnt);
            }
        }
        __syncthreads();
    }
}

static int32_t pq_train_dev_impl(int device, const float *d_vectors, int64_t n, int dim, int m, int k, int center, int iters,
                                 uint64_t seed, float *d_out_codebooks, float *d_out_gcent) {
    JV_REQUIRE(n >= 1 && dim >= 1 && m >= 1 && m <= dim && k >= 1 && k <= 256, "bad PQ shape");
    JV_REQUIRE(k <= n, "K (%d) exceeds the number of training vectors", k);
    JV_REQUIRE(d_vectors && d_out_codebooks && (!center || d_out_gcent), "NULL buffer");
    const int64_t nblk = (n + kKmppBlock - 1) / kKmppBlock;
    JV_REQUIRE(nblk * 8 <= 160 * 1024, "training sample too large (%lld > %d vectors)", (long long)n, 160 * 1024 / 8 * kKmppBlock);
    PqShape s;
    s.init(dim, m, k);
    DevBuf dsize, doff, dcb, d2, assign;
    std::vector<int32_t> cbo(m);
    for (int i = 0; i < m; i++) cbo[i] = (int32_t)s.cb_off[i];
    JV_TRY(dsize.alloc((size_t)m * 4));
    JV_TRY(doff.alloc((size_t)m * 4));
    JV_TRY(dcb.alloc((size_t)m * 4));
    JV_TRY(d2.alloc((size_t)m * n * 4));
    JV_TRY(assign.alloc((size_t)m * n));
    JV_CUDA_TRY(cudaMemcpy(dsize.p, s.size.data(), (size_t)m * 4, cudaMemcpyHostToDevice));
    JV_CUDA_TRY(cudaMemcpy(doff.p, s.off.data(), (size_t)m * 4, cudaMemcpyHostToDevice));
    JV_CUDA_TRY(cudaMemcpy(dcb.p, cbo.data(), (size_t)m * 4, cudaMemcpyHostToDevice));
    if (center) {
        mean_kernel<<<(dim + 127) / 128, 128>>>(d_vectors, n, dim, d_out_gcent);
        JV_CUDA_TRY(cudaGetLastError());
    }
    const size_t smem = (size_t)nblk * 8;
    JV_CUDA_TRY(cudaFuncSetAttribute(pq_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pq_train_kernel<<<m, kTrainThreads, smem>>>(d_vectors, n, dim, k, iters, seed, center ? d_out_gcent : nullptr,
                                                dsize.as<int32_t>(), doff.as<int32_t>(), dcb.as<int32_t>(), d_out_codebooks,
                                                d2.as<float>(), assign.as<uint8_t>(), nblk);
    JV_CUDA_TRY(cudaGetLastError());
    JV_CUDA_TRY(cudaDeviceSynchronize());
    (void)device;
    return JV_OK;
}

// ================================================================================================
// Vamana construction
// ================================================================================================
constexpr int kPruneThreads = 256;
constexpr int kPruneWarps = kPruneThreads / 32;
constexpr int kAppendCap = 2048;    // old list ++ incoming back-links, in edge order, cut here
constexpr int kPruneMaxCands = 1024; // after sorting best first

struct BuildParams {
    const float *vectors;
    const float *vec_norm;
    int64_t n;
    int dim, sim, R, Rb, beam;
    float overflow, alpha;
    int32_t *adj;  // [n * Rb]
    float *ads;    // [n * Rb] score of the neighbour w.r.t. the owner
    int32_t *deg;  // [n]
    const int32_t *order;
};

__device__ __forceinline__ float pair_score(const BuildParams &p, int32_t c, int32_t s, int lane, bool vec4) {
    const float *cv = p.vectors + (int64_t)c * p.dim, *sv = p.vectors + (int64_t)s * p.dim;
    const float raw = p.sim == JV_SIM_EUCLIDEAN ? jv_warp_reduce_pair<true>(cv, sv, p.dim, lane, vec4)
                                                : jv_warp_reduce_pair<false>(cv, sv, p.dim, lane, vec4);
    const bool cosine = p.sim == JV_SIM_COSINE;
    return jv_finish_score(p.sim, raw, cosine ? __ldg(p.vec_norm + c) : 0.f, cosine ? __ldg(p.vec_norm + s) : 0.f);
}

// retainDiverse (oracle retain_diverse): candidates best first in shared memory; block-cooperative.
// Every thread keeps identical control flow; sel/sels/taken are written redundantly with identical values.
__device__ int retain_diverse_block(const BuildParams &p, const int32_t *cn, const float *cs, int nc, uint8_t *taken,
                                    int32_t *sel, float *sels) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool vec4 = (p.dim & 3) == 0;
    for (int i = tid; i < nc; i += kPruneThreads) taken[i] = 0;
    __syncthreads();
    int nsel = 0;
    for (float a = 1.0f; a <= __fadd_rn(p.alpha, 1e-6f) && nsel < p.R; a = __fadd_rn(a, 0.2f)) {
        for (int i = 0; i < nc && nsel < p.R; i++) {
            if (taken[i]) continue;
            const int32_t c = cn[i];
            const float thresh = __fmul_rn(cs[i], a);
            int viol = 0;
            for (int j = warp; j < nsel; j += kPruneWarps)
                if (pair_score(p, c, sel[j], lane, vec4) > thresh) viol = 1;
            viol = __syncthreads_or(viol);
            if (!viol) {
                taken[i] = 1;
                sel[nsel] = c;
                sels[nsel] = cs[i];
                nsel++;
            }
        }
    }
    __syncthreads();
    return nsel;
}

// (1) new nodes of the batch: candidates = approximate result list of the beam search -> out-edges + edge keys
__global__ void __launch_bounds__(kPruneThreads)
prune_new_kernel(const BuildParams p, int64_t done, int bs, const uint64_t *__restrict__ approx_keys,
                 const int32_t *__restrict__ approx_count, int32_t *out_nodes, float *out_scores, uint64_t *edges, int64_t n_edges_pad) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int32_t *cn = reinterpret_cast<int32_t *>(smem_raw);
    float *cs = reinterpret_cast<float *>(cn + p.beam);
    int32_t *sel = reinterpret_cast<int32_t *>(cs + p.beam);
    float *sels = reinterpret_cast<float *>(sel + p.R);
    uint8_t *taken = reinterpret_cast<uint8_t *>(sels + p.R);
    const int b = blockIdx.x, tid = threadIdx.x;
    const int32_t node = p.order[done + b];
    const int nc = approx_count[b];
    for (int i = tid; i < nc; i += kPruneThreads) {
        const uint64_t k = approx_keys[(int64_t)b * p.beam + i];
        cn[i] = jv_key_id(k);
        cs[i] = jv_key_score(k);
    }
    __syncthreads();
    const int nsel = retain_diverse_block(p, cn, cs, nc, taken, sel, sels);
    for (int j = tid; j < p.Rb; j += kPruneThreads) {
        p.adj[(int64_t)node * p.Rb + j] = j < nsel ? sel[j] : -1;
        p.ads[(int64_t)node * p.Rb + j] = j < nsel ? sels[j] : 0.f;
    }
    for (int j = tid; j < p.R; j += kPruneThreads) {
        const int64_t e = (int64_t)b * p.R + j;
        out_nodes[e] = j < nsel ? sel[j] : -1;
        out_scores[e] = j < nsel ? sels[j] : 0.f;
        edges[e] = j < nsel ? (((uint64_t)(uint32_t)sel[j] << 32) | (uint32_t)e) : ~0ull;
    }
    if (tid == 0) p.deg[node] = nsel;
    // pad the tail of the edge array (last CTA)
    if (b == bs - 1)
        for (int64_t e = (int64_t)bs * p.R + tid; e < n_edges_pad; e += kPruneThreads) edges[e] = ~0ull;
}

// (2) bitonic sort of the edge keys (target << 32 | edge id), ascending, in global memory
constexpr int kSortChunk = 2048;
__global__ void __launch_bounds__(kSortChunk / 2) bitonic_local_kernel(uint64_t *keys, int64_t n2, int64_t size_from, int64_t size_to) {
    // sorts/merges inside chunks of kSortChunk keys in shared memory for all stages with stride < kSortChunk
    __shared__ uint64_t s[kSortChunk];
    const int64_t base = (int64_t)blockIdx.x * kSortChunk;
    const int tid = threadIdx.x;
    s[tid] = keys[base + tid];
    s[tid + kSortChunk / 2] = keys[base + tid + kSortChunk / 2];
    __syncthreads();
    for (int64_t size = size_from; size <= size_to; size <<= 1) {
        int64_t stride = size >> 1;
        if (stride >= kSortChunk) stride = kSortChunk >> 1;
        for (; stride > 0; stride >>= 1) {
            const int lo = 2 * tid - (tid & ((int)stride - 1));
            const int hi = lo + (int)stride;
            const bool asc = ((base + lo) & size) == 0;
            const uint64_t a = s[lo], b = s[hi];
            if (asc ? (a > b) : (a < b)) {
                s[lo] = b;
                s[hi] = a;
            }
            __syncthreads();
        }
    }
    keys[base + tid] = s[tid];
    keys[base + tid + kSortChunk / 2] = s[tid + kSortChunk / 2];
}
__global__ void bitonic_global_kernel(uint64_t *keys, int64_t n2, int64_t size, int64_t stride) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (n2 >> 1)) return;
    const int64_t lo = 2 * i - (i & (stride - 1));
    const int64_t hi = lo + stride;
    const bool asc = (lo & size) == 0;
    const uint64_t a = keys[lo], b = keys[hi];
    if (asc ? (a > b) : (a < b)) {
        keys[lo] = b;
        keys[hi] = a;
    }
}
static int32_t sort_u64(cudaStream_t st, uint64_t *keys, int64_t n2) {
    // n2: power of two, multiple of kSortChunk
    const int chunks = (int)(n2 / kSortChunk);
    bitonic_local_kernel<<<chunks, kSortChunk / 2, 0, st>>>(keys, n2, 2, kSortChunk);
    for (int64_t size = 2 * kSortChunk; size <= n2; size <<= 1) {
        for (int64_t stride = size >> 1; stride >= kSortChunk; stride >>= 1)
            bitonic_global_kernel<<<(unsigned)((n2 / 2 + 255) / 256), 256, 0, st>>>(keys, n2, size, stride);
        bitonic_local_kernel<<<chunks, kSortChunk / 2, 0, st>>>(keys, n2, size, size);
    }
    JV_CUDA_TRY(cudaGetLastError());
    return JV_OK;
}

// (3) segment starts: one work item per distinct back-link target
__global__ void segment_kernel(const uint64_t *__restrict__ edges, int64_t ne, int32_t *seg_start, int *seg_count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ne) return;
    const uint64_t k = edges[i];
    if (k == ~0ull) return;
    if (i == 0 || (uint32_t)(edges[i - 1] >> 32) != (uint32_t)(k >> 32)) seg_start[atomicAdd(seg_count, 1)] = (int32_t)i;
}

// sorts `n2` (power of two <= kAppendCap) keys in shared memory, descending
__device__ void block_sort_desc(uint64_t *keys, int n2) {
    const int tid = threadIdx.x;
    for (int size = 2; size <= n2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < (n2 >> 1); i += kPruneThreads) {
                const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const uint64_t a = keys[lo], b = keys[hi];
                if (desc ? (a < b) : (a > b)) {
                    keys[lo] = b;
                    keys[hi] = a;
                }
            }
            __syncthreads();
        }
}

// (4) apply the back-links of one target (mode 0: segments of the sorted edge list) or enforce degree <= R on one
//     node (mode 1: final cleanup).  List = old ++ incoming in edge order, cut at kAppendCap; if longer than the
//     allowed degree it is sorted best first, cut at kPruneMaxCands and re-pruned with retainDiverse to R.
__global__ void __launch_bounds__(kPruneThreads)
backlink_kernel(const BuildParams p, int mode, int64_t done, const uint64_t *__restrict__ edges, int64_t ne,
                const int32_t *__restrict__ seg_start, const int *__restrict__ seg_count, const float *__restrict__ out_scores,
                int *work_counter, int64_t n_items) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem_raw);              // [kAppendCap]
    int32_t *cn = reinterpret_cast<int32_t *>(keys + kAppendCap);         // [kPruneMaxCands]
    float *cs = reinterpret_cast<float *>(cn + kPruneMaxCands);           // [kPruneMaxCands]
    int32_t *sel = reinterpret_cast<int32_t *>(cs + kPruneMaxCands);      // [R]
    float *sels = reinterpret_cast<float *>(sel + p.R);                   // [R]
    uint8_t *taken = reinterpret_cast<uint8_t *>(sels + p.R);             // [kPruneMaxCands]
    __shared__ long long s_item;
    __shared__ int s_end;
    const int tid = threadIdx.x;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(work_counter, 1);
        __syncthreads();
        const int64_t item = s_item;
        const int64_t limit = mode == 0 ? (int64_t)*seg_count : n_items;
        if (item >= limit) break;
        int32_t u;
        int m = 0;
        int64_t s0 = 0;
        if (mode == 0) {
            s0 = seg_start[item];
            u = (int32_t)(uint32_t)(edges[s0] >> 32);
            // segment length: first position whose target differs (bounded scan, cooperative)
            if (tid == 0) s_end = 0x7fffffff;
            __syncthreads();
            for (int64_t i = s0 + 1 + tid;; i += kPruneThreads) {
                const bool stop = i >= ne || edges[i] == ~0ull || (uint32_t)(edges[i] >> 32) != (uint32_t)u;
                if (stop) atomicMin(&s_end, (int)(i - s0));
                if (__syncthreads_or(stop)) break;
            }
            m = s_end;
        } else {
            u = (int32_t)item;
        }
        const int d0 = p.deg[u];
        // every thread must have read the old degree before thread 0 publishes the new one at the end of the
        // (barrier-free) append path below — without this a late warp sees the updated degree, computes a larger
        // total and takes the prune branch on its own (found by compute-sanitizer synccheck)
        __syncthreads();
        if (mode == 1 && d0 <= p.R) continue;
        int total = d0 + m;
        if (total > kAppendCap) total = kAppendCap;
        const float limit_deg = mode == 0 ? __fmul_rn((float)p.R, p.overflow) : (float)p.R;
        if ((float)total > limit_deg) {
            int n2 = 1;
            while (n2 < total) n2 <<= 1;
            for (int i = tid; i < n2; i += kPruneThreads) {
                uint64_t k = 0ull;
                if (i < d0) {
                    k = jv_mk_key(p.ads[(int64_t)u * p.Rb + i], p.adj[(int64_t)u * p.Rb + i]);
                } else if (i < total) {
                    const uint32_t e = (uint32_t)edges[s0 + (i - d0)];
                    k = jv_mk_key(out_scores[e], p.order[done + e / p.R]);
                }
                keys[i] = k;
            }
            __syncthreads();
            block_sort_desc(keys, n2);
            const int nc = total < kPruneMaxCands ? total : kPruneMaxCands;
            for (int i = tid; i < nc; i += kPruneThreads) {
                cn[i] = jv_key_id(keys[i]);
                cs[i] = jv_key_score(keys[i]);
            }
            __syncthreads();
            const int nsel = retain_diverse_block(p, cn, cs, nc, taken, sel, sels);
            for (int j = tid; j < p.Rb; j += kPruneThreads) {
                p.adj[(int64_t)u * p.Rb + j] = j < nsel ? sel[j] : -1;
                p.ads[(int64_t)u * p.Rb + j] = j < nsel ? sels[j] : 0.f;
            }
            if (tid == 0) p.deg[u] = nsel;
        } else if (m > 0) {
            for (int i = d0 + tid; i < total; i += kPruneThreads) {
                const uint32_t e = (uint32_t)edges[s0 + (i - d0)];
                p.adj[(int64_t)u * p.Rb + i] = p.order[done + e / p.R];
                p.ads[(int64_t)u * p.Rb + i] = out_scores[e];
            }
            if (tid == 0) p.deg[u] = total;
        }
    }
}

// leading-segment merge (JVectorWriter.java:1166-1341, insert-only part): the first n0 ordinals keep their graph.  One warp per
// seed node copies its row (stride R -> Rb) and recomputes the cached neighbour scores (exact pair scores: what the reference
// reads back from the neighbours-score-cache file).  *bad is raised when a neighbour id lies outside [0, n0).
__global__ void seed_kernel(const BuildParams p, const int32_t *__restrict__ seed_adj, int64_t n0, int *bad) {
    const int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (u >= n0) return;
    const bool vec4 = (p.dim & 3) == 0;
    int du = 0;
    for (int j = 0; j < p.R; j++) {
        const int32_t nb = __ldg(seed_adj + u * p.R + j);
        if (nb < 0) break;
        if (nb >= n0) {
            if (lane == 0) atomicExch(bad, 1);
            break;
        }
        const float s = pair_score(p, (int32_t)u, nb, lane, vec4);
        if (lane == 0) {
            p.adj[u * p.Rb + j] = nb;
            p.ads[u * p.Rb + j] = s;
        }
        du++;
    }
    if (lane == 0) p.deg[u] = du;
}

__global__ void order_kernel(int32_t *order, int64_t n, int32_t entry) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == entry)
        order[0] = entry;
    else
        order[i < entry ? i + 1 : i] = (int32_t)i;
}
__global__ void fill_i32_kernel(int32_t *a, int64_t n, int32_t v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}
__global__ void export_adj_kernel(const int32_t *__restrict__ adj, const int32_t *__restrict__ deg, int64_t n, int R, int Rb,
                                  int32_t *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * R) return;
    const int64_t u = i / R;
    const int j = (int)(i - u * R);
    out[i] = j < deg[u] ? adj[u * Rb + j] : -1;
}

// n0 > 0: seeded build (jv_graph_extend): ordinals [0, n0) keep the graph d_seed_adj [n0][R] and its entry node
static int32_t graph_build_dev_impl(int device, const float *d_vectors, int64_t n, int dim, int sim, int R, int beam,
                                    float overflow, float alpha, int32_t *d_out_adj, int32_t *out_entry, int64_t n0 = 0,
                                    const int32_t *d_seed_adj = nullptr, int32_t seed_entry = 0) {
    JV_REQUIRE(n >= 1 && n < 0x7fffffffLL && dim >= 1, "bad n/dim");
    JV_REQUIRE(n0 >= 0 && n0 <= n && (n0 == 0 || (d_seed_adj != nullptr && seed_entry >= 0 && seed_entry < n0)),
               "seed graph: n0 must be in [0, n], seed_entry in [0, n0)");
    JV_REQUIRE(R >= 1 && R <= 96 && beam >= 1 && beam <= 1024, "max_degree must be in [1,96], beam_width in [1,1024]");
    JV_REQUIRE(sim >= JV_SIM_EUCLIDEAN && sim <= JV_SIM_MIP, "unknown similarity");
    JV_REQUIRE(overflow >= 1.0f && overflow <= 1.3f && alpha >= 1.0f, "neighbor_overflow must be in [1,1.3], alpha >= 1");
    JV_REQUIRE(d_vectors && d_out_adj && out_entry, "NULL buffer");
    const int bsim = sim == JV_SIM_MIP ? JV_SIM_DOT : sim; // build scores are plain jVector scores
    int Rb = (int)floorf((float)R * overflow);
    if (Rb < R) Rb = R;

    // a private index object over the caller's vectors; adjacency = the build-time lists (stride Rb)
    jv_index ix;
    ix.device = device;
    ix.sim = bsim;
    ix.dim = dim;
    ix.R = Rb;
    ix.n = n;
    ix.has_pq = false;
    ix.vectors_dev = const_cast<float *>(d_vectors);
    cudaDeviceProp prop;
    JV_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    ix.sm_count = prop.multiProcessorCount;
    ix.smem_optin = prop.sharedMemPerBlockOptin;
    JV_TRY(ix.dbg.alloc(128));
    JV_CUDA_TRY(cudaMemset(ix.dbg.p, 0, 128));
    JV_TRY(ix.adjacency.alloc((size_t)n * Rb * 4));
    SearchCtx ctx;
    struct CtxGuard {
        SearchCtx &c;
        ~CtxGuard() { c.destroy(); }
    } guard{ctx};
    JV_TRY(ctx.init(device));
    cudaStream_t st = ctx.stream;
    if (bsim == JV_SIM_COSINE) {
        JV_TRY(ix.vec_norm.alloc((size_t)n * 4));
        JV_TRY(launch_vec_norms(st, d_vectors, n, dim, ix.vec_norm.as<float>()));
    }
    DevBuf ads, deg, order, mean, e_doc, e_score, e_cnt;
    JV_TRY(ads.alloc((size_t)n * Rb * 4));
    JV_TRY(deg.alloc((size_t)n * 4));
    JV_TRY(order.alloc((size_t)n * 4));
    JV_TRY(mean.alloc((size_t)dim * 4));
    JV_TRY(e_doc.alloc(4));
    JV_TRY(e_score.alloc(4));
    JV_TRY(e_cnt.alloc(4));
    fill_i32_kernel<<<(unsigned)(((int64_t)n * Rb + 255) / 256), 256, 0, st>>>(ix.adjacency.as<int32_t>(), (int64_t)n * Rb, -1);
    JV_CUDA_TRY(cudaMemsetAsync(deg.p, 0, (size_t)n * 4, st));
    JV_CUDA_TRY(cudaMemsetAsync(ads.p, 0, (size_t)n * Rb * 4, st));

    int launches = 0;
    int32_t entry = 0;
    if (n0 > 0) {
        entry = seed_entry; // the leading graph keeps its entry node; insertion order = ordinal order
        order_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(order.as<int32_t>(), n, 0);
    } else {
        // entry = medoid: best exact score against the mean vector (ties -> lower ordinal) = brute force with nq = 1, k = 1
        mean_kernel<<<(dim + 127) / 128, 128, 0, st>>>(d_vectors, n, dim, mean.as<float>());
        JV_CUDA_TRY(cudaGetLastError());
        JV_TRY(launch_exact_topk(&ix, &ctx, mean.as<float>(), 1, 1, nullptr, 0, e_doc.as<int32_t>(), e_score.as<float>(),
                                 e_cnt.as<int32_t>(), &launches));
        JV_CUDA_TRY(cudaMemcpyAsync(&entry, e_doc.p, 4, cudaMemcpyDeviceToHost, st));
        JV_CUDA_TRY(cudaStreamSynchronize(st));
        JV_REQUIRE(entry >= 0 && entry < n, "medoid search failed");
        order_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(order.as<int32_t>(), n, entry);
    }
    *out_entry = entry;
    ix.entry = entry;

    BuildParams bp;
    bp.vectors = d_vectors;
    bp.vec_norm = ix.vec_norm.as<float>();
    bp.n = n;
    bp.dim = dim;
    bp.sim = bsim;
    bp.R = R;
    bp.Rb = Rb;
    bp.beam = beam;
    bp.overflow = overflow;
    bp.alpha = alpha;
    bp.adj = ix.adjacency.as<int32_t>();
    bp.ads = ads.as<float>();
    bp.deg = deg.as<int32_t>();
    bp.order = order.as<int32_t>();

    int64_t bcap = n / 50; // prefix doubling, capped at 2 % of n (integer, as in the oracle) and 8192
    if (bcap > 8192) bcap = 8192;
    if (bcap < 1) bcap = 1;
    int64_t epad = 1;
    while (epad < bcap * R) epad <<= 1;
    if (epad < kSortChunk) epad = kSortChunk;
    DevBuf approx_keys, approx_count, out_nodes, out_scores, edges, seg_start, counters;
    JV_TRY(approx_keys.alloc((size_t)bcap * beam * 8));
    JV_TRY(approx_count.alloc((size_t)bcap * 4));
    JV_TRY(out_nodes.alloc((size_t)bcap * R * 4));
    JV_TRY(out_scores.alloc((size_t)bcap * R * 4));
    JV_TRY(edges.alloc((size_t)epad * 8));
    JV_TRY(seg_start.alloc((size_t)bcap * R * 4));
    JV_TRY(counters.alloc(8));

    const size_t smem_new = (size_t)beam * 8 + (size_t)R * 8 + (size_t)beam + 16;
    const size_t smem_back = (size_t)kAppendCap * 8 + (size_t)kPruneMaxCands * 8 + (size_t)R * 8 + kPruneMaxCands + 16;
    JV_CUDA_TRY(cudaFuncSetAttribute(prune_new_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_new));
    JV_CUDA_TRY(cudaFuncSetAttribute(backlink_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_back));
    int occ_back = 1;
    JV_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_back, backlink_kernel, kPruneThreads, smem_back));
    const int grid_back = ix.sm_count * (occ_back > 0 ? occ_back : 1);

    int64_t done = 1;
    if (n0 > 0) {
        JV_CUDA_TRY(cudaMemsetAsync(counters.p, 0, 8, st));
        seed_kernel<<<(unsigned)((n0 * 32 + 255) / 256), 256, 0, st>>>(bp, d_seed_adj, n0, counters.as<int>());
        JV_CUDA_TRY(cudaGetLastError());
        int bad = 0;
        JV_CUDA_TRY(cudaMemcpyAsync(&bad, counters.p, 4, cudaMemcpyDeviceToHost, st));
        JV_CUDA_TRY(cudaStreamSynchronize(st));
        JV_REQUIRE(bad == 0, "seed graph: a neighbour id lies outside [0, n0)");
        done = n0;
    }
    while (done < n) {
        int64_t bs = done < bcap ? done : bcap;
        if (bs > n - done) bs = n - done;
        // 1. beam search of every batch node against the frozen graph (exact scores, L = beamWidth)
        SearchLaunch a;
        a.d_queries = nullptr;
        a.d_query_ids = order.as<int32_t>() + done;
        a.nq = (int)bs;
        a.rerank_k = beam;
        a.threshold = 0.f;
        a.d_accept = nullptr;
        a.accept_stride_words = 0;
        a.d_approx_keys = approx_keys.as<uint64_t>();
        a.d_approx_count = approx_count.as<int32_t>();
        a.d_stats = nullptr;
        a.entry_override = entry;
        a.n_limit = n;
        a.expand_width = -1; // reference order exactly: the oracle's builder must reproduce this graph bit for bit
        JV_TRY(launch_search(&ix, &ctx, a, &launches));
        // 2. out-edges of the new nodes + edge keys
        int64_t ne_pad = 1;
        while (ne_pad < bs * R) ne_pad <<= 1;
        if (ne_pad < kSortChunk) ne_pad = kSortChunk;
        prune_new_kernel<<<(unsigned)bs, kPruneThreads, smem_new, st>>>(bp, done, (int)bs, approx_keys.as<uint64_t>(),
                                                                         approx_count.as<int32_t>(), out_nodes.as<int32_t>(),
                                                                         out_scores.as<float>(), edges.as<uint64_t>(), ne_pad);
        JV_CUDA_TRY(cudaGetLastError());
        // 3. group the back-links by target, in batch order
        JV_TRY(sort_u64(st, edges.as<uint64_t>(), ne_pad));
        JV_CUDA_TRY(cudaMemsetAsync(counters.p, 0, 8, st));
        segment_kernel<<<(unsigned)((bs * R + 255) / 256), 256, 0, st>>>(edges.as<uint64_t>(), bs * R, seg_start.as<int32_t>(),
                                                                          counters.as<int>());
        // 4. append / re-prune
        backlink_kernel<<<grid_back, kPruneThreads, smem_back, st>>>(bp, 0, done, edges.as<uint64_t>(), bs * R,
                                                                     seg_start.as<int32_t>(), counters.as<int>(),
                                                                     out_scores.as<float>(), counters.as<int>() + 1, 0);
        JV_CUDA_TRY(cudaGetLastError());
        done += bs;
    }
    // cleanup: enforce degree <= R everywhere, then export
    JV_CUDA_TRY(cudaMemsetAsync(counters.p, 0, 8, st));
    backlink_kernel<<<grid_back, kPruneThreads, smem_back, st>>>(bp, 1, 0, nullptr, 0, nullptr, counters.as<int>(), nullptr,
                                                                 counters.as<int>() + 1, n);
    export_adj_kernel<<<(unsigned)(((int64_t)n * R + 255) / 256), 256, 0, st>>>(bp.adj, bp.deg, n, R, Rb, d_out_adj);
    JV_CUDA_TRY(cudaGetLastError());
    JV_CUDA_TRY(cudaStreamSynchronize(st));
    return JV_OK;
}


// ================================================================================================
// Delete consolidation (GraphIndexBuilder.removeDeletedNodes behind markNodeDeleted + cleanup, JVectorWriter.java:1318-1327;
// FreshDiskANN section 4.2): same definition as the oracle's jvo_graph_remove_deleted.
// ================================================================================================
// pass 1, one warp per node: deleted rows are emptied, rows without a deleted neighbour are copied, the rest is queued
__global__ void consolidate_scan_kernel(const int32_t *__restrict__ adj, const uint8_t *__restrict__ deleted, int64_t n, int R,
                                        int32_t *__restrict__ out_adj, int32_t *todo, int *n_todo) {
    const int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (u >= n) return;
    const bool dead = deleted[u] != 0;
    bool hit = false;
    for (int r0 = 0; r0 < R; r0 += 32) {
        const int j = r0 + lane;
        const int32_t nb = j < R ? __ldg(adj + u * R + j) : -1;
        if (j < R) out_adj[u * R + j] = dead ? -1 : nb;
        hit |= nb >= 0 && deleted[nb] != 0;
    }
    hit = __any_sync(0xffffffffu, hit);
    if (!dead && hit && lane == 0) todo[atomicAdd(n_todo, 1)] = (int32_t)u;
}

// ordered compaction of one chunk of kPruneThreads flags: returns this thread's output position (base + rank) and advances base
__device__ __forceinline__ int block_rank(bool flag, int &base, int *warp_tot) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bal = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < kPruneWarps; w++) {
        const int c = warp_tot[w];
        before += w < warp ? c : 0;
        total += c;
    }
    const int pos = base + before + __popc(bal & ((1u << lane) - 1u));
    base += total;
    __syncthreads();
    return pos;
}

// pass 2, one block per queued node (entry_mode: the deleted entry node; only the best live candidate is reported)
__global__ void __launch_bounds__(kPruneThreads)
consolidate_kernel(const BuildParams p, const int32_t *__restrict__ adj, const uint8_t *__restrict__ deleted,
                   const int32_t *__restrict__ todo, int entry_mode, int32_t entry, int32_t *__restrict__ out_adj, int32_t *out_entry) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem_raw);              // [kAppendCap]
    int32_t *ids = reinterpret_cast<int32_t *>(keys + kAppendCap);        // [kAppendCap]
    int32_t *cn = ids + kAppendCap;                                       // [kPruneMaxCands]
    float *cs = reinterpret_cast<float *>(cn + kPruneMaxCands);           // [kPruneMaxCands]
    int32_t *sel = reinterpret_cast<int32_t *>(cs + kPruneMaxCands);      // [R]
    float *sels = reinterpret_cast<float *>(sel + p.R);                   // [R]
    int32_t *nbr = reinterpret_cast<int32_t *>(sels + p.R);               // [R]
    uint8_t *taken = reinterpret_cast<uint8_t *>(nbr + p.R);              // [kPruneMaxCands]
    __shared__ int warp_tot[kPruneWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, R = p.R;
    const int32_t u = entry_mode ? entry : todo[blockIdx.x];
    for (int j = tid; j < R; j += kPruneThreads) nbr[j] = __ldg(adj + (int64_t)u * R + j);
    __syncthreads();
    // candidates in slot order: segment 0 = live neighbours of u, segment 1 + j = live neighbours k != u of deleted neighbour j
    int m = 0;
    for (int s0 = 0; s0 < (R + 1) * R; s0 += kPruneThreads) {
        const int s = s0 + tid;
        int32_t cand = -1;
        if (s < (R + 1) * R) {
            const int seg = s / R, idx = s - seg * R;
            if (seg == 0) {
                cand = nbr[idx];
                if (cand >= 0 && deleted[cand]) cand = -1;
            } else {
                const int32_t j = nbr[seg - 1];
                if (j >= 0 && deleted[j]) {
                    cand = __ldg(adj + (int64_t)j * R + idx);
                    if (cand == u || (cand >= 0 && deleted[cand])) cand = -1;
                }
            }
        }
        const int pos = block_rank(cand >= 0, m, warp_tot);
        if (cand >= 0 && pos < kAppendCap) ids[pos] = cand;
    }
    if (m > kAppendCap) m = kAppendCap;
    __syncthreads();
    const bool vec4 = (p.dim & 3) == 0;
    for (int i = warp; i < m; i += kPruneWarps) {
        const float sc = pair_score(p, u, ids[i], lane, vec4);
        if (lane == 0) keys[i] = jv_mk_key(sc, ids[i]);
    }
    int n2 = 1;
    while (n2 < m) n2 <<= 1;
    for (int i = m + tid; i < n2; i += kPruneThreads) keys[i] = 0ull;
    __syncthreads();
    block_sort_desc(keys, n2);
    // duplicates (a node reachable through several deleted neighbours) carry equal keys and are adjacent
    int w = 0;
    for (int i0 = 0; i0 < m; i0 += kPruneThreads) {
        const int i = i0 + tid;
        const bool uniq = i < m && (i == 0 || keys[i] != keys[i - 1]);
        const int pos = block_rank(uniq, w, warp_tot);
        if (uniq && pos < kPruneMaxCands) {
            cn[pos] = jv_key_id(keys[i]);
            cs[pos] = jv_key_score(keys[i]);
        }
    }
    const int nc = w < kPruneMaxCands ? w : kPruneMaxCands;
    __syncthreads();
    if (entry_mode) {
        if (tid == 0) *out_entry = nc > 0 ? cn[0] : -1;
        return;
    }
    const int nsel = retain_diverse_block(p, cn, cs, nc, taken, sel, sels);
    for (int j = tid; j < R; j += kPruneThreads) out_adj[(int64_t)u * R + j] = j < nsel ? sel[j] : -1;
}

__global__ void first_live_kernel(const uint8_t *__restrict__ deleted, int64_t n, int *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !deleted[i]) atomicMin(out, (int)i);
}

static int32_t graph_remove_deleted_dev_impl(int device, const float *d_vectors, int64_t n, int dim, int sim, int R, float alpha,
                                             const int32_t *d_adj, const uint8_t *d_deleted, int32_t entry, int32_t *d_out_adj,
                                             int32_t *out_entry) {
    JV_REQUIRE(n >= 1 && n < 0x7fffffffLL && dim >= 1, "bad n/dim");
    JV_REQUIRE(R >= 1 && R <= 96 && alpha >= 1.0f, "max_degree must be in [1,96], alpha >= 1");
    JV_REQUIRE(sim >= JV_SIM_EUCLIDEAN && sim <= JV_SIM_MIP, "unknown similarity");
    JV_REQUIRE(d_vectors && d_adj && d_deleted && d_out_adj && out_entry, "NULL buffer");
    JV_REQUIRE(entry >= 0 && entry < n, "entry node outside [0, n)");
    const int bsim = sim == JV_SIM_MIP ? JV_SIM_DOT : sim;
    cudaStream_t st = nullptr;
    JV_CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard {
        cudaStream_t s;
        ~StreamGuard() { cudaStreamDestroy(s); }
    } sg{st};
    DevBuf norms, todo, counters;
    BuildParams bp;
    memset(&bp, 0, sizeof(bp));
    if (bsim == JV_SIM_COSINE) {
        JV_TRY(norms.alloc((size_t)n * 4));
        JV_TRY(launch_vec_norms(st, d_vectors, n, dim, norms.as<float>()));
    }
    bp.vectors = d_vectors;
    bp.vec_norm = norms.as<float>();
    bp.n = n;
    bp.dim = dim;
    bp.sim = bsim;
    bp.R = R;
    bp.Rb = R;
    bp.alpha = alpha;
    JV_TRY(todo.alloc((size_t)n * 4));
    JV_TRY(counters.alloc(8));
    JV_CUDA_TRY(cudaMemsetAsync(counters.p, 0, 8, st));
    consolidate_scan_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(d_adj, d_deleted, n, R, d_out_adj, todo.as<int32_t>(),
                                                                              counters.as<int>());
    JV_CUDA_TRY(cudaGetLastError());
    int n_todo = 0;
    uint8_t entry_dead = 0;
    JV_CUDA_TRY(cudaMemcpyAsync(&n_todo, counters.p, 4, cudaMemcpyDeviceToHost, st));
    JV_CUDA_TRY(cudaMemcpyAsync(&entry_dead, d_deleted + entry, 1, cudaMemcpyDeviceToHost, st));
    JV_CUDA_TRY(cudaStreamSynchronize(st));
    const size_t smem = (size_t)kAppendCap * 12 + (size_t)kPruneMaxCands * 9 + (size_t)R * 12 + 16;
    JV_CUDA_TRY(cudaFuncSetAttribute(consolidate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (n_todo > 0) {
        consolidate_kernel<<<(unsigned)n_todo, kPruneThreads, smem, st>>>(bp, d_adj, d_deleted, todo.as<int32_t>(), 0, 0, d_out_adj, nullptr);
        JV_CUDA_TRY(cudaGetLastError());
    }
    int32_t e = entry;
    if (entry_dead) {
        int32_t *d_e = counters.as<int32_t>() + 1;
        consolidate_kernel<<<1, kPruneThreads, smem, st>>>(bp, d_adj, d_deleted, nullptr, 1, entry, nullptr, d_e);
        JV_CUDA_TRY(cudaGetLastError());
        JV_CUDA_TRY(cudaMemcpyAsync(&e, d_e, 4, cudaMemcpyDeviceToHost, st));
        JV_CUDA_TRY(cudaStreamSynchronize(st));
        if (e < 0) { // no live node around the old entry: the lowest live ordinal
            const int big = 0x7fffffff;
            JV_CUDA_TRY(cudaMemcpyAsync(d_e, &big, 4, cudaMemcpyHostToDevice, st));
            first_live_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_deleted, n, d_e);
            JV_CUDA_TRY(cudaMemcpyAsync(&e, d_e, 4, cudaMemcpyDeviceToHost, st));
            JV_CUDA_TRY(cudaStreamSynchronize(st));
            if (e == big) e = -1;
        }
    }
    JV_CUDA_TRY(cudaStreamSynchronize(st));
    *out_entry = e;
    return JV_OK;
}

}  // namespace jv

using namespace jv;

extern "C" {

int32_t jv_pq_train_dev(int32_t device, const float *d_vectors, int64_t n, int32_t dim, int32_t m, int32_t k, int32_t center,
                        int32_t iters, uint64_t seed, float *d_out_codebooks, float *d_out_gcent) {
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || device < 0 || device >= cnt) {
        set_error("no usable CUDA device %d; libjvgpu has no CPU fallback", device);
        return JV_ERR_CUDA;
    }
    DeviceGuard guard(device);
    return pq_train_dev_impl(device, d_vectors, n, dim, m, k, center, iters, seed, d_out_codebooks, d_out_gcent);
}

int32_t jv_pq_train(int32_t device, const float *vectors, int64_t n, int32_t dim, int32_t m, int32_t k, int32_t center, int32_t iters,
                    uint64_t seed, float *out_codebooks, float *out_gcent) {
    JV_REQUIRE(vectors && out_codebooks && n >= 1 && dim >= 1, "bad arguments");
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || device < 0 || device >= cnt) {
        set_error("no usable CUDA device %d; libjvgpu has no CPU fallback", device);
        return JV_ERR_CUDA;
    }
    DeviceGuard guard(device);
    DevBuf dx, dcb, dg;
    JV_TRY(dx.alloc((size_t)n * dim * 4));
    JV_TRY(dcb.alloc((size_t)k * dim * 4));
    JV_TRY(dg.alloc((size_t)dim * 4));
    JV_CUDA_TRY(cudaMemcpy(dx.p, vectors, (size_t)n * dim * 4, cudaMemcpyHostToDevice));
    JV_TRY(pq_train_dev_impl(device, dx.as<float>(), n, dim, m, k, center, iters, seed, dcb.as<float>(), dg.as<float>()));
    JV_CUDA_TRY(cudaMemcpy(out_codebooks, dcb.p, (size_t)k * dim * 4, cudaMemcpyDeviceToHost));
    if (center && out_gcent) JV_CUDA_TRY(cudaMemcpy(out_gcent, dg.p, (size_t)dim * 4, cudaMemcpyDeviceToHost));
    return JV_OK;
}

int32_t jv_graph_build_dev(int32_t device, const float *d_vectors, int64_t n, int32_t dim, int32_t similarity, int32_t max_degree,
                           int32_t beam_width, float neighbor_overflow, float alpha, int32_t *d_out_adjacency,
                           int32_t *out_entry_node) {
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || device < 0 || device >= cnt) {
        set_error("no usable CUDA device %d; libjvgpu has no CPU fallback", device);
        return JV_ERR_CUDA;
    }
    DeviceGuard guard(device);
    return graph_build_dev_impl(device, d_vectors, n, dim, similarity, max_degree, beam_width, neighbor_overflow, alpha,
                                d_out_adjacency, out_entry_node);
}

int32_t jv_graph_build(int32_t device, const float *vectors, int64_t n, int32_t dim, int32_t similarity, int32_t max_degree,
                       int32_t beam_width, float neighbor_overflow, float alpha, int32_t *out_adjacency, int32_t *out_entry_node) {
    JV_REQUIRE(vectors && out_adjacency && out_entry_node && n >= 1 && dim >= 1 && max_degree >= 1, "bad arguments");
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || device < 0 || device >= cnt) {
        set_error("no usable CUDA device %d; libjvgpu has no CPU fallback", device);
        return JV_ERR_CUDA;
    }
    DeviceGuard guard(device);
    DevBuf dx, dadj;
    JV_TRY(dx.alloc((size_t)n * dim * 4));
    JV_TRY(dadj.alloc((size_t)n * max_degree * 4));
    JV_CUDA_TRY(cudaMemcpy(dx.p, vectors, (size_t)n * dim * 4, cudaMemcpyHostToDevice));
    JV_TRY(graph_build_dev_impl(device, dx.as<float>(), n, dim, similarity, max_degree, beam_width, neighbor_overflow, alpha,
                                dadj.as<int32_t>(), out_entry_node));
    JV_CUDA_TRY(cudaMemcpy(out_adjacency, dadj.p, (size_t)n * max_degree * 4, cudaMemcpyDeviceToHost));
    return JV_OK;
}

// Graph build with PQ build scores (BuildScoreProvider.pqBuildScoreProvider, JVectorWriter.java:238-244, 1143-1151): every score the
// builder takes — search for a node's neighbours, diversity pruning, the approximate centroid — is a score between PQ
// RECONSTRUCTIONS (the provider decodes node i and scores it against the codes of the others, never touching the fp32 vectors),
// i.e. the exact builder run on decode(codes).  The reconstructions are made on the device and dropped afterwards.
int32_t jv_graph_build_pq_dev(int32_t device, const uint8_t *d_codes, int64_t n, int32_t dim, int32_t similarity, int32_t pq_m, int32_t pq_k,
                              const float *d_codebooks, const float *d_gcent, int32_t max_degree, int32_t beam_width,
                              float neighbor_overflow, float alpha, int32_t *d_out_adjacency, int32_t *out_entry_node) {
    JV_REQUIRE(d_codes && d_codebooks && d_out_adjacency && out_entry_node && n >= 1 && dim >= 1 && max_degree >= 1, "bad arguments");
    JV_REQUIRE(pq_m >= 1 && pq_m <= dim && pq_k >= 1 && pq_k <= 256, "bad PQ shape");
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || device < 0 || device >= cnt) {
        set_error("no usable CUDA device %d; libjvgpu has no CPU fallback", device);
        return JV_ERR_CUDA;
    }
    DeviceGuard guard(device);
    PqShape s;
    s.init(dim, pq_m, pq_k);
    DevBuf dx;
    JV_TRY(dx.alloc((size_t)n * dim * 4));
    JV_TRY(launch_pq_decode(nullptr, s, d_codes, n, d_codebooks, d_gcent, dx.as<float>()));
    JV_CUDA_TRY(cudaDeviceSynchronize());
    return graph_build_dev_impl(device, dx.as<float>(), n, dim, similarity, max_degree, beam_width, neighbor_overflow, alpha,
                                d_out_adjacency, out_entry_node);
}

int32_t jv_graph_build_pq(int32_t device, const uint8_t *codes, int64_t n, int32_t dim, int32_t similarity, int32_t pq_m, int32_t pq_k,
                          const float *codebooks, const float *gcent, int32_t max_degree, int32_t beam_width, float neighbor_overflow,
                          float alpha, int32_t *out_adjacency, int32_t *out_entry_node) {
    JV_REQUIRE(codes && codebooks && out_adjacency && out_entry_node && n >= 1 && dim >= 1 && max_degree >= 1, "bad arguments");
    JV_REQUIRE(pq_m >= 1 && pq_m <= dim && pq_k >= 1 && pq_k <= 256, "bad PQ shape");
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || device < 0 || device >= cnt) {
        set_error("no usable CUDA device %d; libjvgpu has no CPU fallback", device);
        return JV_ERR_CUDA;
    }
    DeviceGuard guard(device);
    PqShape s;
    s.init(dim, pq_m, pq_k);
    DevBuf dc, dcb, dg, dadj;
    JV_TRY(dc.alloc((size_t)n * pq_m));
    JV_TRY(dcb.alloc((size_t)s.cb_floats * 4));
    JV_TRY(dadj.alloc((size_t)n * max_degree * 4));
    JV_CUDA_TRY(cudaMemcpy(dc.p, codes, (size_t)n * pq_m, cudaMemcpyHostToDevice));
    JV_CUDA_TRY(cudaMemcpy(dcb.p, codebooks, (size_t)s.cb_floats * 4, cudaMemcpyHostToDevice));
    if (gcent) {
        JV_TRY(dg.alloc((size_t)dim * 4));
        JV_CUDA_TRY(cudaMemcpy(dg.p, gcent, (size_t)dim * 4, cudaMemcpyHostToDevice));
    }
    JV_TRY(jv_graph_build_pq_dev(device, dc.as<uint8_t>(), n, dim, similarity, pq_m, pq_k, dcb.as<float>(), dg.as<float>(), max_degree,
                                 beam_width, neighbor_overflow, alpha, dadj.as<int32_t>(), out_entry_node));
    JV_CUDA_TRY(cudaMemcpy(out_adjacency, dadj.p, (size_t)n * max_degree * 4, cudaMemcpyDeviceToHost));
    return JV_OK;
}

int32_t jv_graph_extend_dev(int32_t device, const float *d_vectors, int64_t n, int64_t n0, const int32_t *d_seed_adjacency,
                            int32_t seed_entry, int32_t dim, int32_t similarity, int32_t max_degree, int32_t beam_width,
                            float neighbor_overflow, float alpha, int32_t *d_out_adjacency) {
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || device < 0 || device >= cnt) {
        set_error("no usable CUDA device %d; libjvgpu has no CPU fallback", device);
        return JV_ERR_CUDA;
    }
    JV_REQUIRE(n0 >= 1, "n0 must be >= 1 (use jv_graph_build for an empty seed)");
    DeviceGuard guard(device);
    int32_t entry = 0;
    return graph_build_dev_impl(device, d_vectors, n, dim, similarity, max_degree, beam_width, neighbor_overflow, alpha,
                                d_out_adjacency, &entry, n0, d_seed_adjacency, seed_entry);
}

int32_t jv_graph_extend(int32_t device, const float *vectors, int64_t n, int64_t n0, const int32_t *seed_adjacency, int32_t seed_entry,
                        int32_t dim, int32_t similarity, int32_t max_degree, int32_t beam_width, float neighbor_overflow, float alpha,
                        int32_t *out_adjacency) {
    JV_REQUIRE(vectors && seed_adjacency && out_adjacency && n >= 1 && n0 >= 1 && n0 <= n && dim >= 1 && max_degree >= 1, "bad arguments");
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || device < 0 || device >= cnt) {
        set_error("no usable CUDA device %d; libjvgpu has no CPU fallback", device);
        return JV_ERR_CUDA;
    }
    DeviceGuard guard(device);
    DevBuf dx, dadj, dseed;
    JV_TRY(dx.alloc((size_t)n * dim * 4));
    JV_TRY(dadj.alloc((size_t)n * max_degree * 4));
    JV_TRY(dseed.alloc((size_t)n0 * max_degree * 4));
    JV_CUDA_TRY(cudaMemcpy(dx.p, vectors, (size_t)n * dim * 4, cudaMemcpyHostToDevice));
    JV_CUDA_TRY(cudaMemcpy(dseed.p, seed_adjacency, (size_t)n0 * max_degree * 4, cudaMemcpyHostToDevice));
    int32_t entry = 0;
    JV_TRY(graph_build_dev_impl(device, dx.as<float>(), n, dim, similarity, max_degree, beam_width, neighbor_overflow, alpha,
                                dadj.as<int32_t>(), &entry, n0, dseed.as<int32_t>(), seed_entry));
    JV_CUDA_TRY(cudaMemcpy(out_adjacency, dadj.p, (size_t)n * max_degree * 4, cudaMemcpyDeviceToHost));
    return JV_OK;
}

int32_t jv_graph_remove_deleted_dev(int32_t device, const float *d_vectors, int64_t n, int32_t dim, int32_t similarity,
                                    int32_t max_degree, float alpha, const int32_t *d_adjacency, const uint8_t *d_deleted,
                                    int32_t entry_node, int32_t *d_out_adjacency, int32_t *out_entry_node) {
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || device < 0 || device >= cnt) {
        set_error("no usable CUDA device %d; libjvgpu has no CPU fallback", device);
        return JV_ERR_CUDA;
    }
    DeviceGuard guard(device);
    return graph_remove_deleted_dev_impl(device, d_vectors, n, dim, similarity, max_degree, alpha, d_adjacency, d_deleted, entry_node,
                                         d_out_adjacency, out_entry_node);
}

int32_t jv_graph_remove_deleted(int32_t device, const float *vectors, int64_t n, int32_t dim, int32_t similarity, int32_t max_degree,
                                float alpha, const int32_t *adjacency, const uint8_t *deleted, int32_t entry_node,
                                int32_t *out_adjacency, int32_t *out_entry_node) {
    JV_REQUIRE(vectors && adjacency && deleted && out_adjacency && out_entry_node && n >= 1 && dim >= 1 && max_degree >= 1, "bad arguments");
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || device < 0 || device >= cnt) {
        set_error("no usable CUDA device %d; libjvgpu has no CPU fallback", device);
        return JV_ERR_CUDA;
    }
    DeviceGuard guard(device);
    DevBuf dx, dadj, dout, ddel;
    JV_TRY(dx.alloc((size_t)n * dim * 4));
    JV_TRY(dadj.alloc((size_t)n * max_degree * 4));
    JV_TRY(dout.alloc((size_t)n * max_degree * 4));
    JV_TRY(ddel.alloc((size_t)n));
    JV_CUDA_TRY(cudaMemcpy(dx.p, vectors, (size_t)n * dim * 4, cudaMemcpyHostToDevice));
    JV_CUDA_TRY(cudaMemcpy(dadj.p, adjacency, (size_t)n * max_degree * 4, cudaMemcpyHostToDevice));
    JV_CUDA_TRY(cudaMemcpy(ddel.p, deleted, (size_t)n, cudaMemcpyHostToDevice));
    JV_TRY(graph_remove_deleted_dev_impl(device, dx.as<float>(), n, dim, similarity, max_degree, alpha, dadj.as<int32_t>(),
                                         ddel.as<uint8_t>(), entry_node, dout.as<int32_t>(), out_entry_node));
    JV_CUDA_TRY(cudaMemcpy(out_adjacency, dout.p, (size_t)n * max_degree * 4, cudaMemcpyDeviceToHost));
    return JV_OK;
}

}  // extern "C"
