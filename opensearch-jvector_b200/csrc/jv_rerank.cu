// jv_rerank.cu — K3 exact rerank + top-k, K5 brute-force exact top-k, K7 merge of per-shard lists.
//
// K3 replaces the rerank step inside GraphSearcher.search fed by view.rerankerFor(q, sim)
//    (JVectorReader.java:355) plus the ordinal->doc mapping and collector hand-off (JVectorReader.java:175-177).
// K5 replaces JVectorVectorScorer.score() driven by Lucene's exactSearch (JVectorVectorScorer.java:36-53).
// K7 is the on-device analogue of Lucene TopDocs.merge across segments/shards.
// All exact scores use the canonical fp32 reduction of jv_common.cuh, bit-identical to the oracle, so
// top-k ids match exactly with ties broken towards the lower docId.
#include "jv_internal.h"
#include "jv_rerank_body.cuh"

namespace jv {

// ------------------------------------------------------------------------------------------------
// K3: one CTA (4 warps) per query.  Gathers <= rerank_k inline fp32 vectors (dim*4 B each, coalesced
// 16-byte loads), scores them exactly, selects the top k by rank counting.
// ------------------------------------------------------------------------------------------------
constexpr int kRerankThreads = 128;

// THREADS = 128: throughput shape (16 CTAs per SM in flight); 512: small batches — one query's <= rerankK rows are gathered by 16 warps
// instead of 4 (latency: 39 -> ~12 us for a single query at 768-d)
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
rerank_kernel(const float *__restrict__ vectors, const float *__restrict__ vec_norm, const int32_t *__restrict__ ord_to_doc,
              int dim, int sim, int has_pq, const float *__restrict__ queries, int k, int L, float rerank_floor,
              const uint64_t *__restrict__ approx_keys, const int32_t *__restrict__ approx_count, int32_t *out_doc,
              float *out_score, int32_t *out_count, jv_query_stats *stats, NvqView nvq, RowMap rows) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sq = reinterpret_cast<float *>(smem_raw);
    const size_t qb = (((size_t)dim * 4) + 15) & ~(size_t)15;
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem_raw + qb);
    if (nvq.bytes != nullptr) { // per-warp decode buffers behind the keys
        nvq.xbuf = reinterpret_cast<float *>(smem_raw + qb + ((((size_t)L * 8) + 15) & ~(size_t)15)); // 16-B aligned: float4 reads
        nvq.consts = nvq.xbuf + (size_t)(THREADS / 32) * ((dim + 3) & ~3);
    }
    const int qi = blockIdx.x;
    const bool vec4 = (dim & 3) == 0 && ((reinterpret_cast<uintptr_t>(queries) & 15) == 0);
    const int reranked = rerank_query<THREADS>(vectors, vec_norm, ord_to_doc, dim, sim, has_pq, queries + (int64_t)qi * dim, vec4, k,
                                                      approx_count[qi], rerank_floor, approx_keys + (int64_t)qi * L, sq, keys,
                                                      out_doc + (int64_t)qi * k, out_score + (int64_t)qi * k, out_count + qi, nvq, rows);
    if (threadIdx.x == 0 && stats) stats[qi].reranked = reranked;
}


// ------------------------------------------------------------------------------------------------
// Rerank rows of a batch, de-duplicated (rerank vectors in pinned host memory, BASELINE config 5): the candidates of the 10k
// queries of a batch overlap, and every row read by the rerank kernel crosses PCIe.  Mark the rows the batch needs in a bitmap
// over the ordinals, rank the set bits (popcount prefix sum), stream every needed row ONCE from host memory into a dense staging
// array in HBM (one warp per row, 16-byte loads, many rows in flight), and let the rerank kernel read the staging array through
// the row map.  Same rows, same arithmetic: ids and score bits are unchanged.
// ------------------------------------------------------------------------------------------------
__global__ void dd_mark_kernel(const uint64_t *__restrict__ approx_keys, const int32_t *__restrict__ approx_count, int nq, int L, int64_t n,
                               uint32_t *bitmap) {
    const int64_t total = (int64_t)nq * L;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(i / L), j = (int)(i - (int64_t)q * L);
        if (j >= approx_count[q]) continue;
        const uint64_t key = approx_keys[i];
        if (key == 0ull) continue;
        const int32_t node = jv_key_id(key);
        if (node >= 0 && node < n) atomicOr(bitmap + (node >> 5), 1u << (node & 31));
    }
}
constexpr int kDdBlock = 256, kDdWordsPerThread = 4, kDdWordsPerBlock = kDdBlock * kDdWordsPerThread;
// pass 1: set bits per block of 1024 bitmap words
__global__ void __launch_bounds__(kDdBlock) dd_block_sums_kernel(const uint32_t *__restrict__ bitmap, int64_t words, int32_t *block_sums) {
    __shared__ int s_warp[kDdBlock / 32];
    const int64_t w0 = (int64_t)blockIdx.x * kDdWordsPerBlock + (int64_t)threadIdx.x * kDdWordsPerThread;
    int c = 0;
#pragma unroll
    for (int i = 0; i < kDdWordsPerThread; i++)
        if (w0 + i < words) c += __popc(bitmap[w0 + i]);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) c += __shfl_xor_sync(JV_FULL_MASK, c, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < kDdBlock / 32; i++) t += s_warp[i];
        block_sums[blockIdx.x] = t;
    }
}
// pass 2 (one block): exclusive prefix sum of the block sums in place, total behind them
__global__ void __launch_bounds__(1024) dd_scan_blocks_kernel(int32_t *block_sums, int nblocks, int32_t *total) {
    __shared__ int s_part[1024];
    const int per = (nblocks + 1023) / 1024, b0 = threadIdx.x * per;
    int c = 0;
    for (int i = 0; i < per; i++)
        if (b0 + i < nblocks) c += block_sums[b0 + i];
    s_part[threadIdx.x] = c;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) { // Hillis-Steele inclusive scan
        const int v = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = s_part[threadIdx.x] - c; // exclusive prefix of this thread's range
    for (int i = 0; i < per; i++)
        if (b0 + i < nblocks) {
            const int v = block_sums[b0 + i];
            block_sums[b0 + i] = run;
            run += v;
        }
    if (threadIdx.x == 1023) *total = s_part[1023];
}
// pass 3: rank base of every bitmap word + the list of marked ordinals in ascending order
__global__ void __launch_bounds__(kDdBlock) dd_rank_kernel(const uint32_t *__restrict__ bitmap, int64_t words, const int32_t *__restrict__ block_sums,
                                                           int32_t *base, int32_t *uniq) {
    __shared__ int s_warp[kDdBlock / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t w0 = (int64_t)blockIdx.x * kDdWordsPerBlock + (int64_t)threadIdx.x * kDdWordsPerThread;
    uint32_t bits[kDdWordsPerThread];
    int c = 0;
#pragma unroll
    for (int i = 0; i < kDdWordsPerThread; i++) {
        bits[i] = w0 + i < words ? bitmap[w0 + i] : 0u;
        c += __popc(bits[i]);
    }
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(JV_FULL_MASK, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int run = block_sums[blockIdx.x] + incl - c;
    for (int i = 0; i < warp; i++) run += s_warp[i];
#pragma unroll
    for (int i = 0; i < kDdWordsPerThread; i++) {
        if (w0 + i >= words) break;
        base[w0 + i] = run;
        uint32_t b = bits[i];
        while (b) {
            const int bit = __ffs(b) - 1;
            b &= b - 1u;
            uniq[run++] = (int32_t)((w0 + i) * 32 + bit);
        }
    }
}
// gather: one warp per needed row, host (zero-copy) -> staging in HBM
__global__ void __launch_bounds__(256) dd_gather_kernel(const float *__restrict__ vectors, const int32_t *__restrict__ uniq, const int32_t *__restrict__ total,
                                                        int dim, float *staging) {
    const int n_rows = *total, lane = threadIdx.x & 31;
    const int warps = (int)((gridDim.x * blockDim.x) >> 5);
    const bool v4 = (dim & 3) == 0;
    for (int r = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < n_rows; r += warps) {
        const float *src = vectors + (int64_t)uniq[r] * dim;
        float *dst = staging + (int64_t)r * dim;
        if (v4) {
            for (int i = lane; i < dim / 4; i += 32) reinterpret_cast<float4 *>(dst)[i] = __ldg(reinterpret_cast<const float4 *>(src) + i);
        } else {
            for (int i = lane; i < dim; i += 32) dst[i] = __ldg(src + i);
        }
    }
}

// builds the row map of a batch on ctx->stream; returns the staging array to read instead of the host vectors
static int32_t dedupe_rows(jv_index *ix, SearchCtx *ctx, int nq, int L, const uint64_t *d_approx_keys, const int32_t *d_approx_count, RowMap *rows,
                           const float **staging, int *launches) {
    const int64_t words = (ix->n + 31) / 32;
    const int nblocks = (int)((words + kDdWordsPerBlock - 1) / kDdWordsPerBlock);
    int64_t cap = (int64_t)nq * L; // rows the batch can need at most
    if (cap > ix->n) cap = ix->n;
    JV_TRY(ctx->dd_bitmap.ensure((size_t)words * 4));
    JV_TRY(ctx->dd_base.ensure((size_t)words * 4));
    JV_TRY(ctx->dd_blocks.ensure((size_t)(nblocks + 1) * 4));
    JV_TRY(ctx->dd_uniq.ensure((size_t)cap * 4));
    JV_TRY(ctx->dd_rows.ensure((size_t)cap * ix->dim * 4));
    uint32_t *bitmap = ctx->dd_bitmap.as<uint32_t>();
    int32_t *blocks = ctx->dd_blocks.as<int32_t>(), *total = blocks + nblocks;
    JV_CUDA_TRY(cudaMemsetAsync(bitmap, 0, (size_t)words * 4, ctx->stream));
    int64_t mb = ((int64_t)nq * L + 255) / 256;
    if (mb > 148 * 32) mb = 148 * 32;
    dd_mark_kernel<<<(int)mb, 256, 0, ctx->stream>>>(d_approx_keys, d_approx_count, nq, L, ix->n, bitmap);
    dd_block_sums_kernel<<<nblocks, kDdBlock, 0, ctx->stream>>>(bitmap, words, blocks);
    dd_scan_blocks_kernel<<<1, 1024, 0, ctx->stream>>>(blocks, nblocks, total);
    dd_rank_kernel<<<nblocks, kDdBlock, 0, ctx->stream>>>(bitmap, words, blocks, ctx->dd_base.as<int32_t>(), ctx->dd_uniq.as<int32_t>());
    dd_gather_kernel<<<ix->sm_count * 8, 256, 0, ctx->stream>>>(ix->vectors_dev, ctx->dd_uniq.as<int32_t>(), total, ix->dim, ctx->dd_rows.as<float>());
    JV_CUDA_TRY(cudaGetLastError());
    if (launches) *launches += 5;
    rows->bitmap = bitmap;
    rows->base = ctx->dd_base.as<int32_t>();
    *staging = ctx->dd_rows.as<float>();
    return JV_OK;
}

int32_t launch_rerank(jv_index *ix, SearchCtx *ctx, const float *d_queries, int nq, int k, int rerank_k, float rerank_floor,
                      const uint64_t *d_approx_keys, const int32_t *d_approx_count, int32_t *d_out_doc, float *d_out_score,
                      int32_t *d_out_count, jv_query_stats *d_stats, int *launches) {
    if (nq <= 0) return JV_OK;
    const int threads = nq <= ix->sm_count ? 512 : kRerankThreads; // a batch that leaves SMs empty: wide CTAs
    size_t smem = ((((size_t)ix->dim * 4) + 15) & ~(size_t)15) + (size_t)rerank_k * 8;
    NvqView nvq{nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr};
    if (ix->has_nvq) {
        nvq.bytes = ix->nvq_bytes.as<uint8_t>();
        nvq.params = ix->nvq_params.as<float>();
        nvq.gmean = ix->nvq_gmean.as<float>();
        nvq.off = ix->nvq_off.as<int32_t>();
        nvq.m = ix->nvq_m;
        smem = ((smem + 15) & ~(size_t)15) + (size_t)(threads / 32) * (((((size_t)ix->dim * 4) + 15) & ~(size_t)15) + (size_t)ix->nvq_m * 16);
    }
    JV_REQUIRE(smem <= ix->smem_optin - 1024, "rerank_k %d too large for shared memory", rerank_k);
    // rerank vectors in host memory: every row crosses PCIe, so large batches gather the rows they need once (JVGPU_RERANK_DEDUPE=1
    // forces it for any batch, =0 turns it off)
    RowMap rows{nullptr, nullptr};
    const float *vectors = ix->vectors_dev;
    const int dd = q8_knobs().rerank_dedupe;
    if (ix->vectors_on_host && ix->has_pq && !ix->has_nvq && dd != 0 && (dd == 1 || (int64_t)nq * rerank_k >= 65536)) {
        if (dedupe_rows(ix, ctx, nq, rerank_k, d_approx_keys, d_approx_count, &rows, &vectors, launches) != JV_OK) {
            cudaGetLastError(); // no room for the staging array: read the rows in place, as before
            rows = RowMap{nullptr, nullptr};
            vectors = ix->vectors_dev;
        }
    }
    if (threads == 512) {
        JV_CUDA_TRY(cudaFuncSetAttribute(rerank_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rerank_kernel<512><<<nq, 512, smem, ctx->stream>>>(
            vectors, ix->vec_norm.as<float>(), ix->ord_to_doc.as<int32_t>(), ix->dim, ix->sim, ix->has_pq ? 1 : 0,
            d_queries, k, rerank_k, rerank_floor, d_approx_keys, d_approx_count, d_out_doc, d_out_score, d_out_count, d_stats, nvq, rows);
    } else {
        JV_CUDA_TRY(cudaFuncSetAttribute(rerank_kernel<kRerankThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rerank_kernel<kRerankThreads><<<nq, kRerankThreads, smem, ctx->stream>>>(
            vectors, ix->vec_norm.as<float>(), ix->ord_to_doc.as<int32_t>(), ix->dim, ix->sim, ix->has_pq ? 1 : 0,
            d_queries, k, rerank_k, rerank_floor, d_approx_keys, d_approx_count, d_out_doc, d_out_score, d_out_count, d_stats, nvq, rows);
    }
    JV_CUDA_TRY(cudaGetLastError());
    if (launches) *launches += 1;
    return JV_OK;
}

// ------------------------------------------------------------------------------------------------
// K7: merge `group` lists of k (doc, score) per query into one sorted list of k.  One CTA per
// (query, output group): keys to shared memory, bitonic sort descending, first k out.
// ------------------------------------------------------------------------------------------------
constexpr int kMergeThreads = 256;

__global__ void __launch_bounds__(kMergeThreads)
merge_kernel(int g, int group, int nq, int k, const int32_t *__restrict__ docs, const float *__restrict__ scores,
             int32_t *out_doc, float *out_score, int32_t *out_count, int n2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem_raw);
    const int qi = blockIdx.x, og = blockIdx.y, tid = threadIdx.x;
    const int g0 = og * group, g1 = min(g, g0 + group);
    const int nin = (g1 - g0) * k;
    for (int i = tid; i < n2; i += kMergeThreads) {
        uint64_t key = 0ull;
        if (i < nin) {
            const int s = g0 + i / k, j = i % k;
            const int64_t idx = ((int64_t)s * nq + qi) * k + j;
            const int32_t d = docs[idx];
            if (d >= 0) key = jv_mk_key(scores[idx], d);
        }
        keys[i] = key;
    }
    __syncthreads();
    for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < (n2 >> 1); i += kMergeThreads) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
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
    int cnt_local = 0;
    for (int j = tid; j < k; j += kMergeThreads) {
        const uint64_t key = j < n2 ? keys[j] : 0ull;
        const int64_t o = ((int64_t)og * nq + qi) * k + j;
        out_doc[o] = key ? jv_key_id(key) : -1;
        out_score[o] = key ? jv_key_score(key) : 0.f;
        cnt_local += key ? 1 : 0;
    }
    __shared__ int s_cnt;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    if (cnt_local) atomicAdd(&s_cnt, cnt_local);
    __syncthreads();
    if (tid == 0 && out_count && gridDim.y == 1) out_count[qi] = s_cnt;
}

int32_t launch_merge_topk(cudaStream_t stream, int g, int nq, int k, const int32_t *d_docs, const float *d_scores,
                          int32_t *d_out_doc, float *d_out_score, int32_t *d_out_count) {
    if (nq <= 0) return JV_OK;
    JV_REQUIRE(g >= 1 && k >= 1 && k <= 4096, "merge: need g >= 1 and 1 <= k <= 4096");
    const int max_keys = 4096;
    int group = max_keys / k;
    if (group < 2) group = 2;
    DevBuf tmp_doc[2], tmp_score[2];
    const int32_t *in_d = d_docs;
    const float *in_s = d_scores;
    int cur_g = g, flip = 0;
    for (;;) {
        const int ng = (cur_g + group - 1) / group;
        const int per = cur_g < group ? cur_g : group;
        int n2 = 1;
        while (n2 < per * k) n2 <<= 1;
        int32_t *od = d_out_doc;
        float *os = d_out_score;
        if (ng > 1) {
            JV_TRY(tmp_doc[flip].alloc((size_t)ng * nq * k * 4));
            JV_TRY(tmp_score[flip].alloc((size_t)ng * nq * k * 4));
            od = tmp_doc[flip].as<int32_t>();
            os = tmp_score[flip].as<float>();
        }
        const size_t smem = (size_t)n2 * 8;
        JV_CUDA_TRY(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        merge_kernel<<<dim3(nq, ng), kMergeThreads, smem, stream>>>(cur_g, group, nq, k, in_d, in_s, od, os, d_out_count, n2);
        JV_CUDA_TRY(cudaGetLastError());
        if (ng == 1) break;
        in_d = od;
        in_s = os;
        cur_g = ng;
        flip ^= 1;
    }
    if (tmp_doc[0].p || tmp_doc[1].p) JV_CUDA_TRY(cudaStreamSynchronize(stream)); // temporaries die with this scope
    return JV_OK;
}

// ------------------------------------------------------------------------------------------------
// K5: brute force.  Grid = (query tiles, doc slices).  A CTA stages TQ queries in shared memory; each
// warp keeps DPW doc rows in registers (canonical lane layout: lane j owns float4 chunks j, j+32, ..)
// and sweeps the query tile, so a doc row is read from HBM once per CTA and every query float4 read
// from shared memory feeds DPW dot products.  Per-query top-k lists live in shared memory; inserts
// are rare after warm-up (~k ln(n/k) per query) and serialised by a per-query lock.
// ------------------------------------------------------------------------------------------------
constexpr int kExactThreads = 256;

template <int NCH, int DPW>
__global__ void __launch_bounds__(kExactThreads)
exact_kernel(const float *__restrict__ vectors, const float *__restrict__ vec_norm, const int32_t *__restrict__ ord_to_doc,
             int64_t n, int dim, int sim, float mul, const float *__restrict__ queries, int nq, int k, int TQ,
             const uint64_t *__restrict__ accept, int64_t accept_stride, int64_t slice_len, int32_t *part_doc, float *part_score) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q0 = blockIdx.x * TQ;
    const int tq = min(TQ, nq - q0);
    const int n4 = dim >> 2;
    float4 *sq4 = reinterpret_cast<float4 *>(smem_raw);
    unsigned char *sp = smem_raw + (size_t)TQ * dim * 4;
    uint64_t *topk = reinterpret_cast<uint64_t *>(sp); // [TQ][k] ascending, slot 0 = worst
    sp += (size_t)TQ * k * 8;
    float *qnorm = reinterpret_cast<float *>(sp);
    sp += (size_t)TQ * 4;
    int *cnt = reinterpret_cast<int *>(sp);
    sp += (size_t)TQ * 4;
    int *lock = reinterpret_cast<int *>(sp);
    sp += (size_t)TQ * 4;
    unsigned long long *thr = reinterpret_cast<unsigned long long *>(smem_raw + ((((size_t)(sp - smem_raw)) + 7) & ~(size_t)7));

    for (int i = tid; i < tq * n4; i += kExactThreads) {
        const int t = i / n4, c = i - t * n4;
        sq4[t * n4 + c] = __ldg(reinterpret_cast<const float4 *>(queries + (int64_t)(q0 + t) * dim) + c);
    }
    for (int t = tid; t < TQ; t += kExactThreads) {
        cnt[t] = 0;
        lock[t] = 0;
        thr[t] = 0ull;
    }
    __syncthreads();
    for (int t = warp; t < tq; t += kExactThreads / 32) {
        const float *qv = reinterpret_cast<const float *>(sq4 + t * n4);
        float qn = jv_warp_reduce_pair<false>(qv, queries + (int64_t)(q0 + t) * dim, dim, lane, true);
        if (lane == 0) qnorm[t] = qn;
    }
    __syncthreads();

    const int64_t d_begin = (int64_t)blockIdx.y * slice_len;
    const int64_t d_end = min(n, d_begin + slice_len);
    const bool l2 = sim == JV_SIM_EUCLIDEAN;
    for (int64_t base = d_begin + (int64_t)warp * DPW; base < d_end; base += (int64_t)(kExactThreads / 32) * DPW) {
        float4 x[DPW][NCH];
        int32_t doc[DPW];
        float xn[DPW];
#pragma unroll
        for (int d = 0; d < DPW; d++) {
            const int64_t o = base + d;
            doc[d] = -1;
            xn[d] = 0.f;
            if (o < d_end) {
                doc[d] = ord_to_doc ? __ldg(ord_to_doc + o) : (int32_t)o;
                if (sim == JV_SIM_COSINE) xn[d] = __ldg(vec_norm + o);
                const float4 *row = reinterpret_cast<const float4 *>(vectors + o * dim);
#pragma unroll
                for (int c = 0; c < NCH; c++) {
                    const int i = lane + 32 * c;
                    x[d][c] = i < n4 ? __ldg(row + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
        bool any = false;
#pragma unroll
        for (int d = 0; d < DPW; d++) any |= doc[d] >= 0;
        if (!any) continue;
        for (int t = 0; t < tq; t++) {
            float p[DPW][4];
#pragma unroll
            for (int d = 0; d < DPW; d++) p[d][0] = p[d][1] = p[d][2] = p[d][3] = 0.f;
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                const int i = lane + 32 * c;
                if (i < n4) {
                    const float4 qv = sq4[t * n4 + i];
#pragma unroll
                    for (int d = 0; d < DPW; d++) {
                        if (l2) {
                            float e0 = qv.x - x[d][c].x, e1 = qv.y - x[d][c].y, e2 = qv.z - x[d][c].z, e3 = qv.w - x[d][c].w;
                            p[d][0] = __fmaf_rn(e0, e0, p[d][0]);
                            p[d][1] = __fmaf_rn(e1, e1, p[d][1]);
                            p[d][2] = __fmaf_rn(e2, e2, p[d][2]);
                            p[d][3] = __fmaf_rn(e3, e3, p[d][3]);
                        } else {
                            p[d][0] = __fmaf_rn(qv.x, x[d][c].x, p[d][0]);
                            p[d][1] = __fmaf_rn(qv.y, x[d][c].y, p[d][1]);
                            p[d][2] = __fmaf_rn(qv.z, x[d][c].z, p[d][2]);
                            p[d][3] = __fmaf_rn(qv.w, x[d][c].w, p[d][3]);
                        }
                    }
                }
            }
            const uint64_t *abits = accept ? accept + (int64_t)(q0 + t) * accept_stride : nullptr;
#pragma unroll
            for (int d = 0; d < DPW; d++) {
                float v = __fadd_rn(__fadd_rn(p[d][0], p[d][1]), __fadd_rn(p[d][2], p[d][3]));
                v = jv_warp_sum_canonical(v);
                if (doc[d] < 0) continue;
                if (abits && !jv_doc_accepted(abits, doc[d])) continue;
                const float s = jv_finish_score(sim, v, qnorm[t], xn[d]) * mul;
                const uint64_t key = jv_mk_key(s, doc[d]);
                if (key > *reinterpret_cast<volatile unsigned long long *>(&thr[t])) {
                    if (lane == 0) {
                        while (atomicCAS(&lock[t], 0, 1) != 0) {
                        }
                        __threadfence_block();
                        volatile uint64_t *lst = topk + (size_t)t * k;
                        int c = *reinterpret_cast<volatile int *>(&cnt[t]);
                        if (c < k) {
                            int pos = c; // insert keeping ascending order
                            while (pos > 0 && lst[pos - 1] > key) {
                                lst[pos] = lst[pos - 1];
                                pos--;
                            }
                            lst[pos] = key;
                            c++;
                            *reinterpret_cast<volatile int *>(&cnt[t]) = c;
                            if (c == k) *reinterpret_cast<volatile unsigned long long *>(&thr[t]) = lst[0];
                        } else if (key > lst[0]) {
                            int pos = 0; // drop the worst, shift down
                            while (pos + 1 < k && lst[pos + 1] < key) {
                                lst[pos] = lst[pos + 1];
                                pos++;
                            }
                            lst[pos] = key;
                            *reinterpret_cast<volatile unsigned long long *>(&thr[t]) = lst[0];
                        }
                        __threadfence_block();
                        atomicExch(&lock[t], 0);
                    }
                    __syncwarp();
                }
            }
        }
    }
    __syncthreads();
    // emit per-slice partial lists, best first
    for (int i = tid; i < tq * k; i += kExactThreads) {
        const int t = i / k, j = i - t * k;
        const int c = cnt[t];
        const int64_t o = ((int64_t)blockIdx.y * nq + (q0 + t)) * k + j;
        if (j < c) {
            const uint64_t key = topk[(size_t)t * k + (c - 1 - j)];
            part_doc[o] = jv_key_id(key);
            part_score[o] = jv_key_score(key);
        } else {
            part_doc[o] = -1;
            part_score[o] = 0.f;
        }
    }
}

// generic fallback (dim % 4 != 0 or very large dim): one warp per (query, doc) through the shared reducer
__global__ void __launch_bounds__(kExactThreads)
exact_kernel_generic(const float *__restrict__ vectors, const float *__restrict__ vec_norm,
                     const int32_t *__restrict__ ord_to_doc, int64_t n, int dim, int sim, float mul,
                     const float *__restrict__ queries, int nq, int k, const uint64_t *__restrict__ accept,
                     int64_t accept_stride, int64_t slice_len, int32_t *part_doc, float *part_score) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qi = blockIdx.x;
    float *sq = reinterpret_cast<float *>(smem_raw);
    uint64_t *topk = reinterpret_cast<uint64_t *>(smem_raw + ((((size_t)dim * 4) + 15) & ~(size_t)15));
    __shared__ int s_cnt, s_lock;
    __shared__ float s_qn;
    for (int i = tid; i < dim; i += kExactThreads) sq[i] = queries[(int64_t)qi * dim + i];
    if (tid == 0) s_cnt = 0, s_lock = 0;
    __syncthreads();
    if (warp == 0) {
        float qn = jv_warp_reduce_pair<false>(sq, queries + (int64_t)qi * dim, dim, lane, false);
        if (lane == 0) s_qn = qn;
    }
    __syncthreads();
    const uint64_t *abits = accept ? accept + (int64_t)qi * accept_stride : nullptr;
    const int64_t d_begin = (int64_t)blockIdx.y * slice_len, d_end = min(n, d_begin + slice_len);
    for (int64_t o = d_begin + warp; o < d_end; o += kExactThreads / 32) {
        const int32_t doc = ord_to_doc ? ord_to_doc[o] : (int32_t)o;
        if (doc < 0 || (abits && !jv_doc_accepted(abits, doc))) continue;
        const float *x = vectors + o * dim;
        float raw = sim == JV_SIM_EUCLIDEAN ? jv_warp_reduce_pair<true>(sq, x, dim, lane, false)
                                            : jv_warp_reduce_pair<false>(sq, x, dim, lane, false);
        const float s = jv_finish_score(sim, raw, s_qn, sim == JV_SIM_COSINE ? vec_norm[o] : 0.f) * mul;
        const uint64_t key = jv_mk_key(s, doc);
        if (lane == 0) {
            while (atomicCAS(&s_lock, 0, 1) != 0) {
            }
            __threadfence_block();
            volatile uint64_t *lst = topk;
            int c = *reinterpret_cast<volatile int *>(&s_cnt);
            if (c < k) {
                int pos = c;
                while (pos > 0 && lst[pos - 1] > key) {
                    lst[pos] = lst[pos - 1];
                    pos--;
                }
                lst[pos] = key;
                *reinterpret_cast<volatile int *>(&s_cnt) = c + 1;
            } else if (key > lst[0]) {
                int pos = 0;
                while (pos + 1 < k && lst[pos + 1] < key) {
                    lst[pos] = lst[pos + 1];
                    pos++;
                }
                lst[pos] = key;
            }
            __threadfence_block();
            atomicExch(&s_lock, 0);
        }
        __syncwarp();
    }
    __syncthreads();
    for (int j = tid; j < k; j += kExactThreads) {
        const int c = s_cnt;
        const int64_t o = ((int64_t)blockIdx.y * nq + qi) * k + j;
        if (j < c) {
            const uint64_t key = topk[c - 1 - j];
            part_doc[o] = jv_key_id(key);
            part_score[o] = jv_key_score(key);
        } else {
            part_doc[o] = -1;
            part_score[o] = 0.f;
        }
    }
}

template <int NCH, int DPW>
static int32_t launch_exact_typed(jv_index *ix, SearchCtx *ctx, const float *d_queries, int nq, int k, const uint64_t *d_accept,
                                  int64_t stride, int TQ, int S, int64_t slice_len, size_t smem, int32_t *pd, float *ps) {
    auto kern = exact_kernel<NCH, DPW>;
    JV_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const float mul = ix->sim == JV_SIM_MIP ? 2.0f : 1.0f; // JVectorVectorScorer.java:43-50
    dim3 grid((nq + TQ - 1) / TQ, S);
    kern<<<grid, kExactThreads, smem, ctx->stream>>>(ix->vectors_dev, ix->vec_norm.as<float>(), ix->ord_to_doc.as<int32_t>(), ix->n,
                                                     ix->dim, ix->sim, mul, d_queries, nq, k, TQ, d_accept, stride, slice_len, pd, ps);
    JV_CUDA_TRY(cudaGetLastError());
    return JV_OK;
}

int32_t launch_exact_topk(jv_index *ix, SearchCtx *ctx, const float *d_queries, int nq, int k, const uint64_t *d_accept,
                          int64_t accept_stride_words, int32_t *d_out_doc, float *d_out_score, int32_t *d_out_count,
                          int *launches) {
    if (nq <= 0) return JV_OK;
    JV_REQUIRE(k >= 1 && k <= 1024, "exact_topk: 1 <= k <= 1024");
    if (exact_tc_eligible(ix, nq, k, d_accept)) { // tensor-core candidate generation + canonical fp32 re-scoring (jv_exact_tc.cu)
        bool done = false;
        JV_TRY(launch_exact_topk_tc(ix, ctx, d_queries, nq, k, d_out_doc, d_out_score, d_out_count, launches, &done));
        if (done) return JV_OK;
    }
    return launch_exact_topk_fp32(ix, ctx, d_queries, nq, k, d_accept, accept_stride_words, d_out_doc, d_out_score, d_out_count, launches);
}

// the plain fp32 kernel: every (query, vector) pair through the canonical reduction
int32_t launch_exact_topk_fp32(jv_index *ix, SearchCtx *ctx, const float *d_queries, int nq, int k, const uint64_t *d_accept,
                               int64_t accept_stride_words, int32_t *d_out_doc, float *d_out_score, int32_t *d_out_count,
                               int *launches) {
    if (nq <= 0) return JV_OK;
    const int dim = ix->dim;
    const int nch = (dim + 127) / 128;
    const bool fast = (dim & 3) == 0 && nch <= 16 && ((reinterpret_cast<uintptr_t>(d_queries) & 15) == 0);
    int TQ = 1;
    size_t smem = 0;
    if (fast) {
        const size_t per_q = (size_t)dim * 4 + (size_t)k * 8 + 12 + 8;
        TQ = (int)((96 * 1024) / per_q);
        if (TQ > 32) TQ = 32;
        if (TQ < 1) TQ = 1;
        if (TQ > nq) TQ = nq;
        smem = (size_t)TQ * per_q + 16;
    } else {
        smem = ((((size_t)dim * 4) + 15) & ~(size_t)15) + (size_t)k * 8;
    }
    JV_REQUIRE(smem <= ix->smem_optin - 1024, "exact_topk: k=%d dim=%d do not fit shared memory", k, dim);
    const int qtiles = (nq + TQ - 1) / TQ;
    // enough CTAs for ~2 waves, but keep slices >= 2048 docs so per-slice lists amortise
    int S = (2 * ix->sm_count * 2 + qtiles - 1) / qtiles;
    const int64_t max_s = (ix->n + 2047) / 2048;
    if (S > max_s) S = (int)max_s;
    if (S < 1) S = 1;
    if (S > 65535) S = 65535;
    const int64_t slice_len = (ix->n + S - 1) / S;
    JV_TRY(ctx->slice_doc.ensure((size_t)S * nq * k * 4));
    JV_TRY(ctx->slice_score.ensure((size_t)S * nq * k * 4));
    int32_t *pd = ctx->slice_doc.as<int32_t>();
    float *ps = ctx->slice_score.as<float>();
    int32_t st = JV_OK;
    if (fast) {
#define JV_EX(N, D) st = launch_exact_typed<N, D>(ix, ctx, d_queries, nq, k, d_accept, accept_stride_words, TQ, S, slice_len, smem, pd, ps)
        if (nch <= 1) JV_EX(1, 4);
        else if (nch <= 2) JV_EX(2, 4);
        else if (nch <= 4) JV_EX(4, 2);
        else if (nch <= 6) JV_EX(6, 2);
        else if (nch <= 8) JV_EX(8, 2);
        else if (nch <= 12) JV_EX(12, 1);
        else JV_EX(16, 1);
#undef JV_EX
    } else {
        const float mul = ix->sim == JV_SIM_MIP ? 2.0f : 1.0f;
        JV_CUDA_TRY(cudaFuncSetAttribute(exact_kernel_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        exact_kernel_generic<<<dim3(nq, S), kExactThreads, smem, ctx->stream>>>(
            ix->vectors_dev, ix->vec_norm.as<float>(), ix->ord_to_doc.as<int32_t>(), ix->n, dim, ix->sim, mul, d_queries, nq, k,
            d_accept, accept_stride_words, slice_len, pd, ps);
        JV_CUDA_TRY(cudaGetLastError());
    }
    JV_TRY(st);
    JV_TRY(launch_merge_topk(ctx->stream, S, nq, k, pd, ps, d_out_doc, d_out_score, d_out_count));
    if (launches) *launches += 2;
    return JV_OK;
}

}  // namespace jv
