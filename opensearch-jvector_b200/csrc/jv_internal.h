// jv_internal.h — the opaque index handle and the kernel launchers shared by the .cu files.
#pragma once
#include "jv_common.cuh"

namespace jv {

// per-call scratch: one stream + buffers, recycled through a pool so concurrent Java threads
// (KNNJVectorTests.java:982-1028) never share mutable state.
struct SearchCtx {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaStream_t copy_stream = nullptr;  // H2D of query chunks overlapping the kernels of the previous chunk
    cudaEvent_t chunk_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    DevBuf queries, out_doc, out_score, out_count, stats, accept;
    DevBuf approx_keys, approx_count; // [nq*rerank_k] uint64 keys (score, ordinal), best first
    DevBuf counter;                   // persistent-grid work counter
    DevBuf visited;                   // global visited tables (fallback when they do not fit in smem)
    DevBuf lut8, qparams;             // 8-bit ADC tables of the current chunk + per-query (delta, base, ||q||^2)
    DevBuf slice_doc, slice_score;    // brute force: per-slice partial top-k
    DevBuf dd_bitmap, dd_base, dd_blocks, dd_uniq, dd_rows; // de-duplicated rerank rows of a batch (vectors in host memory): bitmap over
                                      // the ordinals, rank base per bitmap word, block sums, marked ordinals, gathered rows
    DevBuf tc_q, tc_f, tc_chunk, tc_cand, tc_cnt, tc_redo; // brute force on the tensor cores (jv_exact_tc.cu): bf16 queries, per-query floats,
                                      // pass-A chunk maxima, pass-B candidates + counts
    bool lut_timed = false;           // ev[5] was recorded after the first chunk's table build
    bool time_lut = true;             // this launch is the timed chunk of a pipelined host batch
    int last_width = 0, last_kernel = 0; // what the last traversal launch used (jv_batch_timing)
    void *pinned = nullptr;           // host staging
    size_t pinned_bytes = 0;
    int32_t init(int device);
    int32_t ensure_pinned(size_t bytes);
    void destroy();
};

}  // namespace jv

struct jv_index {
    int device = 0;
    int sim = 0, dim = 0, R = 0, entry = 0, max_doc = 0;
    int64_t n = 0;
    uint32_t flags = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    bool has_pq = false;
    bool vectors_on_host = false;
    jv::PqShape pq;
    int code_stride = 0; // bytes per code row (M rounded up to 16)
    jv::DevBuf adjacency, vectors, vec_norm, ord_to_doc, codes, codebooks, gcent, pq_size, pq_off, pq_cboff, node_norm;
    jv::DevBuf dbg;   // int32[4] diagnostic counters + uint64[8] phase cycles
    jv::DevBuf codebooks_h; // fp16 copy of the codebooks (table build of the fast kernel)
    jv::DevBuf fused; // neighbour-interleaved records (optional)
    // 8-bit table path (JV_INDEX_FLAG_LUT_U8): lane-major permuted codes, bounding balls of the subspace codebooks
    jv::DevBuf codes_q8, ball_ctr, ball_rad;
    // NVQ-inline vectors (nvq+pq segments): the reranker scores the dequantised vector
    jv::DevBuf nvq_bytes, nvq_params, nvq_gmean, nvq_off;
    bool has_nvq = false;
    bool nvq_only = false;   // no auxiliary PQ: traversal scored by the NVQ reranker = exact scores over the dequantised vectors
    bool fp32_given = true;  // fp32 inline vectors came with the index (brute force needs them on NVQ segments)
    int nvq_m = 0;
    // tensor-core brute force (jv_exact_tc.cu): bf16 copy of the vectors (COSINE: normalised), EUCLIDEAN bias, max ||x||^2; lazy
    jv::DevBuf tc_base, tc_bias, tc_maxnorm;
    bool tc_ready = false;
    int tc_dp = 0;
    int64_t tc_batches = 0, tc_fallbacks = 0; // diagnostics (jv_index_debug_counter 4, 5)
    bool q8_ok = false;
    int q8_nj = 0; // 32-subspace blocks per code row (M rounded up to 32)
    int fused_stride = 0;
    float *vectors_dev = nullptr; // device-visible pointer to the fp32 vectors (HBM or mapped pinned host)
    void *vectors_host = nullptr;
    int64_t device_bytes = 0;
    std::mutex mu;
    std::vector<jv::SearchCtx *> pool;
    jv::SearchCtx *acquire();
    void release(jv::SearchCtx *c);
};

namespace jv {

// diagnostic environment knobs, read once per process (never on the search path)
struct Q8Knobs {
    int occ = 0;        // JVGPU_Q8_OCC: cap the CTAs per SM
    int chunk = 0;      // JVGPU_Q8_CHUNK: table staging buffer of n queries
    int warps = 0;      // JVGPU_Q8_WARPS: warps per CTA (4 or 8; manager / scorer kernel: 4, 5 or 8 = 3, 4 or 7 scorer warps)
    bool prof = false;  // JVGPU_PROFILE: per-phase cycle counters (synchronous kernel)
    bool fused = false; // JVGPU_Q8_FUSED: K3 as the epilogue of the synchronous kernel
    bool sync = false;  // JVGPU_Q8_SYNC: the round-synchronous kernel of jv_q8.cu instead of the manager / scorer kernel (jv_q8_beam.cu)
    int depth = 0;      // JVGPU_Q8_DEPTH: steps in flight of the manager / scorer kernel (1 or 2; default 2)
    int rerank_dedupe = -1; // JVGPU_RERANK_DEDUPE: gather the rerank rows of a batch once when the vectors live in host memory (1 always, 0 never, unset = large batches)
    bool h2d_single = false; // JVGPU_H2D_SINGLE: no chunked H2D pipeline in jv_search_batch
    int exact_tc = -1;  // JVGPU_EXACT_TC: brute force on the tensor cores (jv_exact_tc.cu): 1 always (tests), 0 never, unset = by size
};
const Q8Knobs &q8_knobs();   // the environment is read once per process ...
void q8_knobs_refresh();     // ... and again on jv_index_debug_counter(which = 200) (tests and sweeps change the knobs)

// ---- kernel launchers (each returns a jv_status; all work is enqueued on `stream`) ----

// K1+K2 / K4: graph traversal with ADC (PQ) or exact scoring; writes the approximate result list.
struct SearchLaunch {
    const float *d_queries;
    const int32_t *d_query_ids = nullptr; // builder: query i = stored vector query_ids[i]
    int nq;
    int rerank_k;
    float threshold;
    const uint64_t *d_accept;
    int64_t accept_stride_words;
    uint64_t *d_approx_keys; // [nq * rerank_k]
    int32_t *d_approx_count; // [nq]
    jv_query_stats *d_stats; // [nq]
    int expand_width = 0;    // 0 = default (4), >= 1 explicit width (fast kernel); -1 = strict reference-order kernel
    int entry_override;      // -1 = index entry
    int64_t n_limit;         // nodes >= n_limit are ignored (graph builder); n for queries
    // optional: let the traversal kernel run K3 as its epilogue (production 8-bit path); fuse_k = 0 -> never fused
    int fuse_k = 0;
    float rerank_floor = 0.f;
    int32_t *d_out_doc = nullptr;
    float *d_out_score = nullptr;
    int32_t *d_out_count = nullptr;
};
// *reranked (nullable) is set when the launched kernel already produced the final top-k (fused K3)
int32_t launch_search(jv_index *ix, SearchCtx *ctx, const SearchLaunch &a, int *launches, bool *reranked = nullptr);

// K1+K2 with the 8-bit table (jv_q8.cu)
bool q8_search_supported(const jv_index *ix, int L, int R, bool filtered);
int q8_lut_bytes(int nj); // table bytes per query in the bank-interleaved layout
int32_t launch_search_q8(jv_index *ix, SearchCtx *ctx, const SearchLaunch &a, int *launches, bool *reranked);
int32_t launch_lut_q8(jv_index *ix, cudaStream_t stream, const float *d_queries, int nq, uint8_t *d_lut, float4 *d_qparams);
int32_t launch_permute_codes(cudaStream_t stream, const uint8_t *d_codes, int64_t n, int M, int stride, int NJ, uint8_t *d_out);

// K3: exact rerank of the approximate list + top-k + ordinal->doc mapping
int32_t launch_rerank(jv_index *ix, SearchCtx *ctx, const float *d_queries, int nq, int k, int rerank_k, float rerank_floor,
                      const uint64_t *d_approx_keys, const int32_t *d_approx_count, int32_t *d_out_doc, float *d_out_score,
                      int32_t *d_out_count, jv_query_stats *d_stats, int *launches);

// K5
int32_t launch_exact_topk(jv_index *ix, SearchCtx *ctx, const float *d_queries, int nq, int k, const uint64_t *d_accept,
                          int64_t accept_stride_words, int32_t *d_out_doc, float *d_out_score, int32_t *d_out_count,
                          int *launches);

int32_t launch_exact_topk_fp32(jv_index *ix, SearchCtx *ctx, const float *d_queries, int nq, int k, const uint64_t *d_accept,
                               int64_t accept_stride_words, int32_t *d_out_doc, float *d_out_score, int32_t *d_out_count,
                               int *launches);
// K5 on the tensor cores (jv_exact_tc.cu): *done = false -> a candidate list overflowed, use the fp32 kernel
bool exact_tc_eligible(const jv_index *ix, int nq, int k, const uint64_t *d_accept);
int32_t launch_exact_topk_tc(jv_index *ix, SearchCtx *ctx, const float *d_queries, int nq, int k, int32_t *d_out_doc, float *d_out_score,
                             int32_t *d_out_count, int *launches, bool *done);

// K7
int32_t launch_merge_topk(cudaStream_t stream, int g, int nq, int k, const int32_t *d_docs, const float *d_scores,
                          int32_t *d_out_doc, float *d_out_score, int32_t *d_out_count);

// K6
int32_t launch_pq_encode(cudaStream_t stream, const PqShape &shape, const float *d_vectors, int64_t n,
                         const float *d_codebooks, const float *d_gcent, uint8_t *d_out, int out_stride);

int32_t launch_pq_decode(cudaStream_t stream, const PqShape &shape, const uint8_t *d_codes, int64_t n, const float *d_codebooks,
                         const float *d_gcent, float *d_out);

// K1 (test hook) + ADC on explicit pairs
int32_t launch_pq_lut(jv_index *ix, cudaStream_t stream, const float *d_queries, int nq, float *d_lut);
int32_t launch_adc_pairs(jv_index *ix, cudaStream_t stream, const float *d_queries, int nq, const int32_t *d_nodes,
                         int per_query, float *d_out);

// index-creation helpers
int32_t launch_vec_norms(cudaStream_t stream, const float *d_vectors, int64_t n, int dim, float *d_out);
int32_t launch_node_norms(cudaStream_t stream, const jv_index *ix, float *d_out);
int32_t launch_build_fused(cudaStream_t stream, jv_index *ix);
int32_t launch_f32_to_f16(cudaStream_t stream, const float *d_in, int64_t n, void *d_out);

}  // namespace jv
