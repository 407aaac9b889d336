/*
 * jvgpu.h — C-ABI of libjvgpu.so, the B200 (sm_100a) implementation of the
 * opensearch-jvector query hot path.
 *
 * This header is the whole drop-in boundary: the Java codec
 * (JVectorReader / JVectorWriter / JVectorIndexQuantization) binds these
 * symbols with java.lang.foreign (Panama FFM) downcall handles — see
 * INTEGRATION.md for the binding a maintainer would add.  Everything is plain
 * pointers, sizes and int32 status codes: no structs by value, no callbacks,
 * no C++ / torch types.
 *
 * Conventions
 *  - every function returns 0 (JV_OK) or a negative jv_status; the message is
 *    available from jv_last_error() (thread local).  Nothing throws or aborts
 *    across the boundary.  There is NO CPU fallback: without a usable CUDA
 *    device every compute entry point returns JV_ERR_CUDA.
 *  - all entry points are re-entrant and thread-safe; an index handle may be
 *    searched from many threads concurrently (reference behaviour:
 *    KNNJVectorTests.java:982-1028).
 *  - the caller owns every in/out buffer; jv_index_create() copies what it
 *    needs to the device and retains no host pointer.
 *  - "host" entry points take host pointers and do their own H2D/D2H; the
 *    "_dev" variants take device pointers that live on the index's device
 *    (used by the multi-GPU driver and by the resident-input benchmark).
 *
 * Reference citations are relative to /root/reference/src/main/java/org/
 * opensearch/knn/index/codec/jvector/ unless a longer path is given.
 */
#ifndef JVGPU_H
#define JVGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define JV_API __attribute__((visibility("default")))
#else
#define JV_API
#endif

#define JVGPU_VERSION_MAJOR 0
#define JVGPU_VERSION_MINOR 1

/* ---- status codes --------------------------------------------------- */
typedef enum jv_status {
    JV_OK = 0,
    JV_ERR_INVALID_ARGUMENT = -1, /* Java side: IllegalArgumentException          */
    JV_ERR_CUDA = -2,             /* Java side: IOException (no device / CUDA err) */
    JV_ERR_OUT_OF_MEMORY = -3,    /* Java side: IOException                        */
    JV_ERR_UNSUPPORTED = -4,      /* Java side: UnsupportedOperationException      */
    JV_ERR_INTERNAL = -5,
    JV_ERR_CORRUPT = -6           /* Java side: CorruptIndexException / IOException (segment files) */
} jv_status;

/* ---- similarity ordinals --------------------------------------------
 * = the ordinals of Lucene's VectorSimilarityFunction (FieldInfo.getVectorSimilarityFunction().ordinal()).  The meta file's
 * simOrd (JVectorReader.java:389-409: distFuncToOrd = indexOf in [EUCLIDEAN, DOT_PRODUCT, COSINE, DOT_PRODUCT]) is 0, 1 or 2:
 * a MAXIMUM_INNER_PRODUCT field is stored as 1, so the loader needs FieldInfo to tell MIP from DOT
 * (jv_segment_set_lucene_similarity).
 *   0 EUCLIDEAN  score = 1/(1+||a-b||^2)
 *   1 DOT        score = (1+a.b)/2
 *   2 COSINE     score = (1+cos)/2
 *   3 MIP        Lucene MAXIMUM_INNER_PRODUCT -> jVector DOT_PRODUCT.  On the
 *                un-quantised traversal (JVectorReader.java:220-239,358-363)
 *                and the brute-force scorer (JVectorVectorScorer.java:43-50)
 *                the score is doubled (= 1+a.b); the PQ reranker is NOT
 *                wrapped (JVectorReader.java:352-356) and stays (1+a.b)/2.
 */
#define JV_SIM_EUCLIDEAN 0
#define JV_SIM_DOT 1
#define JV_SIM_COSINE 2
#define JV_SIM_MIP 3

/* ---- index description ------------------------------------------------
 * Decoded arrays of one field of one segment, i.e. what FieldEntry holds
 * after JVectorReader.java:284-337 loaded OnDiskGraphIndex + PQVectors +
 * GraphNodeIdToDocMap.  All pointers are HOST pointers, read during
 * jv_index_create() only.
 */
#define JV_INDEX_FLAG_FUSED_LAYOUT 1u /* also keep a neighbour-interleaved copy of the PQ codes
                                          (R*M bytes per node) so one expansion is one contiguous read */
#define JV_INDEX_FLAG_LUT_F16 2u      /* hold the per-query ADC table in fp16 in shared memory
                                          (steering scores only; final scores are the exact rerank) */
#define JV_INDEX_FLAG_NO_VECTORS_ON_DEVICE 4u /* keep the fp32 vectors in pinned host memory (cfg 5): the rerank reads them over PCIe;
                                               * large batches gather the rows they need once into HBM (de-duplicated), results unchanged */
#define JV_INDEX_FLAG_LUT_U8 8u      /* production traversal: per-query ADC table quantised to bytes with one scale per
                                          query (integer sums, fused-ADC style), built by a batched kernel and staged
                                          into shared memory with TMA; needs K = 256 and dim % M == 0 with sub-vector
                                          size 2, 4 or 8, otherwise the fp16/fp32 table path is used.  Steering scores
                                          only; final scores are the exact rerank. */

typedef struct jv_index_desc {
    int32_t struct_size;       /* = sizeof(jv_index_desc), for forward compatibility          */
    int32_t similarity;        /* JV_SIM_*                                                      */
    int32_t dim;               /* vector dimension                                              */
    int32_t max_degree;        /* R: row stride of `adjacency`                                  */
    int64_t n;                 /* number of graph nodes (ordinals 0..n-1)                       */
    int32_t entry_node;        /* OnDiskGraphIndex.View.entryNode()                             */
    int32_t max_doc;           /* Lucene maxDoc of the segment (size of accept bitsets)         */
    const int32_t *adjacency;  /* [n * max_degree] level-0 neighbours, -1 padded                */
    const float *vectors;      /* [n * dim] inline fp32 vectors (InlineVectors feature)         */
    const int32_t *ord_to_doc; /* [n] GraphNodeIdToDocMap.getLuceneDocId; NULL = identity; -1 = deleted */
    /* product quantisation (all NULL/0 when the segment is not quantised: n < 1024) */
    int32_t pq_m;              /* number of subspaces M (0 = no PQ)                             */
    int32_t pq_k;              /* centroids per subspace K <= 256                               */
    const float *pq_codebooks; /* concatenated per subspace: [K * sub_m] fp32, sub sizes per A.3 */
    const float *pq_global_centroid; /* [dim] or NULL (only EUCLIDEAN is centred, JVectorIndexQuantization.java:127) */
    const uint8_t *pq_codes;   /* [n * M]                                                       */
    int32_t device;            /* CUDA device ordinal                                           */
    uint32_t flags;            /* JV_INDEX_FLAG_*                                               */
    /* ---- struct_size >= 128: NVQ-inline segments ("nvq+pq", JVectorWriter.java:436-466).  The graph nodes carry
     * 8-bit non-uniform-quantised vectors instead of fp32; the reranker (view.rerankerFor, JVectorReader.java:352-358)
     * scores the dequantised vector: JVectorIndexQuantization.java:316-361.  `vectors` may then be NULL (brute force is
     * unsupported on such segments, as in the reference: JVectorQuantizedNvqVectorValues.java:33-36).  With pq_m = 0 the segment
     * is NVQ-only (JVectorReader.java:357-358): the traversal itself is scored by the NVQ reranker, un-wrapped; the inline
     * vectors are dequantised once on the device (n * dim * 4 bytes) and traversed exactly. */
    int32_t nvq_m;             /* number of NVQ sub-vectors (0 = no NVQ); sizes per the PQ split rule */
    int32_t nvq_reserved;
    const uint8_t *nvq_bytes;  /* [n * dim] quantised components                                */
    const float *nvq_params;   /* [n * nvq_m * 4] growthRate, midpoint, minValue, maxValue per sub-vector */
    const float *nvq_global_mean; /* [dim] NVQuantization.globalMean                            */
} jv_index_desc;
#define JV_INDEX_DESC_SIZE_V1 96 /* struct_size of the layout without the NVQ fields: still accepted */

typedef struct jv_index jv_index; /* opaque */

/* ---- per-query statistics -----------------------------------------------
 * The four counters JVectorReader.java:183-193 feeds into KNNCounter.
 */
typedef struct jv_query_stats {
    int32_t visited;        /* SearchResult.getVisitedCount()            */
    int32_t expanded;       /* SearchResult.getExpandedCount()           */
    int32_t expanded_base;  /* SearchResult.getExpandedCountBaseLayer()  */
    int32_t reranked;       /* SearchResult.getRerankedCount()           */
} jv_query_stats;

/* Device-side durations of the last batch, CUDA events on the launching stream. */
typedef struct jv_batch_timing {
    float h2d_ms;
    float search_ms;  /* LUT build + graph traversal + ADC scoring kernel (K1+K2 / K4) */
    float rerank_ms;  /* exact rerank + top-k kernel (K3)                              */
    float d2h_ms;
    float total_ms;
    int32_t launches; /* kernels launched for this batch                               */
    float lut_ms;     /* share of search_ms spent in the batched 8-bit table build (0 when the table build is fused
                         into the traversal kernel).  jv_search_batch pipelines large host batches in chunks:
                         search_ms / rerank_ms / lut_ms are then the FIRST chunk's, total_ms the whole batch's */
    int32_t expand_width_used; /* candidates expanded per step by the traversal kernel that ran (0: strict kernel)   */
    int32_t traversal_kernel;  /* JV_KERNEL_*: which traversal kernel served the batch                              */
} jv_batch_timing;

enum {
    JV_KERNEL_STRICT = 0,   /* search_kernel: candidate heap + result heap, reference order (jv_search.cu)             */
    JV_KERNEL_FAST = 1,     /* fast_search_kernel: fp32 / fp16 table or exact scores, wide steps (jv_search_fast.cu)   */
    JV_KERNEL_Q8_SYNC = 2,  /* q8_search_kernel: 8-bit table, round-synchronous CTA per query (jv_q8.cu)               */
    JV_KERNEL_Q8_BEAM = 3   /* q8_beam_kernel: 8-bit table, manager + expander + scorer warps, two steps in flight (jv_q8_beam.cu) */
};

typedef struct jv_search_params {
    int32_t struct_size;   /* = sizeof(jv_search_params)                                        */
    int32_t k;             /* topK  = KnnCollector.k()                                           */
    int32_t rerank_k;      /* k * overQueryFactor, JVectorReader.java:168-169; must be >= k      */
    float threshold;       /* JVectorKnnCollector.getThreshold(), default 0                      */
    float rerank_floor;    /* JVectorKnnCollector.getRerankFloor(), default 0                    */
    int32_t expand_width;  /* traversal schedule (GPU-side knob, not part of the reference API):
                            *   0  default: the production kernel expands the 4 best unexpanded candidates per step (6 with accept_bits)
                            *   1..8 explicit width (clamped to what the kernel supports); 1 = the reference's best-first
                            *      order (up to exact score ties)
                            *  -1  strict kernel: candidate heap + result heap exactly as GraphSearcher;
                            *      always used when threshold > 0; with accept_bits it is used unless the index was created
                            *      with JV_INDEX_FLAG_LUT_U8 and 8 * rerank_k <= 1024 (then accepted and rejected nodes share
                            *      one list of 8 * rerank_k entries in the production kernel)                          */
    /* AcceptDocs by Lucene docId (FixedBitSet words: bit d = word d>>6, bit d&63), NULL = accept all.
     * accept_stride_words == 0: one bitset shared by the batch; else query i uses
     * accept_bits + i*accept_stride_words.  An ordinal is accepted iff ord_to_doc[ord] != -1 and its
     * doc bit is set (JVectorReader.java:157-163).  Rejected nodes are still traversed. */
    const uint64_t *accept_bits;
    int64_t accept_stride_words;
} jv_search_params;

/* ---- library ---------------------------------------------------------- */
JV_API int32_t jv_version(void);                 /* (major << 16) | minor */
JV_API const char *jv_last_error(void);          /* thread-local, never NULL */
JV_API int32_t jv_device_count(int32_t *out_count);

/* ---- page-locked host buffers ---------------------------------------------------------------------------------------
 * jv_search_batch / jv_exact_topk / jv_pq_encode take HOST pointers (what JVectorReader.search has in hand, JVectorReader.java:147:
 * the query float[] and the collector's result arrays).  Copies from pageable memory are staged by the driver and cannot overlap the
 * kernels; from page-locked memory they are asynchronous DMA and the chunked H2D pipeline of jv_search_batch overlaps them with the
 * traversal.  A Panama caller gets page-locked memory in one of two ways (INTEGRATION.md "Host buffers"):
 *   jv_host_alloc      cudaHostAlloc'ed buffer (portable across devices); wrap it with MemorySegment.reinterpret(bytes)
 *   jv_host_register   page-lock memory the JVM already owns (an Arena-allocated off-heap segment); the range must stay
 *                      allocated until jv_host_unregister.  Never register Java heap arrays (the GC moves them).
 * Both are optional: every entry point also accepts pageable pointers. */
JV_API int32_t jv_host_alloc(int64_t bytes, void **out_ptr);
JV_API int32_t jv_host_free(void *ptr);
JV_API int32_t jv_host_register(void *ptr, int64_t bytes);
JV_API int32_t jv_host_unregister(void *ptr);

/* ---- index lifetime: FieldEntry ctor / close, JVectorReader.java:284-337, 367-378 ---- */
JV_API int32_t jv_index_create(const jv_index_desc *desc, jv_index **out_index);
JV_API int32_t jv_index_destroy(jv_index *index);
JV_API int32_t jv_index_device_bytes(const jv_index *index, int64_t *out_bytes);
/* diagnostics since index creation: which = 0 -> queries whose shared-memory visited set filled up in the strict
 * kernel (those searches stop admitting new nodes early; results stay valid but recall may drop);
 * which = 8..15 -> SM cycles the fast kernel spent per phase (setup, table build, select, neighbour rows, scoring,
 * merge, emit, steps), summed over CTAs; which = 100 resets all counters; which = 200 re-reads the JVGPU_* diagnostic
 * environment knobs, which are otherwise read once per process */
JV_API int32_t jv_index_debug_counter(jv_index *index, int32_t which, int64_t *out_value);

/* ---- K1+K2(+K4)+K3: replaces the body of JVectorReader.search, JVectorReader.java:130-210 ----
 * queries [nq*dim] fp32; out_doc/out_score [nq*k] sorted by (score desc, doc asc), unused slots
 * doc=-1, score=0; out_count [nq]; stats [nq] (nullable); timing (nullable).  nq = 1 is legal. */
JV_API int32_t jv_search_batch(jv_index *index, const float *queries, int32_t nq, const jv_search_params *params,
                        int32_t *out_doc, float *out_score, int32_t *out_count, jv_query_stats *stats,
                        jv_batch_timing *timing);
/* same, device pointers on the index's device (params->accept_bits is a device pointer too) */
JV_API int32_t jv_search_batch_dev(jv_index *index, const float *d_queries, int32_t nq, const jv_search_params *params,
                            int32_t *d_out_doc, float *d_out_score, int32_t *d_out_count, jv_query_stats *d_stats,
                            jv_batch_timing *timing);

/* ---- K5: brute-force exact top-k; replaces JVectorVectorScorer.score() driven by Lucene exactSearch,
 * JVectorVectorScorer.java:36-53, JVectorFloatVectorValues.java:189-191.  Scores are jVector-scaled
 * (MIP doubled).  Ties -> lower docId.  accept bits as in jv_search_params (nullable). */
JV_API int32_t jv_exact_topk(jv_index *index, const float *queries, int32_t nq, int32_t k, const uint64_t *accept_bits,
                      int64_t accept_stride_words, int32_t *out_doc, float *out_score, int32_t *out_count);
JV_API int32_t jv_exact_topk_dev(jv_index *index, const float *d_queries, int32_t nq, int32_t k, const uint64_t *d_accept_bits,
                          int64_t accept_stride_words, int32_t *d_out_doc, float *d_out_score, int32_t *d_out_count);

/* ---- K6: PQ encode; replaces PQVectors.encodeAndBuild, JVectorIndexQuantization.java:133 and
 * JVectorWriter.java:1124.  codebooks as in jv_index_desc; global_centroid nullable.
 * out_codes [n*m]; code = first argmin_c ||x_m - C_m[c]||^2 (strict <). */
JV_API int32_t jv_pq_encode(int32_t device, const float *vectors, int64_t n, int32_t dim, int32_t m, int32_t k,
                     const float *codebooks, const float *global_centroid, uint8_t *out_codes, float *out_kernel_ms);
JV_API int32_t jv_pq_encode_dev(int32_t device, const float *d_vectors, int64_t n, int32_t dim, int32_t m, int32_t k,
                         const float *d_codebooks, const float *d_global_centroid, uint8_t *d_out_codes,
                         float *out_kernel_ms);

/* ---- K1 alone (test hook for the ADC table): PQVectors.precomputedScoreFunctionFor, JVectorReader.java:354.
 * out_lut [nq * m * k] fp32: dot (DOT/COSINE/MIP) or squared L2 of the (centred) query sub-vector. */
JV_API int32_t jv_pq_lut(jv_index *index, const float *queries, int32_t nq, float *out_lut);
/* K1, 8-bit flavour (JV_INDEX_FLAG_LUT_U8 indexes; test hook): out_q8 [nq * m * 256] bytes in logical (m, c) order,
 * out_params [nq * 2] = (delta, base): partial-sum estimate = delta * sum_m q8[m][code_m] + base. */
JV_API int32_t jv_pq_lut_q8(jv_index *index, const float *queries, int32_t nq, uint8_t *out_q8, float *out_params);
/* ADC scores of explicit (query, node) pairs through the same device code as the traversal. */
JV_API int32_t jv_pq_adc_scores(jv_index *index, const float *queries, int32_t nq, const int32_t *nodes, int32_t nodes_per_query,
                         float *out_scores);

/* ---- K7: merge per-shard top-k lists (Lucene TopDocs.merge analogue).  lists are laid out
 * [g][nq][k] with doc=-1 padding; docs must already be global ids.  Ties -> lower doc. */
JV_API int32_t jv_merge_topk(int32_t device, int32_t g, int32_t nq, int32_t k, const int32_t *docs, const float *scores,
                      int32_t *out_doc, float *out_score, int32_t *out_count);
/* jv_merge_topk_stream: the same merge enqueued on the caller's CUDA stream (cudaStream_t passed as void*, NULL = the legacy
 * default stream) without a host synchronisation, so that the exchange + merge of one batch overlaps the search of the next. */
JV_API int32_t jv_merge_topk_stream(int32_t device, int32_t g, int32_t nq, int32_t k, const int32_t *d_docs, const float *d_scores,
                             int32_t *d_out_doc, float *d_out_score, int32_t *d_out_count, void *cuda_stream);
JV_API int32_t jv_merge_topk_dev(int32_t device, int32_t g, int32_t nq, int32_t k, const int32_t *d_docs, const float *d_scores,
                          int32_t *d_out_doc, float *d_out_score, int32_t *d_out_count, float *out_kernel_ms);

/* ---- "next" rows: fixtures built on the device ---------------------------
 * PQ codebook training: ProductQuantization.compute(ravv, M, K, center, UNWEIGHTED, ...),
 * JVectorIndexQuantization.java:123-131: k-means++ init + `iters` Lloyd iterations per subspace on
 * the given training sample.  out_codebooks as in jv_index_desc; out_global_centroid [dim] written
 * iff center != 0. */
JV_API int32_t jv_pq_train(int32_t device, const float *vectors, int64_t n, int32_t dim, int32_t m, int32_t k, int32_t center,
                    int32_t iters, uint64_t seed, float *out_codebooks, float *out_global_centroid);
JV_API int32_t jv_pq_train_dev(int32_t device, const float *d_vectors, int64_t n, int32_t dim, int32_t m, int32_t k,
                        int32_t center, int32_t iters, uint64_t seed, float *d_out_codebooks,
                        float *d_out_global_centroid);

/* Vamana graph construction: GraphIndexBuilder(bsp, dim, M, beamWidth, neighborOverflow, alpha, false),
 * JVectorWriter.java:1383-1422.  Batched-insert variant (prefix doubling) with exact build scores.
 * out_adjacency [n*max_degree] (-1 padded), out_entry_node. */
JV_API int32_t jv_graph_build(int32_t device, const float *vectors, int64_t n, int32_t dim, int32_t similarity,
                       int32_t max_degree, int32_t beam_width, float neighbor_overflow, float alpha,
                       int32_t *out_adjacency, int32_t *out_entry_node);
JV_API int32_t jv_graph_build_dev(int32_t device, const float *d_vectors, int64_t n, int32_t dim, int32_t similarity,
                           int32_t max_degree, int32_t beam_width, float neighbor_overflow, float alpha,
                           int32_t *d_out_adjacency, int32_t *out_entry_node);

/* Leading-segment merge, insert-only part (JVectorWriter.tryLeadingSegmentMerge, JVectorWriter.java:1166-1341): the first n0
 * ordinals keep the leading segment's graph (seed_adjacency [n0*max_degree], -1 padded at the end of each row; its entry node
 * stays the entry), the cached neighbour scores are recomputed (exact pair scores, what the neighbours-score-cache file holds),
 * and vectors n0..n-1 are inserted like builder.addGraphNode with the batched schedule of jv_graph_build.  Deleted documents
 * of the leading segment (builder.markNodeDeleted + cleanup, :1318-1327) are consolidated afterwards with
 * jv_graph_remove_deleted in the same ordinal space.  out_adjacency [n*max_degree]. */
JV_API int32_t jv_graph_extend(int32_t device, const float *vectors, int64_t n, int64_t n0, const int32_t *seed_adjacency,
                        int32_t seed_entry, int32_t dim, int32_t similarity, int32_t max_degree, int32_t beam_width,
                        float neighbor_overflow, float alpha, int32_t *out_adjacency);
JV_API int32_t jv_graph_extend_dev(int32_t device, const float *d_vectors, int64_t n, int64_t n0, const int32_t *d_seed_adjacency,
                            int32_t seed_entry, int32_t dim, int32_t similarity, int32_t max_degree, int32_t beam_width,
                            float neighbor_overflow, float alpha, int32_t *d_out_adjacency);

/* Delete consolidation: builder.markNodeDeleted(...) + builder.cleanup() of a merge (JVectorWriter.java:1318-1327), i.e.
 * jVector's GraphIndexBuilder.removeDeletedNodes = FreshDiskANN section 4.2: every live node with a deleted out-neighbour j gets
 * j's live out-neighbours as candidates next to its own live ones, scored exactly, sorted best first and re-pruned with
 * retainDiverse(max_degree, alpha).  deleted [n] (0/1 bytes).  Same ordinal space in and out (rows of deleted nodes come back
 * empty; the writer compacts ordinals afterwards).  A deleted entry node is replaced by the best-scoring live candidate around
 * it, else the lowest live ordinal, else -1 (jVector picks an approximate medoid: documented deviation). */
JV_API int32_t jv_graph_remove_deleted(int32_t device, const float *vectors, int64_t n, int32_t dim, int32_t similarity,
                                int32_t max_degree, float alpha, const int32_t *adjacency, const uint8_t *deleted,
                                int32_t entry_node, int32_t *out_adjacency, int32_t *out_entry_node);
JV_API int32_t jv_graph_remove_deleted_dev(int32_t device, const float *d_vectors, int64_t n, int32_t dim, int32_t similarity,
                                    int32_t max_degree, float alpha, const int32_t *d_adjacency, const uint8_t *d_deleted,
                                    int32_t entry_node, int32_t *d_out_adjacency, int32_t *out_entry_node);

/* ---- PQ decode + graph build with PQ build scores ------------------------------------------------------------------
 * jv_pq_decode: ProductQuantization.decode — out[row] = concatenated centroids of the row's codes (+ global centroid).
 * jv_graph_build_pq: GraphIndexBuilder driven by BuildScoreProvider.pqBuildScoreProvider (JVectorIndexQuantization.java:408,
 * JVectorWriter.java:238-244 at flush, :1143-1151 when a merge rebuilds from scratch): every build-time score is a score between PQ
 * reconstructions; the fp32 vectors are not read.  Same schedule, pruning and outputs as jv_graph_build. */
JV_API int32_t jv_pq_decode(int32_t device, const uint8_t *codes, int64_t n, int32_t dim, int32_t m, int32_t k, const float *codebooks,
                     const float *global_centroid, float *out_vectors);
JV_API int32_t jv_pq_decode_dev(int32_t device, const uint8_t *d_codes, int64_t n, int32_t dim, int32_t m, int32_t k,
                         const float *d_codebooks, const float *d_global_centroid, float *d_out_vectors);
JV_API int32_t jv_graph_build_pq(int32_t device, const uint8_t *codes, int64_t n, int32_t dim, int32_t similarity, int32_t pq_m,
                          int32_t pq_k, const float *codebooks, const float *global_centroid, int32_t max_degree,
                          int32_t beam_width, float neighbor_overflow, float alpha, int32_t *out_adjacency, int32_t *out_entry_node);
JV_API int32_t jv_graph_build_pq_dev(int32_t device, const uint8_t *d_codes, int64_t n, int32_t dim, int32_t similarity, int32_t pq_m,
                              int32_t pq_k, const float *d_codebooks, const float *d_global_centroid, int32_t max_degree,
                              int32_t beam_width, float neighbor_overflow, float alpha, int32_t *d_out_adjacency,
                              int32_t *out_entry_node);

/* ---- "next" row SURVEY 8f-1: segment-file loader ------------------------------------------------
 * Reads the files JVectorWriter persists straight into the decoded arrays of a jv_index_desc, so that
 * the Java side hands over two paths instead of extracting arrays through jVector's API:
 *   <seg>_<sfx>.meta-jvector          CodecUtil index header "JVectorVectorsFormatMeta" (JVectorFormat.java:23,31-33,
 *                                     JVectorWriter.java:134-157) · repeat{ int fieldNumber · VectorIndexFieldMetadata
 *                                     (JVectorWriter.java:528-540) } · int -1 · CodecUtil footer (:573-577);
 *                                     parsed like JVectorReader.java:52-81,255-262 incl. the v0 rule for the missing
 *                                     quantisation-type byte (JVectorWriter.java:551-558) and the doc map
 *                                     (GraphNodeIdToDocMap.java:39-59)
 *   <seg>_<sfx>_<field>.data-jvector  CodecUtil index header "JVectorVectorsFormatIndex" · OnDiskGraphIndex bytes at
 *                                     [indexOffset, +indexLength) · optional PQVectors blob at [pqOffset, +pqLength) ·
 *                                     footer (JVectorWriter.java:383-433,469-510; read at JVectorReader.java:306-331)
 * Lucene framing (magic, version, CRC-32) is big-endian, everything that goes through IndexOutput.writeInt/Long is
 * little-endian (JVectorIndexWriter.java:72-83).  The byte layout INSIDE the two jVector blobs belongs to the un-vendored
 * jar (jvector 4.0.0-rc.9) and is restated from its published format: it has not been checked
 * against a file written by the real plugin, hence the flags below for the open questions of SURVEY B.3.
 * No GPU is needed up to jv_field_data_desc(); jv_segment_index_create() = load + jv_index_create(). */
#define JV_SEGMENT_FLAG_FLOATS_BIG_ENDIAN 1u /* inline vectors / codebooks were written through a big-endian bulk path
                                                 (B.3-1); default: little-endian, what JVectorIndexWriter.writeFloat emits */
#define JV_SEGMENT_FLAG_VERIFY_DATA_CRC 2u   /* also CRC the whole field data file on load (the reference does that only
                                                 in checkIntegrity, JVectorReader.java:87-99) */
#define JV_SEGMENT_FLAG_LENIENT_MAGIC 4u     /* do not fail on unexpected jVector blob magic numbers (B.3-2) */

typedef struct jv_field_meta {        /* one VectorIndexFieldMetadata record */
    int32_t struct_size;              /* = sizeof(jv_field_meta), set by the caller */
    int32_t field_number;
    int32_t vector_encoding;          /* Lucene VectorEncoding ordinal: 0 BYTE, 1 FLOAT32 */
    int32_t similarity;               /* JV_SIM_*: the record's simOrd (0..2; MIP fields read 1 = DOT) until
                                       * jv_segment_set_lucene_similarity() overrides it from FieldInfo */
    int32_t dim;
    int32_t quantization_type;        /* 0 none, 1 PQ, 2 NVQ-inline (JVectorIndexQuantization.QUANTIZATION_TYPE_*) */
    int64_t index_offset, index_length; /* OnDiskGraphIndex bytes inside the field data file */
    int64_t pq_offset, pq_length;     /* PQVectors blob, 0/0 when absent */
    float degree_overflow;
    int32_t graph_nodes;              /* GraphNodeIdToDocMap size */
    int32_t max_doc;                  /* GraphNodeIdToDocMap maxDocs */
    int32_t format_version;           /* version of the meta file's index header (0 or 1) */
} jv_field_meta;

typedef struct jv_segment jv_segment;       /* parsed meta file */
typedef struct jv_field_data jv_field_data; /* decoded arrays of one field data file (host memory) */

JV_API int32_t jv_segment_open(const char *meta_path, uint32_t flags, jv_segment **out_segment);
JV_API int32_t jv_segment_close(jv_segment *segment);
JV_API int32_t jv_segment_field_count(const jv_segment *segment, int32_t *out_count);
JV_API int32_t jv_segment_field_meta(const jv_segment *segment, int32_t i, jv_field_meta *out_meta);
/* FieldInfo.getVectorSimilarityFunction() of field i (JV_SIM_* = the Lucene enum ordinal).  Needed for MAXIMUM_INNER_PRODUCT
 * fields, whose meta record says DOT_PRODUCT: the x2 wrap of the un-quantised traversal and of the brute-force scorer
 * (JVectorReader.java:220-239, JVectorVectorScorer.java:43-50) depends on it.  Fails when it contradicts the record. */
JV_API int32_t jv_segment_set_lucene_similarity(jv_segment *segment, int32_t i, int32_t lucene_similarity);
/* ordinal -> Lucene docId, -1 for deleted ordinals; capacity >= graph_nodes */
JV_API int32_t jv_segment_field_doc_map(const jv_segment *segment, int32_t i, int32_t *out_ord_to_doc, int32_t capacity);
JV_API int32_t jv_segment_load_field(const jv_segment *segment, int32_t i, const char *field_data_path, uint32_t flags,
                              jv_field_data **out_data);
/* fills every array pointer / shape of *out_desc (device = 0, flags = 0: the caller sets those); the pointers stay
 * valid until jv_field_data_free() */
JV_API int32_t jv_field_data_desc(const jv_field_data *data, jv_index_desc *out_desc);
JV_API int32_t jv_field_data_free(jv_field_data *data);
/* FieldEntry ctor in one call (JVectorReader.java:284-337): load, create the device index, drop the host arrays */
JV_API int32_t jv_segment_index_create(const jv_segment *segment, int32_t i, const char *field_data_path, int32_t device,
                                uint32_t index_flags, uint32_t load_flags, jv_index **out_index);
/* CodecUtil.checksumEntireFile of one file (JVectorReader.checkIntegrity, JVectorReader.java:87-99) */
JV_API int32_t jv_file_check_integrity(const char *path);

#ifdef __cplusplus
}
#endif
#endif /* JVGPU_H */
