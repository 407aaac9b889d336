// jv_q8.cuh — pieces shared by the two traversal kernels of the 8-bit table path: jv_q8.cu (K1 + the round-synchronous
// kernel: filtered queries, lists longer than 64) and jv_q8_beam.cu (manager / scorer warps, the production kernel).
#pragma once
#include <stdlib.h>

#include <type_traits>

#include "jv_search_common.cuh"

namespace jv {

// ---------------------------------------------------------------------------------------------------------------
// small PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { // release.cta: the arriving thread's earlier writes are visible to the waiter
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
template <int BYTES> __device__ __forceinline__ void cp_async(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(dst)), "l"(src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------------------------------------------------------
// code rows (codes_q8): lane sl (0..7) of a row group owns subspaces m = 8t + sl, t = 0 .. 4*NJ-1, as NJ 4-byte words.
// The words are stored in chunks of 4, 2 and 1 words per lane so that every lane fetches its part of a row with 16-byte
// (then 8-, then 4-byte) vector loads and the 8 lanes of a group cover each chunk contiguously:
//     NJ = 6 (M = 192):  [8 lanes x 16 B][8 lanes x 8 B]        NJ = 3:  [8 x 8 B][8 x 4 B]        NJ = 8:  [8 x 16 B][8 x 16 B]
// ---------------------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int q8_word_offset(int NJ, int sl, int j) {
    const int n4 = NJ & ~3;
    if (j < n4) return 32 * (j & ~3) + sl * 16 + 4 * (j & 3);
    if (NJ - n4 >= 2 && j < n4 + 2) return 32 * n4 + sl * 8 + 4 * (j - n4);
    return 32 * j + sl * 4;
}

template <int NJ> __device__ __forceinline__ void q8_load_row(const unsigned char *row, int sl, uint32_t (&cw)[NJ]) {
    constexpr int n4 = NJ & ~3, rem = NJ - n4;
#pragma unroll
    for (int j0 = 0; j0 < n4; j0 += 4)
        asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(cw[j0]), "=r"(cw[j0 + 1]), "=r"(cw[j0 + 2]), "=r"(cw[j0 + 3])
                     : "l"(row + 32 * j0 + sl * 16));
    if constexpr (rem >= 2)
        asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(cw[n4]), "=r"(cw[n4 + 1]) : "l"(row + 32 * n4 + sl * 8));
    if constexpr (rem & 1) asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(cw[NJ - 1]) : "l"(row + 32 * (NJ - 1) + sl * 4));
}

// ---------------------------------------------------------------------------------------------------------------
// traversal parameters and list keys
// ---------------------------------------------------------------------------------------------------------------
constexpr int kQMaxE = 8; // candidates per step (synchronous kernel) / pipelined expansions in flight

struct Q8Params {
    const int32_t *adjacency;
    const uint8_t *codes_q8;
    const float *node_norm;
    const uint8_t *lut;      // [nq][lutb]
    const float4 *qparams;   // [nq]
    uint64_t *approx_keys;   // [nq][L]
    int32_t *approx_count;
    jv_query_stats *stats;
    int *work_counter;
    int *dbg;
    int64_t n;
    int nq, L, R, entry, sim, NJ, lutb, hash_log2, E, surv_cap;
    // filtered queries (FILT instantiation): accept bits by Lucene docId (JVectorReader.java:157-163); Lc = list capacity
    int Lc;
    const uint64_t *accept;
    int64_t accept_stride;
    // fused K3 (exact rerank + top-k as the epilogue of every query; fuse_k = 0: write approx_keys for a separate rerank kernel)
    int fuse_k, dim;
    float rerank_floor;
    const float *queries, *vectors, *vec_norm;
    const int32_t *ord_to_doc;
    int32_t *out_doc, *out_count;
    float *out_score;
};

// list key: order word (32 bits) | (0x7fffffff - node) << 1 | unexpanded.  The order word is the integer ADC sum itself
// for DOT/MIP (larger = better), its complement for EUCLIDEAN, and the ordered-float score for COSINE (the cosine
// decoder divides by the node norm, so its order is not the order of the sums).
__device__ __forceinline__ uint64_t qkey_pack(uint32_t ord, int32_t node) {
    return ((uint64_t)ord << 32) | ((uint64_t)(uint32_t)(0x7fffffff - node) << 1) | 1ull;
}
__device__ __forceinline__ int32_t qkey_node(uint64_t k) { return 0x7fffffff - (int32_t)((k >> 1) & 0x7fffffffu); }
// filtered flavour: 30-bit node field, bit 1 = "accepted by the filter" (the lowest bit of the comparable part key >> 1; it
// never decides an order because (order word, node) is already unique), bit 0 = unexpanded
__device__ __forceinline__ uint64_t qkey_pack_f(uint32_t ord, int32_t node, bool acc) {
    return ((uint64_t)ord << 32) | ((uint64_t)(uint32_t)(0x3fffffff - node) << 2) | (acc ? 2ull : 0ull) | 1ull;
}
__device__ __forceinline__ int32_t qkey_node_f(uint64_t k) { return 0x3fffffff - (int32_t)((k >> 2) & 0x3fffffffu); }

// visited filter: true when `nb` was NOT present (and records it).  2 tags of 15 bits + valid bit per word; (set, tag) is
// a bijection of the ordinal when n <= 2^(set_bits+15), so there are no false positives; evictions only cause re-scoring.
__device__ __forceinline__ bool q_filter_insert(uint32_t *filter, int set_bits, bool tagged, int32_t nb) {
    if (tagged) {
        const uint32_t x = ((uint32_t)nb * 0x9E3779B1u) & ((1u << (set_bits + 15)) - 1u);
        const uint32_t set = x >> 15, tag = (x & 0x7fffu) | 0x8000u;
        uint32_t old = filter[set];
        for (;;) {
            if ((old & 0xffffu) == tag || (old >> 16) == tag) return false;
            const uint32_t seen = atomicCAS(&filter[set], old, (old << 16) | tag);
            if (seen == old) return true;
            old = seen;
        }
    } else {
        const uint32_t h = ((uint32_t)nb * 2654435761u) >> (32 - set_bits);
        return atomicExch(&filter[h], (uint32_t)nb) != (uint32_t)nb;
    }
}

// jv_q8_beam.cu
bool q8_beam_supported(const jv_index *ix, int L, int R, int E);
int32_t launch_q8_beam(jv_index *ix, SearchCtx *ctx, Q8Params &p);

}  // namespace jv
