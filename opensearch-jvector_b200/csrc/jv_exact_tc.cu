// jv_exact_tc.cu — K5 on the tensor cores: brute-force exact top-k as a two-stage kernel.
//
// Reference: JVectorVectorScorer.score() driven by Lucene's exactSearch (JVectorVectorScorer.java:36-53,
// JVectorFloatVectorValues.java:189-191) — an [nq x dim] x [dim x n] contraction followed by a top-k per query.
//
// Stage 1 (tensor cores, bf16 operands, fp32 accumulation in TMEM) produces, per query, a SUPERSET of the true top-k; stage 2
// re-scores that superset with the canonical fp32 reduction of the plain kernel (jv_rerank.cu exact_kernel) and ranks it by
// (score, docId), so ids and score bits are identical to the fp32 path.  The superset is guaranteed by a rigorous error bound:
//   |bf16(q).bf16(x) accumulated in fp32  -  q.x|  <=  eps = c * ||q|| * max||x||,   c = 2^-8 (two roundings to 8 significand
//   bits) + dim * 2^-23 (accumulation) + slack = 0.0045
//   - pass A (a strided sample of the vectors): per query the k-th largest of the per-32-column maxima m_k; k distinct vectors
//     have an approximate score >= m_k, hence the true k-th best score is >= m_k - eps and every member of the true top-k has an
//     approximate score >= m_k - 2 eps  =: thr
//   - pass B (all vectors): every (query, vector) with approximate score >= thr is appended to the query's candidate list
//   - select: the k-th largest approximate score t_k of the list is the k-th largest of ALL vectors; by the same argument every
//     member of the true top-k has an approximate score >= t_k - 2 eps; those (typically k + a few) are re-scored exactly.
//   A candidate list that overflows (adversarial data: thousands of vectors within 2 eps of the k-th) sets a flag and the
//   batch is answered by the fp32 kernel instead — correctness never depends on the bound being tight.
//
// The GEMM kernel is hand-written tcgen05: one CTA per (128-query tile, slice of vectors), persistent over the 256-vector
// tiles of its slice; warp 0 = TMA producer (cp.async.bulk.tensor, 128-byte swizzle, 4-stage mbarrier pipeline of 64-wide K
// blocks), warp 1 = MMA issuer (one elected thread: tcgen05.mma.cta_group::1.kind::f16, M = 128, N = 256, K = 16 per
// instruction, 4 per K block; tcgen05.commit releases the stage / publishes the accumulator), warps 2..5 = epilogue
// (tcgen05.ld 32x32b.x32: thread = query row, 32 columns per load; threshold test in registers).  The accumulator is double
// buffered in TMEM (2 x 256 columns) so the epilogue of tile t overlaps the MMAs of tile t + 1.
//
// Used for unfiltered brute force on device-resident fp32 vectors when the contraction is large enough to pay for the bf16
// copy of the vectors (made once per index, lazily) — or always with JVGPU_EXACT_TC=1 (tests); JVGPU_EXACT_TC=0 disables it.
#include <cuda.h>
#include <cuda_bf16.h>

#include "jv_q8.cuh"

namespace jv {

constexpr int kTcBM = 128, kTcBN = 256, kTcBK = 64, kTcStages = 4;
constexpr int kTcStageBytes = (kTcBM + kTcBN) * kTcBK * 2; // 48 KB
constexpr int kTcThreads = 192;
constexpr int kTcCap = 4096;     // candidates per query (pass B)
constexpr int kTcCap2 = 1024;    // candidates per query that are re-scored exactly
constexpr float kTcEpsFactor = 0.0045f;

struct TcParams {
    const float *bias;         // [n] additive term per vector (EUCLIDEAN: -||x||^2 / 2), nullable
    const int32_t *ord_to_doc; // nullable; < 0 = deleted
    const float *thr;          // [nq] (pass B)
    float *chunkmax;           // [nq][nchunks] (pass A)
    unsigned long long *cand;  // [nq][kTcCap]  (approx ordered-float << 32 | ordinal)
    int *cand_count;           // [nq]
    int64_t n_cols;            // columns of the (possibly strided) view
    int64_t col_stride;        // ordinal = column * col_stride
    int64_t cols_per_cta;      // multiple of 256
    int nq, kblocks, nchunks;
};

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32 (both operands K-major)
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout, sm_100): K-major tile of rows x 64 bf16 written by TMA with
// the 128-byte swizzle: start address >> 4 in bits [0,14), leading byte offset (unused for one swizzle atom along K) 0, stride
// byte offset = 8 rows * 128 B = 1024 (>> 4) in bits [32,46), descriptor version 1 in bits [46,48), layout SWIZZLE_128B = 2 in
// bits [61,64)
__device__ __forceinline__ uint64_t tc_smem_desc(const void *tile) {
    const uint32_t lo = (smem_u32(tile) >> 4) & 0x3fffu;
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = BF16 (1 << 7, 1 << 10), both K-major, N >> 3 in
// bits [17,23), M >> 4 in bits [24,29)
constexpr uint32_t kTcIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTcBN >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);

#define JV_TC_LD32(taddr, v)                                                                                                         \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, " \
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                                  \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),        \
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),            \
                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),           \
                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                         \
                 : "r"(taddr))

// pass B: (query q, column) reached the threshold
__device__ __noinline__ void tc_append(unsigned long long *cand, int *cand_count, const int32_t *ord_to_doc, int64_t n_cols, int64_t col_stride, int q,
                                       float s, int64_t cj) {
    if (cj >= n_cols) return;
    const int64_t ord = cj * col_stride;
    if (ord_to_doc != nullptr && __ldg(ord_to_doc + ord) < 0) return; // deleted
    const int slot = atomicAdd(cand_count + q, 1);
    if (slot < kTcCap) cand[(int64_t)q * kTcCap + slot] = ((unsigned long long)jv_f2ord(s) << 32) | (unsigned long long)(uint32_t)ord;
}

// MODE 0: per-(query, 32 columns) maxima of the approximate scores (pass A).  MODE 1: append (approx score, ordinal) of every
// column whose approximate score reaches thr[query] (pass B).  BIAS: add bias[ordinal] to the dot product (EUCLIDEAN).
template <int MODE, bool BIAS>
__global__ void __launch_bounds__(kTcThreads, 1)
exact_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
    extern __shared__ unsigned char tc_smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kTcStages * kTcStageBytes);
    uint64_t *full = bars, *empty = bars + kTcStages, *tfull = bars + 2 * kTcStages, *tempty = bars + 2 * kTcStages + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kTcStages + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qtile = blockIdx.x;
    const int64_t col0 = (int64_t)blockIdx.y * p.cols_per_cta;
    int64_t span = p.n_cols - col0;
    if (span > p.cols_per_cta) span = p.cols_per_cta;
    const int ntiles = span > 0 ? (int)((span + kTcBN - 1) / kTcBN) : 0;
    const int KB = p.kblocks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kTcStages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(&tfull[b], 1);
            mbar_init(&tempty[b], 4); // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) { // TMEM: 512 columns = two 128 x 256 fp32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ============================================================================================ TMA producer
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
            int stage = 0;
            uint32_t phase = 0;
            for (int t = 0; t < ntiles; t++) {
                const int row_b = (int)(col0 + (int64_t)t * kTcBN);
                for (int kb = 0; kb < KB; kb++) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    unsigned char *a = smem + stage * kTcStageBytes, *b = a + kTcBM * kTcBK * 2;
                    mbar_expect_tx(&full[stage], kTcStageBytes);
                    tma_load_2d(a, &tmA, kb * kTcBK, qtile * kTcBM, &full[stage]);
                    tma_load_2d(b, &tmB, kb * kTcBK, row_b, &full[stage]);
                    if (++stage == kTcStages) stage = 0, phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ============================================================================================ MMA issuer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = 0; t < ntiles; t++) {
                const int buf = t & 1;
                const uint32_t tph = (uint32_t)(t >> 1) & 1u;
                mbar_wait(&tempty[buf], tph ^ 1u); // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)buf * kTcBN;
                for (int kb = 0; kb < KB; kb++) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const unsigned char *a = smem + stage * kTcStageBytes, *b = a + kTcBM * kTcBK * 2;
                    const uint64_t ad = tc_smem_desc(a), bd = tc_smem_desc(b);
#pragma unroll
                    for (int k = 0; k < kTcBK / 16; k++) // 16 bf16 = 32 bytes along K inside the swizzle atom: + 2 in the address field
                        tc_mma(d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), kTcIdesc, (kb | k) != 0 ? 1u : 0u);
                    tc_commit(&empty[stage]); // the stage is free once these MMAs have read it
                    if (++stage == kTcStages) stage = 0, phase ^= 1u;
                }
                tc_commit(&tfull[buf]); // accumulator complete
            }
        }
    } else {
        // ============================================================================================ epilogue
        const int quarter = warp & 3; // a warp may touch TMEM lanes 32 * (warp % 4) .. + 31
        const int row = quarter * 32 + lane;
        const int q = qtile * kTcBM + row;
        const bool qok = q < p.nq;
        const float thr = (MODE == 1 && qok) ? __ldg(p.thr + q) : 0.f;
        for (int t = 0; t < ntiles; t++) {
            const int buf = t & 1;
            const uint32_t tph = (uint32_t)(t >> 1) & 1u;
            if (lane == 0) mbar_wait(&tfull[buf], tph); // one lane polls, the warp follows
            __syncwarp();
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < kTcBN / 32; c++) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * kTcBN + c * 32);
                JV_TC_LD32(taddr, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int64_t col = col0 + (int64_t)t * kTcBN + c * 32;
                float myb = 0.f;
                if (BIAS) {
                    const int64_t cj = col + lane;
                    myb = cj < p.n_cols ? __ldg(p.bias + cj * p.col_stride) : 0.f;
                }
                if (MODE == 0) {
                    const bool edge = col + 32 > p.n_cols || p.ord_to_doc != nullptr;
                    float m = -INFINITY;
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        float s = __uint_as_float(v[j]);
                        if (BIAS) s += __shfl_sync(JV_FULL_MASK, myb, j);
                        if (edge) {
                            const int64_t cj = col + j;
                            const bool live = cj < p.n_cols && (p.ord_to_doc == nullptr || __ldg(p.ord_to_doc + cj * p.col_stride) >= 0);
                            if (!live) s = -INFINITY;
                        }
                        m = fmaxf(m, s);
                    }
                    if (qok && col < p.n_cols) p.chunkmax[(int64_t)q * p.nchunks + (col >> 5)] = m;
                } else {
                    // one compare + (rarely taken) branch per element: about one (query, vector) pair in a thousand is a candidate
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        float s = __uint_as_float(v[j]);
                        if (BIAS) s += __shfl_sync(JV_FULL_MASK, myb, j);
                        if (s >= thr && qok) tc_append(p.cand, p.cand_count, p.ord_to_doc, p.n_cols, p.col_stride, q, s, col + j);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

// fp32 rows -> bf16 rows padded to Dp (zero fill); NORMALISE: divide by the canonical norm first (COSINE).  Also writes the bias
// of the row (EUCLIDEAN: -||x||^2 / 2) and folds max ||x||^2 into *max_norm2 (float bits, values >= 0).
template <bool NORMALISE>
__global__ void tc_convert_rows_kernel(const float *__restrict__ src, int64_t n, int dim, int Dp, __nv_bfloat16 *__restrict__ dst,
                                       float *__restrict__ bias, float *__restrict__ norm2_out, unsigned int *max_norm2) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const float *x = src + row * dim;
    const bool v4 = (dim & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    const float n2 = jv_warp_reduce_pair<false>(x, x, dim, lane, v4);
    const float inv = NORMALISE ? (n2 > 0.f ? 1.0f / sqrtf(n2) : 0.f) : 1.0f;
    for (int i = lane; i < Dp; i += 32) dst[row * Dp + i] = __float2bfloat16_rn(i < dim ? x[i] * inv : 0.f);
    if (lane == 0) {
        if (bias) bias[row] = -0.5f * n2;
        if (norm2_out) norm2_out[row] = n2;
        if (max_norm2) atomicMax(max_norm2, __float_as_uint(NORMALISE ? (n2 > 0.f ? 1.0002f : 0.f) : n2));
    }
}

// pass A -> thresholds: thr[q] = (k-th largest chunk maximum) - 2 eps(q); -inf when fewer than k chunks hold a live vector
__global__ void tc_threshold_kernel(const float *__restrict__ chunkmax, int nchunks, int k, const float *__restrict__ qnorm2,
                                    const unsigned int *__restrict__ max_norm2, float *__restrict__ thr, float *__restrict__ eps_out) {
    const int q = blockIdx.x, tid = threadIdx.x;
    const float *cm = chunkmax + (int64_t)q * nchunks;
    __shared__ int s_cnt;
    // bisection on the ordered-float image: the largest value v with count(cm >= v) >= k
    uint32_t lo = 0u, hi = 0xffffffffu;
    const uint32_t ninf = jv_f2ord(-INFINITY);
    for (int it = 0; it < 32; it++) {
        const uint32_t mid = lo + ((hi - lo) >> 1) + ((hi - lo) & 1u);
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        int c = 0;
        for (int i = tid; i < nchunks; i += blockDim.x) c += jv_f2ord(cm[i]) >= mid ? 1 : 0;
        if (c) atomicAdd(&s_cnt, c);
        __syncthreads();
        if (s_cnt >= k) lo = mid; else hi = mid - 1u;
        __syncthreads();
        if (lo == hi) break;
    }
    if (tid == 0) {
        const float eps = kTcEpsFactor * sqrtf(qnorm2[q]) * sqrtf(__uint_as_float(*max_norm2));
        eps_out[q] = eps;
        thr[q] = lo <= ninf ? -INFINITY : jv_ord2f(lo) - 2.0f * eps;
    }
}

// ||q||^2 per query (canonical) + bf16 copy padded to Dp; rows >= nq of the padded tile are zero
__global__ void tc_convert_queries_kernel(const float *__restrict__ q, int nq, int nqp, int dim, int Dp, __nv_bfloat16 *__restrict__ dst,
                                          float *__restrict__ qnorm2) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nqp) return;
    if (row >= nq) {
        for (int i = lane; i < Dp; i += 32) dst[(int64_t)row * Dp + i] = __float2bfloat16_rn(0.f);
        return;
    }
    const float *x = q + (int64_t)row * dim;
    const bool v4 = (dim & 3) == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0;
    const float n2 = jv_warp_reduce_pair<false>(x, x, dim, lane, v4);
    for (int i = lane; i < Dp; i += 32) dst[(int64_t)row * Dp + i] = __float2bfloat16_rn(i < dim ? x[i] : 0.f);
    if (lane == 0) qnorm2[row] = n2;
}

// select: the candidates within 2 eps of the k-th largest approximate score are re-scored with the canonical fp32 reduction and
// ranked by (score, docId) — the arithmetic of exact_kernel (jv_rerank.cu), so ids and score bits match the fp32 path.
constexpr int kTcSelThreads = 256;
__global__ void __launch_bounds__(kTcSelThreads)
tc_select_kernel(const float *__restrict__ vectors, const float *__restrict__ vec_norm, const int32_t *__restrict__ ord_to_doc, int dim,
                 int sim, float mul, const float *__restrict__ queries, int k, const unsigned long long *__restrict__ cand,
                 const int *__restrict__ cand_count, const float *__restrict__ eps, int32_t *__restrict__ out_doc,
                 float *__restrict__ out_score, int32_t *__restrict__ out_count, int *__restrict__ overflow, int *__restrict__ ovf_q) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    float *sq = reinterpret_cast<float *>(sel_smem);
    uint64_t *keys = reinterpret_cast<uint64_t *>(sel_smem + ((((size_t)dim * 4) + 15) & ~(size_t)15)); // [kTcCap2]
    uint32_t *ords = reinterpret_cast<uint32_t *>(keys + kTcCap2);                                       // [kTcCap2]
    __shared__ int s_cnt, s_n2;
    __shared__ float s_qn;
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total = cand_count[q];
    if (total > kTcCap) {
        if (tid == 0) {
            atomicAdd(overflow, 1);
            ovf_q[q] = 1;
        }
        return;
    }
    const unsigned long long *cq = cand + (int64_t)q * kTcCap;
    for (int i = tid; i < dim; i += kTcSelThreads) sq[i] = queries[(int64_t)q * dim + i];
    if (tid == 0) s_n2 = 0;
    __syncthreads();
    if (warp == 0) {
        const float qn = jv_warp_reduce_pair<false>(sq, queries + (int64_t)q * dim, dim, lane, false);
        if (lane == 0) s_qn = qn;
    }
    // k-th largest approximate score (bisection on the ordered-float image)
    uint32_t lo = 0u, hi = 0xffffffffu;
    if (total >= k) {
        for (int it = 0; it < 32; it++) {
            const uint32_t mid = lo + ((hi - lo) >> 1) + ((hi - lo) & 1u);
            if (tid == 0) s_cnt = 0;
            __syncthreads();
            int c = 0;
            for (int i = tid; i < total; i += kTcSelThreads) c += (uint32_t)(cq[i] >> 32) >= mid ? 1 : 0;
            if (c) atomicAdd(&s_cnt, c);
            __syncthreads();
            if (s_cnt >= k) lo = mid; else hi = mid - 1u;
            __syncthreads();
            if (lo == hi) break;
        }
    }
    const float cut = total >= k ? jv_ord2f(lo) - 2.0f * eps[q] : -INFINITY;
    for (int i = tid; i < total; i += kTcSelThreads) {
        const unsigned long long c = cq[i];
        if (jv_ord2f((uint32_t)(c >> 32)) >= cut) {
            const int slot = atomicAdd(&s_n2, 1);
            if (slot < kTcCap2) ords[slot] = (uint32_t)c;
        }
    }
    __syncthreads();
    const int n2 = s_n2;
    if (n2 > kTcCap2) {
        if (tid == 0) {
            atomicAdd(overflow, 1);
            ovf_q[q] = 1;
        }
        return;
    }
    const bool vec4 = (dim & 3) == 0 && (reinterpret_cast<uintptr_t>(vectors) & 15) == 0;
    for (int j = warp; j < n2; j += kTcSelThreads / 32) {
        const int64_t o = ords[j];
        const float *x = vectors + o * dim;
        const float raw = sim == JV_SIM_EUCLIDEAN ? jv_warp_reduce_pair<true>(sq, x, dim, lane, vec4) : jv_warp_reduce_pair<false>(sq, x, dim, lane, vec4);
        const float s = jv_finish_score(sim, raw, s_qn, sim == JV_SIM_COSINE ? vec_norm[o] : 0.f) * mul;
        const int32_t doc = ord_to_doc ? ord_to_doc[o] : (int32_t)o;
        if (lane == 0) keys[j] = jv_mk_key(s, doc);
    }
    __syncthreads();
    for (int j = tid; j < n2; j += kTcSelThreads) { // rank = number of strictly better keys (keys are unique: doc ids differ)
        const uint64_t my = keys[j];
        int rank = 0;
        for (int t = 0; t < n2; t++) rank += keys[t] > my ? 1 : 0;
        if (rank < k) {
            out_doc[(int64_t)q * k + rank] = jv_key_id(my);
            out_score[(int64_t)q * k + rank] = jv_key_score(my);
        }
    }
    const int nout = n2 < k ? n2 : k;
    for (int j = nout + tid; j < k; j += kTcSelThreads) {
        out_doc[(int64_t)q * k + j] = -1;
        out_score[(int64_t)q * k + j] = 0.f;
    }
    if (tid == 0) out_count[q] = nout;
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*TcEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                               CUtensorMapFloatOOBfill);

static TcEncodeFn tc_encode_fn() {
    static TcEncodeFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<TcEncodeFn>(p);
    });
    return fn;
}

// [rows][Dp] bf16, row pitch `pitch_bytes`, box = 64 columns x box_rows rows, 128-byte swizzle, out-of-range rows read as zero
static int32_t tc_make_map(CUtensorMap *tm, const void *base, int64_t rows, int Dp, int64_t pitch_bytes, int box_rows) {
    TcEncodeFn fn = tc_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return JV_ERR_UNSUPPORTED;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)Dp, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)kTcBK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): rows %lld, Dp %d, pitch %lld", (int)r, (long long)rows, Dp, (long long)pitch_bytes);
        return JV_ERR_CUDA;
    }
    return JV_OK;
}

bool exact_tc_eligible(const jv_index *ix, int nq, int k, const uint64_t *d_accept) {
    const int mode = q8_knobs().exact_tc; // -1 by size, 0 off, 1 forced
    if (mode == 0 || d_accept != nullptr || ix->vectors_on_host || ix->vectors_dev == nullptr) return false;
    if (k > 256 || ix->dim > 4096 || ix->n >= (1ll << 31)) return false;
    if (mode == 1) return true;
    // the bf16 copy costs one pass over the vectors (once per index); the contraction must be worth it
    return (double)nq * (double)ix->n * (double)ix->dim >= 2e11 && ix->n >= 65536 && nq >= 64;
}

// bf16 copy of the vectors (COSINE: normalised), per-vector bias (EUCLIDEAN) and max ||x||^2: made once per index
static int32_t tc_prepare_index(jv_index *ix, cudaStream_t stream) {
    std::lock_guard<std::mutex> lk(ix->mu);
    if (ix->tc_ready) return JV_OK;
    const int Dp = (ix->dim + kTcBK - 1) / kTcBK * kTcBK;
    JV_TRY(ix->tc_base.alloc((size_t)ix->n * Dp * 2));
    if (ix->sim == JV_SIM_EUCLIDEAN) JV_TRY(ix->tc_bias.alloc((size_t)ix->n * 4));
    JV_TRY(ix->tc_maxnorm.alloc(4));
    JV_CUDA_TRY(cudaMemsetAsync(ix->tc_maxnorm.p, 0, 4, stream));
    const int wpb = 8;
    const int64_t blocks = (ix->n + wpb - 1) / wpb;
    if (ix->sim == JV_SIM_COSINE)
        tc_convert_rows_kernel<true><<<(unsigned)blocks, wpb * 32, 0, stream>>>(ix->vectors_dev, ix->n, ix->dim, Dp, ix->tc_base.as<__nv_bfloat16>(), nullptr,
                                                                                nullptr, ix->tc_maxnorm.as<unsigned int>());
    else
        tc_convert_rows_kernel<false><<<(unsigned)blocks, wpb * 32, 0, stream>>>(ix->vectors_dev, ix->n, ix->dim, Dp, ix->tc_base.as<__nv_bfloat16>(),
                                                                                 ix->tc_bias.as<float>(), nullptr, ix->tc_maxnorm.as<unsigned int>());
    JV_CUDA_TRY(cudaGetLastError());
    JV_CUDA_TRY(cudaStreamSynchronize(stream));
    ix->tc_dp = Dp;
    ix->device_bytes += (int64_t)ix->tc_base.bytes + (int64_t)ix->tc_bias.bytes;
    ix->tc_ready = true;
    return JV_OK;
}

template <int MODE>
static int32_t tc_launch_gemm(jv_index *ix, cudaStream_t stream, const CUtensorMap &tmA, const CUtensorMap &tmB, TcParams &p, int nqp) {
    const int qtiles = nqp / kTcBM;
    // slices: a whole number of 256-column tiles each; between 2 and 12 waves of CTAs, the count that fills its last wave best
    const int64_t tiles = (p.n_cols + kTcBN - 1) / kTcBN;
    const int sms = ix->sm_count;
    int64_t best_S = 1;
    double best_fill = -1.0;
    for (int64_t S = 1; S <= tiles && S <= 65535 && S * qtiles <= (int64_t)12 * sms; S++) {
        const int64_t tp = (tiles + S - 1) / S, Se = (tiles + tp - 1) / tp, ctas = Se * qtiles;
        if (ctas < 2 * sms && S < tiles) continue;
        const double fill = (double)ctas / (double)(((ctas + sms - 1) / sms) * sms);
        if (fill > best_fill + 1e-9) best_fill = fill, best_S = Se;
    }
    const int64_t tiles_per = (tiles + best_S - 1) / best_S;
    p.cols_per_cta = tiles_per * kTcBN;
    const int64_t S = (tiles + tiles_per - 1) / tiles_per;
    const size_t smem = (size_t)kTcStages * kTcStageBytes + 256 + 1024;
    const dim3 grid((unsigned)qtiles, (unsigned)S);
    if (p.bias) {
        JV_CUDA_TRY(cudaFuncSetAttribute(exact_tc_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        exact_tc_kernel<MODE, true><<<grid, kTcThreads, smem, stream>>>(tmA, tmB, p);
    } else {
        JV_CUDA_TRY(cudaFuncSetAttribute(exact_tc_kernel<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        exact_tc_kernel<MODE, false><<<grid, kTcThreads, smem, stream>>>(tmA, tmB, p);
    }
    JV_CUDA_TRY(cudaGetLastError());
    return JV_OK;
}

// returns JV_OK with *done = false when many candidate lists overflowed (the caller answers the whole batch with the fp32 kernel);
// a few overflowing queries are re-answered by the fp32 kernel here
int32_t launch_exact_topk_tc(jv_index *ix, SearchCtx *ctx, const float *d_queries, int nq, int k, int32_t *d_out_doc, float *d_out_score,
                             int32_t *d_out_count, int *launches, bool *done) {
    *done = false;
    cudaStream_t st = ctx->stream;
    JV_TRY(tc_prepare_index(ix, st));
    const int Dp = ix->tc_dp, KB = Dp / kTcBK;
    const float mul = ix->sim == JV_SIM_MIP ? 2.0f : 1.0f;
    const int kMaxBatch = 16384; // queries per pass (bounds the candidate buffer: 512 MB)
    // sample for pass A: every stride-th vector.  The threshold then sits near rank k * stride of the full set — it has to stay
    // INSIDE the query's neighbourhood: measured at 1M x 768, k = 10, stride 150 puts it among the thousands of far vectors whose
    // scores lie within 2 eps of each other and every list overflows, stride 16 leaves a few hundred candidates.  Pass A costs
    // 1 / stride of pass B.  At least 4 k chunks of 32 columns.
    int64_t stride = 160 / k;
    if (stride < 1) stride = 1;
    while (stride > 1 && (ix->n / stride) < (int64_t)128 * k) stride--;
    const int64_t n_sample = (ix->n + stride - 1) / stride;
    const int nchunks = (int)((n_sample + 31) / 32);
    for (int q0 = 0; q0 < nq; q0 += kMaxBatch) {
        const int nqb = nq - q0 < kMaxBatch ? nq - q0 : kMaxBatch;
        const int nqp = (nqb + kTcBM - 1) / kTcBM * kTcBM;
        const float *dq = d_queries + (int64_t)q0 * ix->dim;
        JV_TRY(ctx->tc_q.ensure((size_t)nqp * Dp * 2));
        JV_TRY(ctx->tc_f.ensure((size_t)nqb * 16 + 16)); // qnorm2 | thr | eps | per-query overflow flags | overflow count
        JV_TRY(ctx->tc_chunk.ensure((size_t)nqb * nchunks * 4));
        JV_TRY(ctx->tc_cand.ensure((size_t)nqb * kTcCap * 8));
        JV_TRY(ctx->tc_cnt.ensure((size_t)nqb * 4));
        float *qn2 = ctx->tc_f.as<float>(), *thr = qn2 + nqb, *eps = thr + nqb;
        int *ovf_q = reinterpret_cast<int *>(eps + nqb), *ovf = ovf_q + nqb;
        tc_convert_queries_kernel<<<(nqp + 7) / 8, 256, 0, st>>>(dq, nqb, nqp, ix->dim, Dp, ctx->tc_q.as<__nv_bfloat16>(), qn2);
        JV_CUDA_TRY(cudaGetLastError());
        JV_CUDA_TRY(cudaMemsetAsync(ctx->tc_cnt.p, 0, (size_t)nqb * 4, st));
        JV_CUDA_TRY(cudaMemsetAsync(ovf_q, 0, (size_t)nqb * 4 + 4, st));
        CUtensorMap tmA, tmS, tmB;
        JV_TRY(tc_make_map(&tmA, ctx->tc_q.p, nqp, Dp, (int64_t)Dp * 2, kTcBM));
        JV_TRY(tc_make_map(&tmS, ix->tc_base.p, n_sample, Dp, (int64_t)Dp * 2 * stride, kTcBN));
        JV_TRY(tc_make_map(&tmB, ix->tc_base.p, ix->n, Dp, (int64_t)Dp * 2, kTcBN));
        TcParams p;
        memset(&p, 0, sizeof(p));
        p.bias = ix->sim == JV_SIM_EUCLIDEAN ? ix->tc_bias.as<float>() : nullptr;
        p.ord_to_doc = ix->ord_to_doc.as<int32_t>();
        p.thr = thr;
        p.chunkmax = ctx->tc_chunk.as<float>();
        p.cand = ctx->tc_cand.as<unsigned long long>();
        p.cand_count = ctx->tc_cnt.as<int>();
        p.nq = nqb;
        p.kblocks = KB;
        p.nchunks = nchunks;
        // pass A: strided sample -> per-chunk maxima -> thresholds
        p.n_cols = n_sample;
        p.col_stride = stride;
        JV_TRY(tc_launch_gemm<0>(ix, st, tmA, tmS, p, nqp));
        tc_threshold_kernel<<<nqb, 128, 0, st>>>(p.chunkmax, nchunks, k, qn2, ix->tc_maxnorm.as<unsigned int>(), thr, eps);
        JV_CUDA_TRY(cudaGetLastError());
        // pass B: all vectors -> candidates
        p.n_cols = ix->n;
        p.col_stride = 1;
        JV_TRY(tc_launch_gemm<1>(ix, st, tmA, tmB, p, nqp));
        const size_t smem = ((((size_t)ix->dim * 4) + 15) & ~(size_t)15) + (size_t)kTcCap2 * 12;
        JV_CUDA_TRY(cudaFuncSetAttribute(tc_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_select_kernel<<<nqb, kTcSelThreads, smem, st>>>(ix->vectors_dev, ix->vec_norm.as<float>(), ix->ord_to_doc.as<int32_t>(), ix->dim, ix->sim, mul, dq, k,
                                                           p.cand, p.cand_count, eps, d_out_doc + (int64_t)q0 * k, d_out_score + (int64_t)q0 * k,
                                                           d_out_count + q0, ovf, ovf_q);
        JV_CUDA_TRY(cudaGetLastError());
        int h_ovf = 0;
        JV_CUDA_TRY(cudaMemcpyAsync(&h_ovf, ovf, 4, cudaMemcpyDeviceToHost, st));
        JV_CUDA_TRY(cudaStreamSynchronize(st));
        if (launches) *launches += 5;
        if (h_ovf) {
            // queries whose candidate list overflowed (thousands of vectors within 2 eps of the k-th best: adversarial or degenerate
            // data) are answered by the fp32 kernel, one by one when they are few, the whole batch otherwise
            ix->tc_fallbacks += h_ovf;
            if (h_ovf > nqb / 8 + 16) return JV_OK; // *done stays false
            std::vector<int> flags((size_t)nqb);
            JV_CUDA_TRY(cudaMemcpyAsync(flags.data(), ovf_q, (size_t)nqb * 4, cudaMemcpyDeviceToHost, st));
            JV_CUDA_TRY(cudaStreamSynchronize(st));
            std::vector<int> redo;
            for (int i = 0; i < nqb; i++)
                if (flags[i]) redo.push_back(i);
            const size_t r = redo.size();
            JV_TRY(ctx->tc_redo.ensure(r * ((size_t)ix->dim * 4 + (size_t)k * 8 + 4)));
            float *rq = ctx->tc_redo.as<float>();
            int32_t *rd = reinterpret_cast<int32_t *>(rq + r * ix->dim);
            float *rs = reinterpret_cast<float *>(rd + r * k);
            int32_t *rc = reinterpret_cast<int32_t *>(rs + r * k);
            for (size_t j = 0; j < r; j++)
                JV_CUDA_TRY(cudaMemcpyAsync(rq + j * ix->dim, dq + (int64_t)redo[j] * ix->dim, (size_t)ix->dim * 4, cudaMemcpyDeviceToDevice, st));
            JV_TRY(launch_exact_topk_fp32(ix, ctx, rq, (int)r, k, nullptr, 0, rd, rs, rc, launches));
            for (size_t j = 0; j < r; j++) {
                const int64_t q = (int64_t)q0 + redo[j];
                JV_CUDA_TRY(cudaMemcpyAsync(d_out_doc + q * k, rd + j * k, (size_t)k * 4, cudaMemcpyDeviceToDevice, st));
                JV_CUDA_TRY(cudaMemcpyAsync(d_out_score + q * k, rs + j * k, (size_t)k * 4, cudaMemcpyDeviceToDevice, st));
                JV_CUDA_TRY(cudaMemcpyAsync(d_out_count + q, rc + j, 4, cudaMemcpyDeviceToDevice, st));
            }
        }
    }
    ix->tc_batches++;
    *done = true;
    return JV_OK;
}

}  // namespace jv
