// jv_pq.cu — K6 PQ encode (+ index-creation helpers that share the codebook code).
//
// Replaces PQVectors.encodeAndBuild(pq, n, ravv, pool) at JVectorIndexQuantization.java:133 (flush) and
// JVectorWriter.java:1124 (merge re-encode with the leading segment's codebook):
//   x' = x - globalCentroid (if any);  code[m] = first argmin_c ||x'_m - C_m[c]||^2, strict '<' scan c = 0..K-1.
// The squared distance is accumulated with fmaf in sub-vector order, exactly as oracle sub_l2sq, so codes are
// bit-identical to the CPU restatement (no "documented fp ties" needed against it).
//
// Work split: thread = one vector, CTA = 256 vectors x MG consecutive subspaces whose codebooks
// (MG * K * sub floats) sit in shared memory and are read as warp-wide broadcasts.  FMA-bound:
// 2*dim*K flop per vector against dim*4 + M bytes (~128 flop/B).
#include "jv_internal.h"

#include <cuda_fp16.h>

namespace jv {

constexpr int kEncThreads = 256;

template <int SUB> struct SubVec;

template <int SUB>
__device__ __forceinline__ float sub_l2sq_reg(const float (&x)[SUB], const float *__restrict__ c) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < SUB; j++) {
        const float d = x[j] - c[j];
        acc = __fmaf_rn(d, d, acc);
    }
    return acc;
}

// uniform sub-vector size SUB in {1,2,3,4,6,8,12,16}: x_m kept in registers
template <int SUB>
__global__ void __launch_bounds__(kEncThreads)
encode_kernel_uniform(const float *__restrict__ x, int64_t n, int dim, int M, int K, const float *__restrict__ cb,
                      const float *__restrict__ g, uint8_t *__restrict__ out, int out_stride, int MG) {
    extern __shared__ __align__(16) float s_cb[]; // [MG][K][SUB]
    const int m0 = blockIdx.y * MG;
    const int mg = min(MG, M - m0);
    const int64_t cb_base = (int64_t)m0 * K * SUB;
    for (int i = threadIdx.x; i < mg * K * SUB; i += kEncThreads) s_cb[i] = __ldg(cb + cb_base + i);
    __syncthreads();
    const int64_t v = (int64_t)blockIdx.x * kEncThreads + threadIdx.x;
    if (v >= n) return;
    const float *xv = x + v * dim;
    uint8_t *ov = out + v * out_stride;
    uint32_t packed = 0;
    for (int ml = 0; ml < mg; ml++) {
        const int m = m0 + ml;
        float xs[SUB];
#pragma unroll
        for (int j = 0; j < SUB; j++) {
            float t = __ldg(xv + m * SUB + j);
            if (g) t = t - __ldg(g + m * SUB + j);
            xs[j] = t;
        }
        const float *cm = s_cb + (size_t)ml * K * SUB;
        float best = INFINITY;
        int idx = 0;
#pragma unroll 8
        for (int c = 0; c < K; c++) {
            const float d2 = sub_l2sq_reg<SUB>(xs, cm + c * SUB);
            if (d2 < best) {
                best = d2;
                idx = c;
            }
        }
        // pack 4 codes per 32-bit store when the group is word aligned
        if ((m0 & 3) == 0 && (out_stride & 3) == 0) {
            packed |= (uint32_t)idx << (8 * (ml & 3));
            if ((ml & 3) == 3 || ml == mg - 1) {
                if ((ml & 3) == 3) {
                    *reinterpret_cast<uint32_t *>(ov + (m & ~3)) = packed;
                } else {
                    for (int b = 0; b <= (ml & 3); b++) ov[(m & ~3) + b] = (uint8_t)(packed >> (8 * b));
                }
                packed = 0;
            }
        } else {
            ov[m] = (uint8_t)idx;
        }
    }
}

// generic: ragged sub-vector sizes (dim % M != 0) or sizes without a specialisation
__global__ void __launch_bounds__(kEncThreads)
encode_kernel_generic(const float *__restrict__ x, int64_t n, int dim, int M, int K, const float *__restrict__ cb,
                      const float *__restrict__ g, const int32_t *__restrict__ size, const int32_t *__restrict__ off,
                      const int32_t *__restrict__ cboff, uint8_t *__restrict__ out, int out_stride, int MG, int max_size) {
    extern __shared__ __align__(16) float s_cb[]; // [MG][K*max_size]
    const int m0 = blockIdx.y * MG;
    const int mg = min(MG, M - m0);
    for (int ml = 0; ml < mg; ml++) {
        const int len = size[m0 + ml];
        const float *src = cb + cboff[m0 + ml];
        for (int i = threadIdx.x; i < K * len; i += kEncThreads) s_cb[(size_t)ml * K * max_size + i] = __ldg(src + i);
    }
    __syncthreads();
    const int64_t v = (int64_t)blockIdx.x * kEncThreads + threadIdx.x;
    if (v >= n) return;
    const float *xv = x + v * dim;
    for (int ml = 0; ml < mg; ml++) {
        const int m = m0 + ml, len = size[m], o = off[m];
        const float *cm = s_cb + (size_t)ml * K * max_size;
        float best = INFINITY;
        int idx = 0;
        for (int c = 0; c < K; c++) {
            float acc = 0.f;
            for (int j = 0; j < len; j++) {
                float t = __ldg(xv + o + j);
                if (g) t = t - __ldg(g + o + j);
                const float d = t - cm[c * len + j];
                acc = __fmaf_rn(d, d, acc);
            }
            if (acc < best) {
                best = acc;
                idx = c;
            }
        }
        out[v * out_stride + m] = (uint8_t)idx;
    }
}

// PQ decode (ProductQuantization.decode): out[row][off_m + j] = C_m[code[row][m]][j] (+ global centroid, one fp32 add per element like
// the CPU restatement).  One thread per output float; codebooks laid out [m][c][size_m] with jVector's sub-vector split.
__global__ void pq_decode_kernel(const uint8_t *__restrict__ codes, int64_t n, int dim, int M, int K, int base, int rem,
                                 const float *__restrict__ codebooks, const float *__restrict__ gcent, float *__restrict__ out) {
    const int64_t total = n * dim;
    const int wide = rem * (base + 1); // the first `rem` subspaces are one element wider
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / dim;
        const int d = (int)(i - row * dim);
        int m, j, size, off;
        if (d < wide) {
            m = d / (base + 1), j = d - m * (base + 1), size = base + 1, off = m * (base + 1);
        } else {
            const int r = d - wide;
            m = rem + r / base, j = r - (r / base) * base, size = base, off = wide + (m - rem) * base;
        }
        const int c = codes[row * M + m];
        float v = __ldg(codebooks + (int64_t)K * off + (int64_t)c * size + j);
        if (gcent) v = __fadd_rn(v, __ldg(gcent + d));
        out[i] = v;
    }
}

int32_t launch_pq_decode(cudaStream_t stream, const PqShape &s, const uint8_t *d_codes, int64_t n, const float *d_codebooks,
                         const float *d_gcent, float *d_out) {
    if (n <= 0) return JV_OK;
    const int64_t total = n * s.dim;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    pq_decode_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_codes, n, s.dim, s.M, s.K, s.dim / s.M, s.dim % s.M, d_codebooks, d_gcent, d_out);
    JV_CUDA_TRY(cudaGetLastError());
    return JV_OK;
}

int32_t launch_pq_encode(cudaStream_t stream, const PqShape &s, const float *d_vectors, int64_t n, const float *d_codebooks,
                         const float *d_gcent, uint8_t *d_out, int out_stride) {
    if (n <= 0) return JV_OK;
    JV_REQUIRE(s.K >= 1 && s.K <= 256, "pq: K must be in [1,256]");
    const int64_t gx = (n + kEncThreads - 1) / kEncThreads;
    JV_REQUIRE(gx <= 0x7fffffff, "pq_encode: too many vectors for one launch");
    const int sub = s.uniform ? s.dim / s.M : 0;
    // subspaces per CTA: keep the staged codebooks <= 32 KB so several CTAs share an SM
    auto pick_mg = [&](int floats_per_sub) {
        int mg = (32 * 1024) / (floats_per_sub * 4);
        if (mg < 1) mg = 1;
        if (mg > s.M) mg = s.M;
        if (mg > 4) mg &= ~3; // keep groups word aligned for packed stores
        return mg;
    };
#define JV_ENC(SUB)                                                                                                        \
    {                                                                                                                      \
        const int MG = pick_mg(s.K * SUB);                                                                                 \
        const size_t smem = (size_t)MG * s.K * SUB * 4;                                                                    \
        JV_CUDA_TRY(cudaFuncSetAttribute(encode_kernel_uniform<SUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        dim3 grid((unsigned)gx, (unsigned)((s.M + MG - 1) / MG));                                                          \
        encode_kernel_uniform<SUB><<<grid, kEncThreads, smem, stream>>>(d_vectors, n, s.dim, s.M, s.K, d_codebooks, d_gcent, \
                                                                         d_out, out_stride, MG);                           \
    }
    switch (sub) {
    case 1: JV_ENC(1) break;
    case 2: JV_ENC(2) break;
    case 3: JV_ENC(3) break;
    case 4: JV_ENC(4) break;
    case 6: JV_ENC(6) break;
    case 8: JV_ENC(8) break;
    case 12: JV_ENC(12) break;
    case 16: JV_ENC(16) break;
    default: {
        // ragged / unusual sizes: per-subspace tables on the device
        DevBuf dsize, doff, dcb;
        std::vector<int32_t> cbo(s.M);
        for (int m = 0; m < s.M; m++) cbo[m] = (int32_t)s.cb_off[m];
        JV_TRY(dsize.alloc(s.M * 4));
        JV_TRY(doff.alloc(s.M * 4));
        JV_TRY(dcb.alloc(s.M * 4));
        JV_CUDA_TRY(cudaMemcpyAsync(dsize.p, s.size.data(), s.M * 4, cudaMemcpyHostToDevice, stream));
        JV_CUDA_TRY(cudaMemcpyAsync(doff.p, s.off.data(), s.M * 4, cudaMemcpyHostToDevice, stream));
        JV_CUDA_TRY(cudaMemcpyAsync(dcb.p, cbo.data(), s.M * 4, cudaMemcpyHostToDevice, stream));
        const int MG = pick_mg(s.K * s.max_size);
        const size_t smem = (size_t)MG * s.K * s.max_size * 4;
        JV_REQUIRE(smem <= 200 * 1024, "pq_encode: sub-vector too large (%d floats)", s.max_size);
        JV_CUDA_TRY(cudaFuncSetAttribute(encode_kernel_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((unsigned)gx, (unsigned)((s.M + MG - 1) / MG));
        encode_kernel_generic<<<grid, kEncThreads, smem, stream>>>(d_vectors, n, s.dim, s.M, s.K, d_codebooks, d_gcent,
                                                                   dsize.as<int32_t>(), doff.as<int32_t>(), dcb.as<int32_t>(),
                                                                   d_out, out_stride, MG, s.max_size);
        JV_CUDA_TRY(cudaGetLastError());
        JV_CUDA_TRY(cudaStreamSynchronize(stream)); // the tables die with this scope
        return JV_OK;
    }
    }
#undef JV_ENC
    JV_CUDA_TRY(cudaGetLastError());
    return JV_OK;
}

// ------------------------------------------------------------------------------------------------
// index-creation helpers
// ------------------------------------------------------------------------------------------------
// canonical ||x||^2 per vector (cosine): one warp per vector
__global__ void vec_norm_kernel(const float *__restrict__ x, int64_t n, int dim, float *out) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool vec4 = (dim & 3) == 0;
    for (int64_t v = w; v < n; v += nw) {
        const float *xv = x + v * dim;
        float s = jv_warp_reduce_pair<false>(xv, xv, dim, lane, vec4);
        if (lane == 0) out[v] = s;
    }
}

int32_t launch_vec_norms(cudaStream_t stream, const float *d_vectors, int64_t n, int dim, float *d_out) {
    if (n <= 0) return JV_OK;
    int64_t blocks = (n * 32 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    vec_norm_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_vectors, n, dim, d_out);
    JV_CUDA_TRY(cudaGetLastError());
    return JV_OK;
}

// ||decode(code)||^2 = sum_m ||C_m[code_m]||^2, sequential in m like oracle pq_node_norm (cosine ADC)
__global__ void node_norm_kernel(const uint8_t *__restrict__ codes, int code_stride, int64_t n, int M, int K,
                                 const float *__restrict__ cb, const int32_t *__restrict__ size,
                                 const int32_t *__restrict__ cboff, float *out) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const uint8_t *row = codes + v * code_stride;
    float acc = 0.f;
    for (int m = 0; m < M; m++) {
        const int len = size[m];
        const float *cv = cb + cboff[m] + (int64_t)row[m] * len;
        float d = 0.f;
        for (int j = 0; j < len; j++) d = __fmaf_rn(cv[j], cv[j], d);
        acc = __fadd_rn(acc, d);
    }
    out[v] = acc;
}

int32_t launch_node_norms(cudaStream_t stream, const jv_index *ix, float *d_out) {
    if (ix->n <= 0) return JV_OK;
    const int64_t blocks = (ix->n + 255) / 256;
    node_norm_kernel<<<(unsigned)blocks, 256, 0, stream>>>(ix->codes.as<uint8_t>(), ix->code_stride, ix->n, ix->pq.M, ix->pq.K,
                                                          ix->codebooks.as<float>(), ix->pq_size.as<int32_t>(),
                                                          ix->pq_cboff.as<int32_t>(), d_out);
    JV_CUDA_TRY(cudaGetLastError());
    return JV_OK;
}

__global__ void f32_to_f16_kernel(const float *__restrict__ in, int64_t n, __half *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __float2half_rn(in[i]);
}

int32_t launch_f32_to_f16(cudaStream_t stream, const float *d_in, int64_t n, void *d_out) {
    if (n <= 0) return JV_OK;
    f32_to_f16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_in, n, reinterpret_cast<__half *>(d_out));
    JV_CUDA_TRY(cudaGetLastError());
    return JV_OK;
}

int32_t launch_build_fused(cudaStream_t, jv_index *) { return JV_OK; }

}  // namespace jv
