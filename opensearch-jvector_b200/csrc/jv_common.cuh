// jv_common.cuh — shared host/device helpers of libjvgpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/jvgpu.h"

// --------------------------------------------------------------------------------------------
// error plumbing: thread-local message, int32 status, never throw across the C boundary
// --------------------------------------------------------------------------------------------
namespace jv {

void set_error(const char *fmt, ...);
const char *get_error();

#define JV_CUDA_TRY(expr)                                                                          \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            jv::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (_e == cudaErrorMemoryAllocation) ? JV_ERR_OUT_OF_MEMORY : JV_ERR_CUDA;         \
        }                                                                                          \
    } while (0)

#define JV_REQUIRE(cond, ...)              \
    do {                                   \
        if (!(cond)) {                     \
            jv::set_error(__VA_ARGS__);    \
            return JV_ERR_INVALID_ARGUMENT; \
        }                                  \
    } while (0)

#define JV_TRY(expr)              \
    do {                          \
        int32_t _s = (expr);      \
        if (_s != JV_OK) return _s; \
    } while (0)

// RAII device buffer (host-side bookkeeping only)
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    int32_t alloc(size_t n) {
        release();
        if (n == 0) return JV_OK;
        JV_CUDA_TRY(cudaMalloc(&p, n));
        bytes = n;
        return JV_OK;
    }
    int32_t ensure(size_t n) { return n <= bytes ? JV_OK : alloc(n); }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// sub-vector split of jVector ProductQuantization: base = dim/M, first dim%M get base+1
struct PqShape {
    int dim = 0, M = 0, K = 0;
    std::vector<int32_t> size, off;
    std::vector<int64_t> cb_off;
    int64_t cb_floats = 0;
    int max_size = 0;
    bool uniform = false;
    void init(int dim_, int M_, int K_) {
        dim = dim_, M = M_, K = K_;
        size.resize(M), off.resize(M), cb_off.resize(M);
        int base = dim / M, rem = dim % M, o = 0;
        int64_t co = 0;
        max_size = 0;
        for (int m = 0; m < M; m++) {
            size[m] = base + (m < rem ? 1 : 0);
            off[m] = o;
            cb_off[m] = co;
            o += size[m];
            co += (int64_t)K * size[m];
            if (size[m] > max_size) max_size = size[m];
        }
        cb_floats = co;
        uniform = rem == 0;
    }
};

}  // namespace jv

// --------------------------------------------------------------------------------------------
// device helpers
// --------------------------------------------------------------------------------------------
#define JV_FULL_MASK 0xffffffffu

// (score, id) -> uint64 key; larger key = better: higher score, then LOWER id
// (jVector NodeQueue / Lucene TopKnnCollector tie rule).
__host__ __device__ __forceinline__ uint32_t jv_f2ord(float f) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    uint32_t b;
    memcpy(&b, &f, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float jv_ord2f(uint32_t u) {
    uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}
__host__ __device__ __forceinline__ uint64_t jv_mk_key(float score, int32_t id) {
    return ((uint64_t)jv_f2ord(score) << 32) | (uint32_t)(~id);
}
__host__ __device__ __forceinline__ int32_t jv_key_id(uint64_t k) { return (int32_t)(~(uint32_t)k); }
__host__ __device__ __forceinline__ float jv_key_score(uint64_t k) { return jv_ord2f((uint32_t)(k >> 32)); }

#ifdef __CUDACC__

__device__ __forceinline__ float jv_warp_sum_canonical(float v) {
    // halving tree over the 32 lane partials: v[j] += v[j^off], off = 16..1 (same tree as the oracle's reduce128)
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = v + __shfl_xor_sync(JV_FULL_MASK, v, off);
    return v;
}

// Canonical fp32 reductions, one warp per (a, b) pair: element i goes to partial (i mod 128) with fmaf,
// lane j owns partials 4j..4j+3; bit-identical to oracle/jv_oracle.c canon_dot / canon_l2sq.
// `a` may live in shared memory, `b` in global; `vec4` = both are 16-byte aligned and dim % 4 == 0.
// BG = `b` lives in global memory (read through the read-only path); BG = false lets `b` be shared memory too
// (e.g. the norm of a query that was staged from pinned host memory: no second trip over PCIe).
template <bool L2, bool BG = true>
__device__ __forceinline__ float jv_warp_reduce_pair(const float *__restrict__ a, const float *__restrict__ b, int dim,
                                                     int lane, bool vec4) {
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
    if (vec4) {
        const float4 *a4 = reinterpret_cast<const float4 *>(a);
        const float4 *b4 = reinterpret_cast<const float4 *>(b);
        const int n4 = dim >> 2;
#pragma unroll 4
        for (int i = lane; i < n4; i += 32) {
            float4 x = a4[i];
            float4 y = BG ? __ldg(b4 + i) : b4[i];
            if (L2) {
                float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
                p0 = __fmaf_rn(d0, d0, p0);
                p1 = __fmaf_rn(d1, d1, p1);
                p2 = __fmaf_rn(d2, d2, p2);
                p3 = __fmaf_rn(d3, d3, p3);
            } else {
                p0 = __fmaf_rn(x.x, y.x, p0);
                p1 = __fmaf_rn(x.y, y.y, p1);
                p2 = __fmaf_rn(x.z, y.z, p2);
                p3 = __fmaf_rn(x.w, y.w, p3);
            }
        }
    } else {
        for (int base = 4 * lane; base < dim; base += 128) {
            float x[4], y[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const bool in = base + c < dim;
                x[c] = in ? a[base + c] : 0.f;
                y[c] = in ? (BG ? __ldg(b + base + c) : b[base + c]) : 0.f; // (0,0) contributes exactly +0 to dot and to L2
            }
            if (L2) {
                float d0 = x[0] - y[0], d1 = x[1] - y[1], d2 = x[2] - y[2], d3 = x[3] - y[3];
                p0 = __fmaf_rn(d0, d0, p0);
                p1 = __fmaf_rn(d1, d1, p1);
                p2 = __fmaf_rn(d2, d2, p2);
                p3 = __fmaf_rn(d3, d3, p3);
            } else {
                p0 = __fmaf_rn(x[0], y[0], p0);
                p1 = __fmaf_rn(x[1], y[1], p1);
                p2 = __fmaf_rn(x[2], y[2], p2);
                p3 = __fmaf_rn(x[3], y[3], p3);
            }
        }
    }
    float v = __fadd_rn(__fadd_rn(p0, p1), __fadd_rn(p2, p3));
    return jv_warp_sum_canonical(v);
}

// jVector VectorSimilarityFunction.compare.  `raw` = canonical dot or squared L2;
// cosine needs ||q||^2 and ||x||^2 (both canonical dots, precomputed).
__device__ __forceinline__ float jv_finish_score(int sim, float raw, float qnorm, float xnorm) {
    if (sim == JV_SIM_EUCLIDEAN) return __fdiv_rn(1.0f, __fadd_rn(1.0f, raw));
    if (sim == JV_SIM_COSINE) {
        float c = (float)((double)raw / sqrt((double)qnorm * (double)xnorm));
        return __fdiv_rn(__fadd_rn(1.0f, c), 2.0f);
    }
    return __fdiv_rn(__fadd_rn(1.0f, raw), 2.0f);
}

__device__ __forceinline__ bool jv_doc_accepted(const uint64_t *__restrict__ bits, int32_t doc) {
    if (doc < 0) return false;
    if (bits == nullptr) return true;
    return (bits[doc >> 6] >> (doc & 63)) & 1ull;
}

#endif  // __CUDACC__
