// jv_api.cu — the extern "C" boundary of libjvgpu.so (include/jvgpu.h): argument validation, index
// lifetime, per-call context pool, H2D/D2H staging, timing.  No torch, no CPU compute fallback.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include "jv_internal.h"
#include "jv_rerank_body.cuh"

namespace jv {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char *get_error() { return g_err; }

int32_t SearchCtx::init(int) {
    JV_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    JV_CUDA_TRY(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    for (auto &e : ev) JV_CUDA_TRY(cudaEventCreate(&e));
    for (auto &e : chunk_ev) JV_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return JV_OK;
}
int32_t SearchCtx::ensure_pinned(size_t bytes) {
    if (bytes <= pinned_bytes) return JV_OK;
    if (pinned) cudaFreeHost(pinned);
    pinned = nullptr;
    pinned_bytes = 0;
    JV_CUDA_TRY(cudaHostAlloc(&pinned, bytes, cudaHostAllocDefault));
    pinned_bytes = bytes;
    return JV_OK;
}
void SearchCtx::destroy() {
    if (stream) cudaStreamDestroy(stream);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    for (auto &e : ev)
        if (e) cudaEventDestroy(e);
    for (auto &e : chunk_ev)
        if (e) cudaEventDestroy(e);
    copy_stream = nullptr;
    if (pinned) cudaFreeHost(pinned);
    stream = nullptr;
    pinned = nullptr;
}

}  // namespace jv

jv::SearchCtx *jv_index::acquire() {
    {
        std::lock_guard<std::mutex> lk(mu);
        if (!pool.empty()) {
            jv::SearchCtx *c = pool.back();
            pool.pop_back();
            return c;
        }
    }
    auto *c = new jv::SearchCtx();
    if (c->init(device) != JV_OK) {
        c->destroy();
        delete c;
        return nullptr;
    }
    return c;
}
void jv_index::release(jv::SearchCtx *c) {
    std::lock_guard<std::mutex> lk(mu);
    pool.push_back(c);
}

using namespace jv;

namespace {

struct CtxLease {
    jv_index *ix;
    SearchCtx *c;
    explicit CtxLease(jv_index *i) : ix(i), c(i->acquire()) {}
    ~CtxLease() {
        if (c) ix->release(c);
    }
};

int32_t check_device(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        set_error("no usable CUDA device (%s); libjvgpu has no CPU fallback", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
        return JV_ERR_CUDA;
    }
    if (device < 0 || device >= count) {
        set_error("device %d out of range (0..%d)", device, count - 1);
        return JV_ERR_INVALID_ARGUMENT;
    }
    return JV_OK;
}

int32_t upload(DevBuf &b, const void *src, size_t bytes, int64_t *total) {
    JV_TRY(b.alloc(bytes));
    if (bytes) JV_CUDA_TRY(cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice));
    *total += (int64_t)bytes;
    return JV_OK;
}

int32_t validate_params(const jv_index *ix, int32_t nq, const jv_search_params *p) {
    JV_REQUIRE(ix != nullptr, "index is NULL");
    JV_REQUIRE(p != nullptr, "params is NULL");
    JV_REQUIRE(p->struct_size == (int32_t)sizeof(jv_search_params), "jv_search_params.struct_size mismatch");
    JV_REQUIRE(nq >= 0, "nq must be >= 0");
    JV_REQUIRE(p->k >= 1, "k must be >= 1");
    JV_REQUIRE(p->rerank_k >= p->k, "rerankK must be >= topK (GraphSearcher contract)");
    JV_REQUIRE(p->rerank_k <= 4096, "rerank_k > 4096 not supported");
    JV_REQUIRE(p->accept_stride_words >= 0, "accept_stride_words must be >= 0");
    JV_REQUIRE(p->expand_width >= -1 && p->expand_width <= 8, "expand_width must be in [-1, 8]");
    return JV_OK;
}

}  // namespace

extern "C" {

int32_t jv_version(void) { return (JVGPU_VERSION_MAJOR << 16) | JVGPU_VERSION_MINOR; }

const char *jv_last_error(void) { return get_error(); }

int32_t jv_device_count(int32_t *out_count) {
    JV_REQUIRE(out_count != nullptr, "out_count is NULL");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        *out_count = 0;
        set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return JV_ERR_CUDA;
    }
    *out_count = c;
    return JV_OK;
}

// page-locked host buffers for the callers of the host-pointer entry points (the FFM shim, JVectorReader.java:147)
int32_t jv_host_alloc(int64_t bytes, void **out_ptr) {
    JV_REQUIRE(out_ptr != nullptr && bytes > 0, "bad arguments");
    *out_ptr = nullptr;
    JV_CUDA_TRY(cudaHostAlloc(out_ptr, (size_t)bytes, cudaHostAllocPortable));
    return JV_OK;
}

int32_t jv_host_free(void *ptr) {
    if (ptr == nullptr) return JV_OK;
    JV_CUDA_TRY(cudaFreeHost(ptr));
    return JV_OK;
}

int32_t jv_host_register(void *ptr, int64_t bytes) {
    JV_REQUIRE(ptr != nullptr && bytes > 0, "bad arguments");
    JV_CUDA_TRY(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
    return JV_OK;
}

int32_t jv_host_unregister(void *ptr) {
    JV_REQUIRE(ptr != nullptr, "bad arguments");
    JV_CUDA_TRY(cudaHostUnregister(ptr));
    return JV_OK;
}

// NVQ-only segments (no auxiliary PQ blob): the traversal is scored by the NVQ reranker itself (JVectorReader.java:357-358:
// DefaultSearchScoreProvider(view.rerankerFor(q, sim)), not MIP-wrapped), i.e. by exact scores against the DEQUANTISED vectors.  They are
// dequantised once, at index creation, with the operations of nvq_decode_warp (jv_rerank_body.cuh; JVectorIndexQuantization.java:316-361):
// one warp per vector.
__global__ void nvq_dequantize_all_kernel(const uint8_t *__restrict__ bytes, const float *__restrict__ params, const float *__restrict__ gmean,
                                          const int32_t *__restrict__ off, int m, int64_t n, int dim, float *__restrict__ out) {
    extern __shared__ float nvq_consts[]; // [warps][m][4]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t node = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (node >= n) return;
    float *c = nvq_consts + (size_t)warp * m * 4;
    if (lane < m) {
        const float *prm = params + (node * m + lane) * 4;
        const float growth = __ldg(prm), midpoint = __ldg(prm + 1), lo = __ldg(prm + 2), hi = __ldg(prm + 3);
        const float delta = __fsub_rn(hi, lo);
        const float sgr = __fdiv_rn(growth, delta);
        const float mid = __fmul_rn(midpoint, delta);
        const float bias = nvq_logistic(lo, sgr, mid);
        c[lane * 4 + 0] = __fdiv_rn(__fsub_rn(nvq_logistic(hi, sgr, mid), bias), 255.0f);
        c[lane * 4 + 1] = bias;
        c[lane * 4 + 2] = __fdiv_rn(1.0f, sgr);
        c[lane * 4 + 3] = mid;
    }
    __syncwarp();
    const uint8_t *b = bytes + node * dim;
    for (int i = lane; i < dim; i += 32) {
        int sub = 0;
        while (sub + 1 < m && i >= __ldg(off + sub + 1)) sub++;
        const float y = nvq_logit(__fmaf_rn((float)__ldg(b + i), c[sub * 4], c[sub * 4 + 1]), c[sub * 4 + 2], c[sub * 4 + 3]);
        out[node * dim + i] = __fadd_rn(y, __ldg(gmean + i));
    }
}

// jv_index_create validation: the decoded arrays come from the caller, and a neighbour id, doc id or code
// out of range would be an out-of-bounds device read later — which poisons the CUDA context of the whole JVM.  flags: 1 adjacency,
// 2 ord_to_doc, 4 PQ code.
__global__ void validate_index_kernel(const int32_t *__restrict__ adjacency, int64_t n, int R, const int32_t *__restrict__ ord_to_doc,
                                      int max_doc, const uint8_t *__restrict__ codes, int code_stride, int M, int K, int *flags) {
    int bad = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = t0; i < n * R; i += stride) {
        const int32_t nb = adjacency[i];
        if (nb < -1 || nb >= n) bad |= 1;
    }
    if (ord_to_doc)
        for (int64_t i = t0; i < n; i += stride) {
            const int32_t d = ord_to_doc[i];
            if (d < -1 || d >= max_doc) bad |= 2;
        }
    if (codes && K < 256)
        for (int64_t i = t0; i < n * M; i += stride) {
            const int64_t row = i / M;
            if (codes[row * code_stride + (i - row * M)] >= K) bad |= 4;
        }
    if (bad) atomicOr(flags, bad);
}

int32_t jv_index_create(const jv_index_desc *desc, jv_index **out) {
    JV_REQUIRE(desc != nullptr && out != nullptr, "desc/out is NULL");
    *out = nullptr;
    JV_REQUIRE(desc->struct_size == (int32_t)sizeof(jv_index_desc) || desc->struct_size == JV_INDEX_DESC_SIZE_V1,
               "jv_index_desc.struct_size mismatch (%d vs %zu or %d)", desc->struct_size, sizeof(jv_index_desc), JV_INDEX_DESC_SIZE_V1);
    jv_index_desc dcopy; // the V1 layout has no NVQ fields: zero-filled
    memset(&dcopy, 0, sizeof(dcopy));
    memcpy(&dcopy, desc, (size_t)desc->struct_size);
    const jv_index_desc *d = &dcopy;
    const bool has_nvq = d->nvq_m > 0;
    JV_REQUIRE(d->similarity >= JV_SIM_EUCLIDEAN && d->similarity <= JV_SIM_MIP, "unknown similarity ordinal %d", d->similarity);
    JV_REQUIRE(d->dim >= 1 && d->n >= 0 && d->n < 0x7fffffffLL, "bad dim/n");
    JV_REQUIRE(d->max_degree >= 1 && d->max_degree <= 128, "max_degree must be in [1,128]");
    JV_REQUIRE(d->n == 0 || d->adjacency, "adjacency is NULL");
    JV_REQUIRE(d->n == 0 || d->vectors || has_nvq, "vectors are NULL (only an NVQ-inline segment may omit them)");
    if (has_nvq) {
        JV_REQUIRE(d->nvq_bytes && d->nvq_params && d->nvq_global_mean, "nvq_bytes/nvq_params/nvq_global_mean are NULL");
        JV_REQUIRE(d->nvq_m <= 32 && d->nvq_m <= d->dim, "nvq_m must be in [1, 32]");
        if (d->pq_m <= 0 && (d->flags & JV_INDEX_FLAG_NO_VECTORS_ON_DEVICE)) {
            set_error("an NVQ-only segment is traversed over its dequantised vectors, which live on the device");
            return JV_ERR_UNSUPPORTED;
        }
    }
    JV_REQUIRE(d->n == 0 || (d->entry_node >= 0 && d->entry_node < d->n), "entry_node out of range");
    JV_REQUIRE(d->max_doc >= 0 && (d->ord_to_doc != nullptr || d->n == 0 || (int64_t)d->max_doc >= d->n),
               "max_doc %d is smaller than the %lld ordinals of an identity doc map", d->max_doc, (long long)d->n);
    const bool has_pq = d->pq_m > 0;
    if (has_pq) {
        JV_REQUIRE(d->pq_codes && d->pq_codebooks, "pq_codes/pq_codebooks are NULL");
        JV_REQUIRE(d->pq_k >= 1 && d->pq_k <= 256 && d->pq_m <= d->dim, "bad PQ shape M=%d K=%d", d->pq_m, d->pq_k);
        if (d->pq_global_centroid && d->similarity != JV_SIM_EUCLIDEAN) {
            set_error("a PQ global centroid is only defined for EUCLIDEAN (JVectorIndexQuantization.java:127)");
            return JV_ERR_UNSUPPORTED;
        }
    }
    JV_TRY(check_device(d->device));
    DeviceGuard guard(d->device);

    auto *ix = new jv_index();
    ix->device = d->device;
    ix->sim = d->similarity;
    ix->dim = d->dim;
    ix->R = d->max_degree;
    ix->entry = d->entry_node;
    ix->max_doc = d->max_doc;
    ix->n = d->n;
    ix->flags = d->flags;
    ix->has_pq = has_pq;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, d->device) == cudaSuccess) {
        ix->sm_count = prop.multiProcessorCount;
        ix->smem_optin = prop.sharedMemPerBlockOptin;
    }
    int64_t total = 0;
    int32_t st = JV_OK;
    auto fail = [&](int32_t s) {
        jv_index_destroy(ix);
        return s;
    };
    if ((st = ix->dbg.alloc(256)) != JV_OK) return fail(st);
    cudaMemset(ix->dbg.p, 0, 256);
    const size_t n = (size_t)d->n;
    if ((st = upload(ix->adjacency, d->adjacency, n * d->max_degree * 4, &total)) != JV_OK) return fail(st);
    if ((d->flags & JV_INDEX_FLAG_NO_VECTORS_ON_DEVICE) && d->vectors) {
        // cfg 5: fp32 rerank vectors stay in pinned, device-mapped host memory
        cudaError_t e = cudaHostAlloc(&ix->vectors_host, n * d->dim * 4 + 16, cudaHostAllocMapped | cudaHostAllocPortable);
        if (e != cudaSuccess) {
            set_error("cudaHostAlloc(vectors): %s", cudaGetErrorString(e));
            return fail(JV_ERR_OUT_OF_MEMORY);
        }
        memcpy(ix->vectors_host, d->vectors, n * d->dim * 4);
        void *dp = nullptr;
        e = cudaHostGetDevicePointer(&dp, ix->vectors_host, 0);
        if (e != cudaSuccess) {
            set_error("cudaHostGetDevicePointer: %s", cudaGetErrorString(e));
            return fail(JV_ERR_CUDA);
        }
        ix->vectors_dev = static_cast<float *>(dp);
        ix->vectors_on_host = true;
    } else if (d->vectors) {
        if ((st = upload(ix->vectors, d->vectors, n * d->dim * 4, &total)) != JV_OK) return fail(st);
        ix->vectors_dev = ix->vectors.as<float>();
    }
    if (has_nvq) {
        PqShape sh;
        sh.init(d->dim, d->nvq_m, 1);
        std::vector<int32_t> off(sh.off.begin(), sh.off.end());
        off.push_back(d->dim);
        if ((st = upload(ix->nvq_bytes, d->nvq_bytes, n * d->dim, &total)) != JV_OK) return fail(st);
        if ((st = upload(ix->nvq_params, d->nvq_params, n * d->nvq_m * 16, &total)) != JV_OK) return fail(st);
        if ((st = upload(ix->nvq_gmean, d->nvq_global_mean, (size_t)d->dim * 4, &total)) != JV_OK) return fail(st);
        if ((st = upload(ix->nvq_off, off.data(), off.size() * 4, &total)) != JV_OK) return fail(st);
        ix->has_nvq = true;
        ix->nvq_m = d->nvq_m;
        ix->fp32_given = d->vectors != nullptr;
        if (!has_pq && n > 0) {
            // NVQ-only: the dequantised vectors ARE the vectors of this index (they replace fp32 inline vectors if those were given too:
            // "NVQ wins", like the rerank of nvq+pq segments)
            ix->vectors.release();
            if ((st = ix->vectors.alloc(n * d->dim * 4)) != JV_OK) return fail(st);
            total += (int64_t)n * d->dim * 4;
            const int wpb = 8;
            nvq_dequantize_all_kernel<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, (size_t)wpb * d->nvq_m * 16>>>(
                ix->nvq_bytes.as<uint8_t>(), ix->nvq_params.as<float>(), ix->nvq_gmean.as<float>(), ix->nvq_off.as<int32_t>(), d->nvq_m, d->n, d->dim,
                ix->vectors.as<float>());
            if (cudaGetLastError() != cudaSuccess) {
                set_error("NVQ dequantisation failed: %s", cudaGetErrorString(cudaGetLastError()));
                return fail(JV_ERR_CUDA);
            }
            ix->vectors_dev = ix->vectors.as<float>();
            ix->vectors_on_host = false;
            ix->nvq_only = true;
        }
    }
    if (d->ord_to_doc)
        if ((st = upload(ix->ord_to_doc, d->ord_to_doc, n * 4, &total)) != JV_OK) return fail(st);
    if (d->similarity == JV_SIM_COSINE && n > 0 && ix->vectors_dev) {
        if ((st = ix->vec_norm.alloc(n * 4)) != JV_OK) return fail(st);
        total += (int64_t)n * 4;
        if ((st = launch_vec_norms(nullptr, ix->vectors_dev, d->n, d->dim, ix->vec_norm.as<float>())) != JV_OK) return fail(st);
    }
    if (has_pq) {
        ix->pq.init(d->dim, d->pq_m, d->pq_k);
        ix->code_stride = (d->pq_m + 15) & ~15;
        if ((st = upload(ix->codebooks, d->pq_codebooks, (size_t)ix->pq.cb_floats * 4, &total)) != JV_OK) return fail(st);
        if (d->flags & JV_INDEX_FLAG_LUT_F16) { // fp16 copy for the fast kernel's table build
            if ((st = ix->codebooks_h.alloc((size_t)ix->pq.cb_floats * 2)) != JV_OK) return fail(st);
            total += ix->pq.cb_floats * 2;
            if ((st = launch_f32_to_f16(nullptr, ix->codebooks.as<float>(), ix->pq.cb_floats, ix->codebooks_h.p)) != JV_OK) return fail(st);
        }
        if (d->pq_global_centroid)
            if ((st = upload(ix->gcent, d->pq_global_centroid, (size_t)d->dim * 4, &total)) != JV_OK) return fail(st);
        std::vector<int32_t> cbo(d->pq_m);
        for (int m = 0; m < d->pq_m; m++) cbo[m] = (int32_t)ix->pq.cb_off[m];
        if ((st = upload(ix->pq_size, ix->pq.size.data(), (size_t)d->pq_m * 4, &total)) != JV_OK) return fail(st);
        if ((st = upload(ix->pq_off, ix->pq.off.data(), (size_t)d->pq_m * 4, &total)) != JV_OK) return fail(st);
        if ((st = upload(ix->pq_cboff, cbo.data(), (size_t)d->pq_m * 4, &total)) != JV_OK) return fail(st);
        if ((st = ix->codes.alloc(n * ix->code_stride)) != JV_OK) return fail(st);
        total += (int64_t)n * ix->code_stride;
        if (n > 0) {
            if (cudaMemset(ix->codes.p, 0, n * ix->code_stride) != cudaSuccess ||
                cudaMemcpy2D(ix->codes.p, ix->code_stride, d->pq_codes, d->pq_m, d->pq_m, n, cudaMemcpyHostToDevice) != cudaSuccess) {
                set_error("uploading PQ codes failed: %s", cudaGetErrorString(cudaGetLastError()));
                return fail(JV_ERR_CUDA);
            }
        }
        if (d->similarity == JV_SIM_COSINE && n > 0) {
            if ((st = ix->node_norm.alloc(n * 4)) != JV_OK) return fail(st);
            total += (int64_t)n * 4;
            if ((st = launch_node_norms(nullptr, ix, ix->node_norm.as<float>())) != JV_OK) return fail(st);
        }
        const int S = ix->pq.uniform ? d->dim / d->pq_m : 0;
        if ((d->flags & JV_INDEX_FLAG_LUT_U8) && d->pq_k == 256 && (S == 2 || S == 4 || S == 8)) {
            // 8-bit table path: bounding ball of every subspace codebook (same double-precision definition as the
            // oracle's pq_ball) + lane-major permuted code rows
            ix->q8_nj = (d->pq_m + 31) / 32;
            std::vector<float> ctr((size_t)d->dim), rad((size_t)d->pq_m);
            for (int m = 0; m < d->pq_m; m++) {
                const float *cb = d->pq_codebooks + ix->pq.cb_off[m];
                for (int j = 0; j < S; j++) {
                    double acc = 0.0;
                    for (int c = 0; c < 256; c++) acc += (double)cb[(size_t)c * S + j];
                    ctr[(size_t)m * S + j] = (float)(acc / 256.0);
                }
                double r2max = 0.0;
                for (int c = 0; c < 256; c++) {
                    double r2 = 0.0;
                    for (int j = 0; j < S; j++) {
                        const double dd = (double)cb[(size_t)c * S + j] - (double)ctr[(size_t)m * S + j];
                        r2 += dd * dd;
                    }
                    if (r2 > r2max) r2max = r2;
                }
                rad[m] = (float)sqrt(r2max) * 1.0009765625f;
            }
            if ((st = upload(ix->ball_ctr, ctr.data(), ctr.size() * 4, &total)) != JV_OK) return fail(st);
            if ((st = upload(ix->ball_rad, rad.data(), rad.size() * 4, &total)) != JV_OK) return fail(st);
            const size_t qb = n * (size_t)ix->q8_nj * 32;
            if ((st = ix->codes_q8.alloc(qb)) != JV_OK) return fail(st);
            total += (int64_t)qb;
            if ((st = launch_permute_codes(nullptr, ix->codes.as<uint8_t>(), d->n, d->pq_m, ix->code_stride, ix->q8_nj,
                                           ix->codes_q8.as<uint8_t>())) != JV_OK)
                return fail(st);
            ix->q8_ok = true;
        }
    }
    if (n > 0) { // range checks on the device copies (dbg slot 3 is the scratch flag word; it is cleared again below)
        int *vflags = ix->dbg.as<int>() + 3;
        validate_index_kernel<<<ix->sm_count * 4, 256>>>(ix->adjacency.as<int32_t>(), d->n, d->max_degree, ix->ord_to_doc.as<int32_t>(), d->max_doc,
                                                         has_pq ? ix->codes.as<uint8_t>() : nullptr, ix->code_stride, d->pq_m, d->pq_k, vflags);
        int h = 0;
        if (cudaMemcpy(&h, vflags, 4, cudaMemcpyDeviceToHost) != cudaSuccess) {
            set_error("index validation failed: %s", cudaGetErrorString(cudaGetLastError()));
            return fail(JV_ERR_CUDA);
        }
        cudaMemset(vflags, 0, 4);
        if (h) {
            set_error("jv_index_desc holds out-of-range data:%s%s%s", (h & 1) ? " adjacency (neighbour ids must be -1 or in [0, n))" : "",
                      (h & 2) ? " ord_to_doc (doc ids must be -1 or in [0, max_doc))" : "", (h & 4) ? " pq_codes (codes must be < pq_k)" : "");
            return fail(JV_ERR_INVALID_ARGUMENT);
        }
    }
    if (cudaDeviceSynchronize() != cudaSuccess) {
        set_error("index creation kernels failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(JV_ERR_CUDA);
    }
    ix->device_bytes = total;
    *out = ix;
    return JV_OK;
}

int32_t jv_index_destroy(jv_index *ix) {
    if (!ix) return JV_OK;
    DeviceGuard guard(ix->device);
    for (auto *c : ix->pool) {
        c->destroy();
        delete c;
    }
    ix->pool.clear();
    if (ix->vectors_host) cudaFreeHost(ix->vectors_host);
    delete ix;
    return JV_OK;
}

int32_t jv_index_device_bytes(const jv_index *ix, int64_t *out_bytes) {
    JV_REQUIRE(ix && out_bytes, "NULL argument");
    *out_bytes = ix->device_bytes;
    return JV_OK;
}

int32_t jv_index_debug_counter(jv_index *ix, int32_t which, int64_t *out_value) {
    JV_REQUIRE(ix && out_value && ((which >= 0 && which < 6) || (which >= 8 && which < 24) || which == 100 || which == 200), "bad arguments");
    if (which == 4 || which == 5) { // brute-force batches answered by the tensor-core path / handed back to the fp32 kernel (overflow)
        *out_value = which == 4 ? ix->tc_batches : ix->tc_fallbacks;
        return JV_OK;
    }
    if (which == 200) { // re-read the diagnostic environment knobs (they are cached: never a getenv on the search path)
        q8_knobs_refresh();
        *out_value = 0;
        return JV_OK;
    }
    DeviceGuard guard(ix->device);
    if (which == 100) { // reset
        JV_CUDA_TRY(cudaMemset(ix->dbg.p, 0, 256));
        *out_value = 0;
        return JV_OK;
    }
    unsigned char raw[256];
    JV_CUDA_TRY(cudaMemcpy(raw, ix->dbg.p, 256, cudaMemcpyDeviceToHost));
    if (which < 4) {
        int32_t v;
        memcpy(&v, raw + which * 4, 4);
        *out_value = v;
    } else { // per-phase SM cycles of the fast traversal kernel (thread 0 of every CTA, summed)
        uint64_t v;
        memcpy(&v, raw + 64 + (which - 8) * 8, 8);
        *out_value = (int64_t)v;
    }
    return JV_OK;
}

// -------------------------------------------------------------------------------------------------
// search
// -------------------------------------------------------------------------------------------------
static int32_t search_core(jv_index *ix, SearchCtx *c, const float *d_queries, int32_t nq, const jv_search_params *p,
                           const uint64_t *d_accept, int32_t *d_out_doc, float *d_out_score, int32_t *d_out_count,
                           jv_query_stats *d_stats, int *launches, bool timed = true) {
    JV_TRY(c->approx_keys.ensure((size_t)nq * p->rerank_k * 8));
    JV_TRY(c->approx_count.ensure((size_t)nq * 4));
    SearchLaunch a;
    a.d_queries = d_queries;
    a.nq = nq;
    a.rerank_k = p->rerank_k;
    a.threshold = p->threshold;
    a.d_accept = d_accept;
    a.accept_stride_words = p->accept_stride_words;
    a.d_approx_keys = c->approx_keys.as<uint64_t>();
    a.d_approx_count = c->approx_count.as<int32_t>();
    a.d_stats = d_stats;
    a.entry_override = -1;
    a.n_limit = ix->n;
    a.expand_width = p->expand_width;
    c->time_lut = timed; // the table-build event (ev[5]) belongs to the chunk whose ev[1..3] are recorded
    if (timed) {
        c->lut_timed = false;
        JV_CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
    }
    a.fuse_k = p->k;
    a.rerank_floor = p->rerank_floor;
    a.d_out_doc = d_out_doc;
    a.d_out_score = d_out_score;
    a.d_out_count = d_out_count;
    bool reranked = false;
    JV_TRY(launch_search(ix, c, a, launches, &reranked));
    if (timed) JV_CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
    if (!reranked)
        JV_TRY(launch_rerank(ix, c, d_queries, nq, p->k, p->rerank_k, p->rerank_floor, a.d_approx_keys, a.d_approx_count, d_out_doc,
                             d_out_score, d_out_count, d_stats, launches));
    if (timed) JV_CUDA_TRY(cudaEventRecord(c->ev[3], c->stream));
    return JV_OK;
}

static void fill_timing(SearchCtx *c, jv_batch_timing *t, int launches, bool host) {
    if (!t) return;
    memset(t, 0, sizeof(*t));
    cudaEventElapsedTime(&t->search_ms, c->ev[1], c->ev[2]);
    cudaEventElapsedTime(&t->rerank_ms, c->ev[2], c->ev[3]);
    if (host) {
        cudaEventElapsedTime(&t->h2d_ms, c->ev[0], c->ev[1]);
        cudaEventElapsedTime(&t->d2h_ms, c->ev[3], c->ev[4]);
        cudaEventElapsedTime(&t->total_ms, c->ev[0], c->ev[4]);
    } else {
        cudaEventElapsedTime(&t->total_ms, c->ev[1], c->ev[3]);
    }
    t->launches = launches;
    if (c->lut_timed) cudaEventElapsedTime(&t->lut_ms, c->ev[1], c->ev[5]);
    t->expand_width_used = c->last_width;
    t->traversal_kernel = c->last_kernel;
}

int32_t jv_search_batch_dev(jv_index *ix, const float *d_queries, int32_t nq, const jv_search_params *p, int32_t *d_out_doc,
                            float *d_out_score, int32_t *d_out_count, jv_query_stats *d_stats, jv_batch_timing *timing) {
    JV_TRY(validate_params(ix, nq, p));
    if (nq == 0) return JV_OK;
    JV_REQUIRE(d_queries && d_out_doc && d_out_score && d_out_count, "NULL buffer");
    JV_REQUIRE(ix->n > 0, "index is empty");
    DeviceGuard guard(ix->device);
    CtxLease lease(ix);
    JV_REQUIRE(lease.c != nullptr, "could not create a search context: %s", get_error());
    SearchCtx *c = lease.c;
    jv_query_stats *st = d_stats;
    if (!st) {
        JV_TRY(c->stats.ensure((size_t)nq * sizeof(jv_query_stats)));
        st = c->stats.as<jv_query_stats>();
    }
    int launches = 0;
    JV_TRY(search_core(ix, c, d_queries, nq, p, p->accept_bits, d_out_doc, d_out_score, d_out_count, st, &launches));
    JV_CUDA_TRY(cudaStreamSynchronize(c->stream));
    fill_timing(c, timing, launches, false);
    return JV_OK;
}

int32_t jv_search_batch(jv_index *ix, const float *queries, int32_t nq, const jv_search_params *p, int32_t *out_doc,
                        float *out_score, int32_t *out_count, jv_query_stats *stats, jv_batch_timing *timing) {
    JV_TRY(validate_params(ix, nq, p));
    if (nq == 0) return JV_OK;
    JV_REQUIRE(queries && out_doc && out_score && out_count, "NULL buffer");
    if (ix->n == 0) { // empty segment: nothing to collect
        for (int64_t i = 0; i < (int64_t)nq * p->k; i++) out_doc[i] = -1, out_score[i] = 0.f;
        for (int i = 0; i < nq; i++) out_count[i] = 0;
        if (stats) memset(stats, 0, sizeof(jv_query_stats) * (size_t)nq);
        if (timing) memset(timing, 0, sizeof(*timing));
        return JV_OK;
    }
    DeviceGuard guard(ix->device);
    CtxLease lease(ix);
    JV_REQUIRE(lease.c != nullptr, "could not create a search context: %s", get_error());
    SearchCtx *c = lease.c;
    const size_t qbytes = (size_t)nq * ix->dim * 4, kb = (size_t)nq * p->k * 4;
    JV_TRY(c->queries.ensure(qbytes));
    // small batches (latency path): ids, scores, counts and counters share one device buffer and leave in ONE copy into the
    // context's page-locked staging area, from where the host splits them into the caller's buffers (4 small copies -> 1)
    const size_t sb = (size_t)nq * sizeof(jv_query_stats);
    const size_t packed_bytes = 2 * kb + (size_t)nq * 4 + sb;
    const bool packed = packed_bytes <= 64 * 1024;
    JV_TRY(c->out_doc.ensure(packed ? packed_bytes : kb));
    JV_TRY(c->out_score.ensure(kb));
    JV_TRY(c->out_count.ensure((size_t)nq * 4));
    JV_TRY(c->stats.ensure(sb));
    if (packed) JV_TRY(c->ensure_pinned(packed_bytes));
    const uint64_t *d_accept = nullptr;
    size_t abytes = 0;
    if (p->accept_bits) {
        const size_t words = (size_t)((ix->max_doc + 63) / 64);
        abytes = (p->accept_stride_words ? (size_t)(nq - 1) * p->accept_stride_words + words : words) * 8;
        JV_TRY(c->accept.ensure(abytes));
        d_accept = c->accept.as<uint64_t>();
    }
    // Queries in pinned (page-locked, UVA-mapped) host memory — the FFM shim's registered buffer, cudaHostAlloc, torch
    // pin_memory — are read by the kernels in place over PCIe: each CTA pulls its 3 KB query when it starts on it, which
    // overlaps with the other CTAs' work instead of a serial nq*dim*4-byte copy in front of the batch.  Pageable memory
    // takes the staged copy.
    const float *d_q = c->queries.as<float>();
    bool zero_copy = false;
    {
        cudaPointerAttributes at;
        void *dp = nullptr;
        // (the 8-bit table path reads every query several times in the batched table build: stage it in HBM)
        if (!ix->q8_ok && cudaPointerGetAttributes(&at, queries) == cudaSuccess && at.type == cudaMemoryTypeHost &&
            cudaHostGetDevicePointer(&dp, const_cast<float *>(queries), 0) == cudaSuccess && dp != nullptr &&
            (reinterpret_cast<uintptr_t>(dp) & 15) == 0) {
            d_q = static_cast<const float *>(dp);
            zero_copy = true;
        } else {
            cudaGetLastError(); // pageable memory: clear the sticky "invalid value" of the attribute query
        }
    }
    JV_CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
    if (abytes) JV_CUDA_TRY(cudaMemcpyAsync(c->accept.p, p->accept_bits, abytes, cudaMemcpyHostToDevice, c->stream));
    int launches = 0;
    // Large staged batches are pipelined: a small first chunk (its copy is the only one that is exposed) and then growing
    // chunks travel on a copy stream while the kernels of the previous chunks (table build, traversal, rerank) run.  Per-query accept bitsets keep the single-shot path (their stride is relative to the batch).
    int bounds[9] = {0, nq, 0, 0, 0, 0, 0, 0, 0}, nchunks = 1;
    if (!zero_copy && nq >= 4096 && !(p->accept_bits && p->accept_stride_words) && !q8_knobs().h2d_single) {
        // geometric ramp: 1/8, 1/4, then the rest in parts of <= 16384 queries; consecutive chunks alternate between two
        // contexts so that a chunk's kernels start as soon as its copy has landed, overlapping the previous chunk's tail
        const int first = nq / 8, second = first + nq / 4;
        bounds[1] = first;
        bounds[2] = second;
        int rest = nq - second, parts = (rest + 16383) / 16384;
        if (parts > 6) parts = 6;
        for (int i = 1; i <= parts; i++) bounds[2 + i] = second + (int)((int64_t)rest * i / parts);
        nchunks = 2 + parts;
    }
    if (nchunks > 1) {
        JV_CUDA_TRY(cudaEventRecord(c->chunk_ev[0], c->stream)); // the copy stream must not overtake earlier work on these buffers
        JV_CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->chunk_ev[0], 0));
        int per = 0;
        for (int ci = 0; ci < nchunks; ci++) per = bounds[ci + 1] - bounds[ci] > per ? bounds[ci + 1] - bounds[ci] : per;
        // Odd chunks run on a second context (own stream + scratch): when the persistent traversal kernel of chunk i drains,
        // the kernels of chunk i+1 fill the freed SMs instead of waiting for the last query of chunk i.
        CtxLease lease2(ix);
        JV_REQUIRE(lease2.c != nullptr, "could not create a search context: %s", get_error());
        SearchCtx *cc[2] = {c, lease2.c};
        // approximate-list scratch for the largest chunk is allocated once, before anything is enqueued
        for (SearchCtx *x : cc) {
            JV_TRY(x->approx_keys.ensure((size_t)per * p->rerank_k * 8));
            JV_TRY(x->approx_count.ensure((size_t)per * 4));
        }
        JV_CUDA_TRY(cudaStreamWaitEvent(cc[1]->stream, c->chunk_ev[0], 0)); // ordered after earlier work on the shared buffers
        for (int ci = 0; ci < nchunks; ci++) {
            const int q0 = bounds[ci], nqc = bounds[ci + 1] - q0;
            JV_CUDA_TRY(cudaMemcpyAsync(c->queries.as<float>() + (size_t)q0 * ix->dim, queries + (size_t)q0 * ix->dim, (size_t)nqc * ix->dim * 4,
                                        cudaMemcpyHostToDevice, c->copy_stream));
            JV_CUDA_TRY(cudaEventRecord(c->chunk_ev[ci], c->copy_stream));
        }
        for (int ci = 0; ci < nchunks; ci++) {
            const int q0 = bounds[ci], nqc = bounds[ci + 1] - q0;
            SearchCtx *x = cc[ci & 1];
            JV_CUDA_TRY(cudaStreamWaitEvent(x->stream, c->chunk_ev[ci], 0));
            JV_TRY(search_core(ix, x, c->queries.as<float>() + (size_t)q0 * ix->dim, nqc, p, d_accept, c->out_doc.as<int32_t>() + (size_t)q0 * p->k,
                               c->out_score.as<float>() + (size_t)q0 * p->k, c->out_count.as<int32_t>() + q0,
                               c->stats.as<jv_query_stats>() + q0, &launches, ci == 0));
        }
        JV_CUDA_TRY(cudaEventRecord(cc[1]->chunk_ev[0], cc[1]->stream)); // join: results leave on the first context's stream
        JV_CUDA_TRY(cudaStreamWaitEvent(c->stream, cc[1]->chunk_ev[0], 0));
        JV_CUDA_TRY(cudaMemcpyAsync(out_doc, c->out_doc.p, kb, cudaMemcpyDeviceToHost, c->stream));
        JV_CUDA_TRY(cudaMemcpyAsync(out_score, c->out_score.p, kb, cudaMemcpyDeviceToHost, c->stream));
        JV_CUDA_TRY(cudaMemcpyAsync(out_count, c->out_count.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, c->stream));
        if (stats)
            JV_CUDA_TRY(cudaMemcpyAsync(stats, c->stats.p, (size_t)nq * sizeof(jv_query_stats), cudaMemcpyDeviceToHost, c->stream));
        JV_CUDA_TRY(cudaEventRecord(c->ev[4], c->stream));
        JV_CUDA_TRY(cudaStreamSynchronize(c->stream)); // (lease2 must not be released before its work is done)
        fill_timing(c, timing, launches, true);
        return JV_OK;
    } else {
        if (!zero_copy) JV_CUDA_TRY(cudaMemcpyAsync(c->queries.p, queries, qbytes, cudaMemcpyHostToDevice, c->stream));
        if (packed) {
            char *base = static_cast<char *>(c->out_doc.p);
            JV_TRY(search_core(ix, c, d_q, nq, p, d_accept, reinterpret_cast<int32_t *>(base), reinterpret_cast<float *>(base + kb),
                               reinterpret_cast<int32_t *>(base + 2 * kb), reinterpret_cast<jv_query_stats *>(base + 2 * kb + (size_t)nq * 4),
                               &launches));
            JV_CUDA_TRY(cudaMemcpyAsync(c->pinned, base, stats ? packed_bytes : packed_bytes - sb, cudaMemcpyDeviceToHost, c->stream));
            JV_CUDA_TRY(cudaEventRecord(c->ev[4], c->stream));
            JV_CUDA_TRY(cudaStreamSynchronize(c->stream));
            const char *h = static_cast<const char *>(c->pinned);
            memcpy(out_doc, h, kb);
            memcpy(out_score, h + kb, kb);
            memcpy(out_count, h + 2 * kb, (size_t)nq * 4);
            if (stats) memcpy(stats, h + 2 * kb + (size_t)nq * 4, sb);
            fill_timing(c, timing, launches, true);
            return JV_OK;
        }
        JV_TRY(search_core(ix, c, d_q, nq, p, d_accept, c->out_doc.as<int32_t>(), c->out_score.as<float>(),
                           c->out_count.as<int32_t>(), c->stats.as<jv_query_stats>(), &launches));
    }
    JV_CUDA_TRY(cudaMemcpyAsync(out_doc, c->out_doc.p, kb, cudaMemcpyDeviceToHost, c->stream));
    JV_CUDA_TRY(cudaMemcpyAsync(out_score, c->out_score.p, kb, cudaMemcpyDeviceToHost, c->stream));
    JV_CUDA_TRY(cudaMemcpyAsync(out_count, c->out_count.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, c->stream));
    if (stats)
        JV_CUDA_TRY(cudaMemcpyAsync(stats, c->stats.p, (size_t)nq * sizeof(jv_query_stats), cudaMemcpyDeviceToHost, c->stream));
    JV_CUDA_TRY(cudaEventRecord(c->ev[4], c->stream));
    JV_CUDA_TRY(cudaStreamSynchronize(c->stream));
    fill_timing(c, timing, launches, true);
    return JV_OK;
}

// -------------------------------------------------------------------------------------------------
// brute force
// -------------------------------------------------------------------------------------------------
int32_t jv_exact_topk_dev(jv_index *ix, const float *d_queries, int32_t nq, int32_t k, const uint64_t *d_accept_bits,
                          int64_t accept_stride_words, int32_t *d_out_doc, float *d_out_score, int32_t *d_out_count) {
    JV_REQUIRE(ix && nq >= 0 && k >= 1, "bad arguments");
    if (nq == 0) return JV_OK;
    JV_REQUIRE(d_queries && d_out_doc && d_out_score && d_out_count, "NULL buffer");
    JV_REQUIRE(ix->n > 0, "index is empty");
    if (!ix->vectors_dev || ix->nvq_only) {
        set_error("brute force is not supported on NVQ-inline segments (JVectorQuantizedNvqVectorValues.java:33-36)");
        return JV_ERR_UNSUPPORTED;
    }
    DeviceGuard guard(ix->device);
    CtxLease lease(ix);
    JV_REQUIRE(lease.c != nullptr, "could not create a search context: %s", get_error());
    int launches = 0;
    JV_TRY(launch_exact_topk(ix, lease.c, d_queries, nq, k, d_accept_bits, accept_stride_words, d_out_doc, d_out_score,
                             d_out_count, &launches));
    JV_CUDA_TRY(cudaStreamSynchronize(lease.c->stream));
    return JV_OK;
}

int32_t jv_exact_topk(jv_index *ix, const float *queries, int32_t nq, int32_t k, const uint64_t *accept_bits,
                      int64_t accept_stride_words, int32_t *out_doc, float *out_score, int32_t *out_count) {
    JV_REQUIRE(ix && nq >= 0 && k >= 1, "bad arguments");
    if (nq == 0) return JV_OK;
    JV_REQUIRE(queries && out_doc && out_score && out_count, "NULL buffer");
    if (ix->n == 0) {
        for (int64_t i = 0; i < (int64_t)nq * k; i++) out_doc[i] = -1, out_score[i] = 0.f;
        for (int i = 0; i < nq; i++) out_count[i] = 0;
        return JV_OK;
    }
    if (!ix->vectors_dev || ix->nvq_only) {
        set_error("brute force is not supported on NVQ-inline segments (JVectorQuantizedNvqVectorValues.java:33-36)");
        return JV_ERR_UNSUPPORTED;
    }
    DeviceGuard guard(ix->device);
    CtxLease lease(ix);
    JV_REQUIRE(lease.c != nullptr, "could not create a search context: %s", get_error());
    SearchCtx *c = lease.c;
    const size_t qbytes = (size_t)nq * ix->dim * 4, kb = (size_t)nq * k * 4;
    JV_TRY(c->queries.ensure(qbytes));
    JV_TRY(c->out_doc.ensure(kb));
    JV_TRY(c->out_score.ensure(kb));
    JV_TRY(c->out_count.ensure((size_t)nq * 4));
    const uint64_t *d_accept = nullptr;
    if (accept_bits) {
        const size_t words = (size_t)((ix->max_doc + 63) / 64);
        const size_t abytes = (accept_stride_words ? (size_t)(nq - 1) * accept_stride_words + words : words) * 8;
        JV_TRY(c->accept.ensure(abytes));
        JV_CUDA_TRY(cudaMemcpyAsync(c->accept.p, accept_bits, abytes, cudaMemcpyHostToDevice, c->stream));
        d_accept = c->accept.as<uint64_t>();
    }
    JV_CUDA_TRY(cudaMemcpyAsync(c->queries.p, queries, qbytes, cudaMemcpyHostToDevice, c->stream));
    int launches = 0;
    JV_TRY(launch_exact_topk(ix, c, c->queries.as<float>(), nq, k, d_accept, accept_stride_words, c->out_doc.as<int32_t>(),
                             c->out_score.as<float>(), c->out_count.as<int32_t>(), &launches));
    JV_CUDA_TRY(cudaMemcpyAsync(out_doc, c->out_doc.p, kb, cudaMemcpyDeviceToHost, c->stream));
    JV_CUDA_TRY(cudaMemcpyAsync(out_score, c->out_score.p, kb, cudaMemcpyDeviceToHost, c->stream));
    JV_CUDA_TRY(cudaMemcpyAsync(out_count, c->out_count.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, c->stream));
    JV_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return JV_OK;
}

// -------------------------------------------------------------------------------------------------
// PQ
// -------------------------------------------------------------------------------------------------
int32_t jv_pq_encode_dev(int32_t device, const float *d_vectors, int64_t n, int32_t dim, int32_t m, int32_t k,
                         const float *d_codebooks, const float *d_gcent, uint8_t *d_out_codes, float *out_kernel_ms) {
    JV_REQUIRE(n >= 0 && dim >= 1 && m >= 1 && m <= dim && k >= 1 && k <= 256, "bad PQ shape");
    if (n == 0) return JV_OK;
    JV_REQUIRE(d_vectors && d_codebooks && d_out_codes, "NULL buffer");
    JV_TRY(check_device(device));
    DeviceGuard guard(device);
    PqShape s;
    s.init(dim, m, k);
    cudaEvent_t e0, e1;
    JV_CUDA_TRY(cudaEventCreate(&e0));
    JV_CUDA_TRY(cudaEventCreate(&e1));
    cudaEventRecord(e0, nullptr);
    int32_t st = launch_pq_encode(nullptr, s, d_vectors, n, d_codebooks, d_gcent, d_out_codes, m);
    cudaEventRecord(e1, nullptr);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    JV_TRY(st);
    JV_CUDA_TRY(e);
    if (out_kernel_ms) *out_kernel_ms = ms;
    return JV_OK;
}

int32_t jv_pq_encode(int32_t device, const float *vectors, int64_t n, int32_t dim, int32_t m, int32_t k, const float *codebooks,
                     const float *gcent, uint8_t *out_codes, float *out_kernel_ms) {
    JV_REQUIRE(n >= 0 && dim >= 1 && m >= 1 && m <= dim && k >= 1 && k <= 256, "bad PQ shape");
    if (n == 0) return JV_OK;
    JV_REQUIRE(vectors && codebooks && out_codes, "NULL buffer");
    JV_TRY(check_device(device));
    DeviceGuard guard(device);
    PqShape s;
    s.init(dim, m, k);
    DevBuf dcb, dg, dx, dout;
    JV_TRY(dcb.alloc((size_t)s.cb_floats * 4));
    JV_CUDA_TRY(cudaMemcpy(dcb.p, codebooks, (size_t)s.cb_floats * 4, cudaMemcpyHostToDevice));
    if (gcent) {
        JV_TRY(dg.alloc((size_t)dim * 4));
        JV_CUDA_TRY(cudaMemcpy(dg.p, gcent, (size_t)dim * 4, cudaMemcpyHostToDevice));
    }
    // chunked so arbitrarily large flushes/merges stream through a bounded device buffer
    const int64_t chunk = std::min<int64_t>(n, std::max<int64_t>(1, ((int64_t)1 << 30) / ((int64_t)dim * 4)));
    JV_TRY(dx.alloc((size_t)chunk * dim * 4));
    JV_TRY(dout.alloc((size_t)chunk * m));
    float total_ms = 0.f;
    for (int64_t o = 0; o < n; o += chunk) {
        const int64_t cn = std::min(chunk, n - o);
        JV_CUDA_TRY(cudaMemcpy(dx.p, vectors + o * dim, (size_t)cn * dim * 4, cudaMemcpyHostToDevice));
        float ms = 0.f;
        JV_TRY(jv_pq_encode_dev(device, dx.as<float>(), cn, dim, m, k, dcb.as<float>(), dg.as<float>(), dout.as<uint8_t>(), &ms));
        total_ms += ms;
        JV_CUDA_TRY(cudaMemcpy(out_codes + o * m, dout.p, (size_t)cn * m, cudaMemcpyDeviceToHost));
    }
    if (out_kernel_ms) *out_kernel_ms = total_ms;
    return JV_OK;
}

int32_t jv_pq_decode_dev(int32_t device, const uint8_t *d_codes, int64_t n, int32_t dim, int32_t m, int32_t k, const float *d_codebooks,
                         const float *d_gcent, float *d_out_vectors) {
    JV_REQUIRE(n >= 0 && dim >= 1 && m >= 1 && m <= dim && k >= 1 && k <= 256, "bad PQ shape");
    if (n == 0) return JV_OK;
    JV_REQUIRE(d_codes && d_codebooks && d_out_vectors, "NULL buffer");
    JV_TRY(check_device(device));
    DeviceGuard guard(device);
    PqShape s;
    s.init(dim, m, k);
    JV_TRY(launch_pq_decode(nullptr, s, d_codes, n, d_codebooks, d_gcent, d_out_vectors));
    JV_CUDA_TRY(cudaDeviceSynchronize());
    return JV_OK;
}

int32_t jv_pq_decode(int32_t device, const uint8_t *codes, int64_t n, int32_t dim, int32_t m, int32_t k, const float *codebooks,
                     const float *gcent, float *out_vectors) {
    JV_REQUIRE(n >= 0 && dim >= 1 && m >= 1 && m <= dim && k >= 1 && k <= 256, "bad PQ shape");
    if (n == 0) return JV_OK;
    JV_REQUIRE(codes && codebooks && out_vectors, "NULL buffer");
    JV_TRY(check_device(device));
    DeviceGuard guard(device);
    PqShape s;
    s.init(dim, m, k);
    DevBuf dcb, dg, dc, dout;
    JV_TRY(dcb.alloc((size_t)s.cb_floats * 4));
    JV_CUDA_TRY(cudaMemcpy(dcb.p, codebooks, (size_t)s.cb_floats * 4, cudaMemcpyHostToDevice));
    if (gcent) {
        JV_TRY(dg.alloc((size_t)dim * 4));
        JV_CUDA_TRY(cudaMemcpy(dg.p, gcent, (size_t)dim * 4, cudaMemcpyHostToDevice));
    }
    const int64_t chunk = std::min<int64_t>(n, std::max<int64_t>(1, ((int64_t)1 << 30) / ((int64_t)dim * 4)));
    JV_TRY(dc.alloc((size_t)chunk * m));
    JV_TRY(dout.alloc((size_t)chunk * dim * 4));
    for (int64_t o = 0; o < n; o += chunk) {
        const int64_t cn = std::min(chunk, n - o);
        JV_CUDA_TRY(cudaMemcpy(dc.p, codes + o * m, (size_t)cn * m, cudaMemcpyHostToDevice));
        JV_TRY(launch_pq_decode(nullptr, s, dc.as<uint8_t>(), cn, dcb.as<float>(), dg.as<float>(), dout.as<float>()));
        JV_CUDA_TRY(cudaMemcpy(out_vectors + o * dim, dout.p, (size_t)cn * dim * 4, cudaMemcpyDeviceToHost));
    }
    return JV_OK;
}

int32_t jv_pq_lut(jv_index *ix, const float *queries, int32_t nq, float *out_lut) {
    JV_REQUIRE(ix && queries && out_lut && nq >= 0, "bad arguments");
    JV_REQUIRE(ix->has_pq, "index has no PQ");
    if (nq == 0) return JV_OK;
    DeviceGuard guard(ix->device);
    DevBuf dq, dl;
    const size_t lb = (size_t)nq * ix->pq.M * ix->pq.K * 4;
    JV_TRY(dq.alloc((size_t)nq * ix->dim * 4));
    JV_TRY(dl.alloc(lb));
    JV_CUDA_TRY(cudaMemcpy(dq.p, queries, (size_t)nq * ix->dim * 4, cudaMemcpyHostToDevice));
    JV_TRY(launch_pq_lut(ix, nullptr, dq.as<float>(), nq, dl.as<float>()));
    JV_CUDA_TRY(cudaMemcpy(out_lut, dl.p, lb, cudaMemcpyDeviceToHost));
    return JV_OK;
}

int32_t jv_pq_lut_q8(jv_index *ix, const float *queries, int32_t nq, uint8_t *out_q8, float *out_params) {
    JV_REQUIRE(ix && queries && out_q8 && out_params && nq >= 0, "bad arguments");
    JV_REQUIRE(ix->has_pq && ix->q8_ok, "index was not created with JV_INDEX_FLAG_LUT_U8 (or its PQ shape is not supported by the 8-bit table)");
    if (nq == 0) return JV_OK;
    DeviceGuard guard(ix->device);
    const int M = ix->pq.M, lutb = q8_lut_bytes(ix->q8_nj);
    DevBuf dq, dl, dp;
    JV_TRY(dq.alloc((size_t)nq * ix->dim * 4));
    JV_TRY(dl.alloc((size_t)nq * lutb));
    JV_TRY(dp.alloc((size_t)nq * sizeof(float4)));
    JV_CUDA_TRY(cudaMemcpy(dq.p, queries, (size_t)nq * ix->dim * 4, cudaMemcpyHostToDevice));
    JV_TRY(launch_lut_q8(ix, nullptr, dq.as<float>(), nq, dl.as<uint8_t>(), dp.as<float4>()));
    JV_CUDA_TRY(cudaDeviceSynchronize());
    std::vector<uint8_t> raw((size_t)nq * lutb);
    std::vector<float> prm((size_t)nq * 4);
    JV_CUDA_TRY(cudaMemcpy(raw.data(), dl.p, raw.size(), cudaMemcpyDeviceToHost));
    JV_CUDA_TRY(cudaMemcpy(prm.data(), dp.p, prm.size() * 4, cudaMemcpyDeviceToHost));
    for (int q = 0; q < nq; q++) { // undo the bank-interleaved layout (header of jv_q8.cu)
        const uint8_t *t = raw.data() + (size_t)q * lutb;
        for (int m = 0; m < M; m++) {
            const int j = m >> 5;
            for (int c = 0; c < 256; c++)
                out_q8[((size_t)q * M + m) * 256 + c] = t[(size_t)(j >> 1) * 16384 + (size_t)(c & 63) * 256 + (j & 1) * 128 + (m & 31) * 4 + (c >> 6)];
        }
        out_params[2 * q] = prm[4 * (size_t)q];
        out_params[2 * q + 1] = prm[4 * (size_t)q + 1];
    }
    return JV_OK;
}

int32_t jv_pq_adc_scores(jv_index *ix, const float *queries, int32_t nq, const int32_t *nodes, int32_t per_query, float *out) {
    JV_REQUIRE(ix && queries && nodes && out && nq >= 0 && per_query >= 1, "bad arguments");
    JV_REQUIRE(ix->has_pq, "index has no PQ");
    if (nq == 0) return JV_OK;
    DeviceGuard guard(ix->device);
    DevBuf dq, dn, dout;
    JV_TRY(dq.alloc((size_t)nq * ix->dim * 4));
    JV_TRY(dn.alloc((size_t)nq * per_query * 4));
    JV_TRY(dout.alloc((size_t)nq * per_query * 4));
    JV_CUDA_TRY(cudaMemcpy(dq.p, queries, (size_t)nq * ix->dim * 4, cudaMemcpyHostToDevice));
    JV_CUDA_TRY(cudaMemcpy(dn.p, nodes, (size_t)nq * per_query * 4, cudaMemcpyHostToDevice));
    JV_TRY(launch_adc_pairs(ix, nullptr, dq.as<float>(), nq, dn.as<int32_t>(), per_query, dout.as<float>()));
    JV_CUDA_TRY(cudaMemcpy(out, dout.p, (size_t)nq * per_query * 4, cudaMemcpyDeviceToHost));
    return JV_OK;
}

// -------------------------------------------------------------------------------------------------
// merge
// -------------------------------------------------------------------------------------------------
int32_t jv_merge_topk_stream(int32_t device, int32_t g, int32_t nq, int32_t k, const int32_t *d_docs, const float *d_scores,
                             int32_t *d_out_doc, float *d_out_score, int32_t *d_out_count, void *cuda_stream) {
    JV_REQUIRE(g >= 1 && nq >= 0 && k >= 1, "bad arguments");
    if (nq == 0) return JV_OK;
    JV_REQUIRE(d_docs && d_scores && d_out_doc && d_out_score, "NULL buffer");
    JV_TRY(check_device(device));
    DeviceGuard guard(device);
    return launch_merge_topk(static_cast<cudaStream_t>(cuda_stream), g, nq, k, d_docs, d_scores, d_out_doc, d_out_score, d_out_count);
}

int32_t jv_merge_topk_dev(int32_t device, int32_t g, int32_t nq, int32_t k, const int32_t *d_docs, const float *d_scores,
                          int32_t *d_out_doc, float *d_out_score, int32_t *d_out_count, float *out_kernel_ms) {
    JV_REQUIRE(g >= 1 && nq >= 0 && k >= 1, "bad arguments");
    if (nq == 0) return JV_OK;
    JV_REQUIRE(d_docs && d_scores && d_out_doc && d_out_score, "NULL buffer");
    JV_TRY(check_device(device));
    DeviceGuard guard(device);
    cudaEvent_t e0, e1;
    JV_CUDA_TRY(cudaEventCreate(&e0));
    JV_CUDA_TRY(cudaEventCreate(&e1));
    cudaEventRecord(e0, nullptr);
    int32_t st = launch_merge_topk(nullptr, g, nq, k, d_docs, d_scores, d_out_doc, d_out_score, d_out_count);
    cudaEventRecord(e1, nullptr);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    JV_TRY(st);
    JV_CUDA_TRY(e);
    if (out_kernel_ms) *out_kernel_ms = ms;
    return JV_OK;
}

int32_t jv_merge_topk(int32_t device, int32_t g, int32_t nq, int32_t k, const int32_t *docs, const float *scores, int32_t *out_doc,
                      float *out_score, int32_t *out_count) {
    JV_REQUIRE(g >= 1 && nq >= 0 && k >= 1, "bad arguments");
    if (nq == 0) return JV_OK;
    JV_REQUIRE(docs && scores && out_doc && out_score, "NULL buffer");
    JV_TRY(check_device(device));
    DeviceGuard guard(device);
    const size_t inb = (size_t)g * nq * k * 4, outb = (size_t)nq * k * 4;
    DevBuf dd, ds, od, os, oc;
    JV_TRY(dd.alloc(inb));
    JV_TRY(ds.alloc(inb));
    JV_TRY(od.alloc(outb));
    JV_TRY(os.alloc(outb));
    JV_TRY(oc.alloc((size_t)nq * 4));
    JV_CUDA_TRY(cudaMemcpy(dd.p, docs, inb, cudaMemcpyHostToDevice));
    JV_CUDA_TRY(cudaMemcpy(ds.p, scores, inb, cudaMemcpyHostToDevice));
    JV_TRY(jv_merge_topk_dev(device, g, nq, k, dd.as<int32_t>(), ds.as<float>(), od.as<int32_t>(), os.as<float>(), oc.as<int32_t>(),
                             nullptr));
    JV_CUDA_TRY(cudaMemcpy(out_doc, od.p, outb, cudaMemcpyDeviceToHost));
    JV_CUDA_TRY(cudaMemcpy(out_score, os.p, outb, cudaMemcpyDeviceToHost));
    if (out_count) JV_CUDA_TRY(cudaMemcpy(out_count, oc.p, (size_t)nq * 4, cudaMemcpyDeviceToHost));
    return JV_OK;
}

}  // extern "C"
