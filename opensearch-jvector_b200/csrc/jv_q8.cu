// jv_q8.cu — production traversal with an 8-bit quantised ADC table (JV_INDEX_FLAG_LUT_U8).
//
// K1 (PQVectors.precomputedScoreFunctionFor, JVectorReader.java:352-357) and K2 (GraphSearcher.search + PQDecoder ADC,
// JVectorReader.java:165-173) split into two kernels so each can be laid out for the hardware:
//
//   lut_q8_kernel     one CTA builds the tables of QT queries at once: thread = subspace, the codebook streams through
//                     shared memory (cp.async, double buffered per warp) and every centroid read is reused for QT
//                     queries from registers.  Entries are quantised to bytes with ONE scale per query (the sum of M
//                     bytes is then an exact integer, independent of summation order) and written in the bank-
//                     interleaved layout the traversal wants.  Definitions = oracle q8_quantise (bit-identical).
//   q8_search_kernel  one CTA (4 warps) per query, persistent grid.  The 48 KB table (M=192) arrives with one TMA bulk
//                     copy (cp.async.bulk + mbarrier); 4 CTAs share an SM.  Per step every warp expands one candidate:
//                     reads its adjacency row (128 B, coalesced), tests the visited filter, and scores its own fresh
//                     neighbours — 8 lanes per code row, 4 rows per warp at a time, table lookups bank-conflict-free by
//                     construction.  Neighbours that beat the list's worst entry are queued and rank-merged into the
//                     sorted list once per step (re-scored duplicates dropped; 3 block barriers / step).  For DOT /
//                     EUCLIDEAN the list is ordered by the integer sum itself (no float arithmetic per row).
//
// Table layout in shared memory (and in the HBM staging buffer), j = m / 32:
//     byte (m, c)  ->  (j / 2) * 16384 + (c % 64) * 256 + (j % 2) * 128 + (m % 32) * 4 + c / 64
// i.e. bank = m % 32 and the offset of code c inside its bank column is ((c & 63) << 8) | (c >> 6): one PRMT (byte 0 from
// cw >> 6, byte 1 from cw) + one LOP3 ((v & 0x3F03) | lane base) per lookup.  Code rows are stored permuted (codes_q8):
// lane sl of a row group owns subspaces m = 8t + sl, its 4*NJ bytes contiguous, so at lookup t the 8 lanes of a group hit
// 8 different banks and the 4 groups of a warp are rotated onto the 4 different bank quarters.
#include <type_traits>

#include "jv_q8.cuh"
#include "jv_rerank_body.cuh"

namespace jv {

static Q8Knobs g_knobs;
static std::once_flag g_knobs_once;

void q8_knobs_refresh() {
    Q8Knobs v;
    if (const char *e = getenv("JVGPU_Q8_OCC")) v.occ = atoi(e);
    if (const char *e = getenv("JVGPU_Q8_CHUNK")) v.chunk = atoi(e);
    if (const char *e = getenv("JVGPU_Q8_WARPS")) v.warps = atoi(e);
    v.prof = getenv("JVGPU_PROFILE") != nullptr;
    v.fused = getenv("JVGPU_Q8_FUSED") != nullptr;
    v.sync = getenv("JVGPU_Q8_SYNC") != nullptr;
    if (const char *e = getenv("JVGPU_Q8_DEPTH")) v.depth = atoi(e);
    if (const char *e = getenv("JVGPU_RERANK_DEDUPE")) v.rerank_dedupe = atoi(e) != 0 ? 1 : 0;
    v.h2d_single = getenv("JVGPU_H2D_SINGLE") != nullptr;
    if (const char *e = getenv("JVGPU_EXACT_TC")) v.exact_tc = atoi(e) != 0 ? 1 : 0;
    g_knobs = v;
}

const Q8Knobs &q8_knobs() {
    std::call_once(g_knobs_once, q8_knobs_refresh);
    return g_knobs;
}

// packed fp32 pairs (FFMA2 on sm_100a): two IEEE fmas per instruction, each lane rounded exactly like a scalar fmaf
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint32_t sat_u8_rn(float x) {
    uint32_t u;
    asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(u) : "f"(x));
    return u;
}

// ---------------------------------------------------------------------------------------------------------------
// index-creation helper: codes [n][stride] -> codes_q8 [n][MP] (lane-major permutation, zero padded)
// ---------------------------------------------------------------------------------------------------------------
__global__ void permute_codes_kernel(const uint8_t *__restrict__ codes, int64_t n, int M, int stride, int NJ, uint8_t *__restrict__ out) {
    const int MP = NJ * 32, seg = NJ * 4;
    const int64_t total = n * MP;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / MP;
        const int o = (int)(i - row * MP);
        const int sl = o / seg, t = o - sl * seg;
        const int m = 8 * t + sl;
        out[row * MP + q8_word_offset(NJ, sl, t >> 2) + (t & 3)] = m < M ? codes[row * stride + m] : (uint8_t)0;
    }
}

int32_t launch_permute_codes(cudaStream_t stream, const uint8_t *d_codes, int64_t n, int M, int stride, int NJ, uint8_t *d_out) {
    if (n == 0) return JV_OK;
    const int64_t total = n * NJ * 32;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 64) blocks = 148 * 64;
    permute_codes_kernel<<<(int)blocks, 256, 0, stream>>>(d_codes, n, M, stride, NJ, d_out);
    JV_CUDA_TRY(cudaGetLastError());
    return JV_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// K1: batched 8-bit table build
// ---------------------------------------------------------------------------------------------------------------
struct LutQ8Params {
    const float *queries; // [nq][dim]
    const float *codebooks, *gcent, *ball_ctr, *ball_rad;
    uint8_t *lut;   // [nq][lutb]
    float4 *qparams; // [nq] (delta, base, ||q||^2, 0)
    int nq, dim, M, NJ, sim, lutb;
};

template <int S, bool L2>
__global__ void __launch_bounds__(256, 2) lut_q8_kernel(const LutQ8Params p) {
    constexpr int QT = S == 2 ? 8 : 32 / S;     // queries per CTA (QT * S <= 32 query registers per thread)
    constexpr int CH = 32 / S;                 // codes per staged chunk: 128 B per subspace
    constexpr int PB = S == 2 ? 8 : 16;        // cp.async piece
    constexpr int PIECES = 128 / PB;           // per subspace per chunk
    constexpr int ROWB = 128 + (S == 2 ? 8 : 16); // padded row: conflict-free vector reads at lane stride
    constexpr int NCHUNK = 256 / CH;
    constexpr int EP = CH / 4;                 // consecutive codes per 64-code quarter in a chunk (32 B per quarter)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const int MP = p.NJ * 32;
    unsigned char *stage = smem_raw + (size_t)warp * 2 * 32 * ROWB;
    float *lo_s = reinterpret_cast<float *>(smem_raw + (size_t)nw * 2 * 32 * ROWB); // [QT][MP]
    __shared__ uint32_t range_s[QT];
    __shared__ float inv_s[QT], qn_s[QT];
    const int q0 = blockIdx.x * QT;
    constexpr bool l2 = L2;

    if (tid < QT) range_s[tid] = 0u;
    __syncthreads();

    auto load_q = [&](int m, float (&qr)[QT][S]) {
#pragma unroll
        for (int t = 0; t < QT; t++) {
            const bool ok = m < p.M && q0 + t < p.nq;
#pragma unroll
            for (int j = 0; j < S; j++) {
                float v = ok ? __ldg(p.queries + (int64_t)(q0 + t) * p.dim + m * S + j) : 0.f;
                if (ok && l2 && p.gcent) v = __fsub_rn(v, __ldg(p.gcent + m * S + j));
                qr[t][j] = v;
            }
        }
    };

    // ---- phase 0: per-subspace bounds from the bounding ball (oracle q8_bounds), shared range per query
    for (int j = warp; j < p.NJ; j += nw) {
        const int m = j * 32 + lane;
        float qr[QT][S];
        load_q(m, qr);
        float ctr[S], rad = 0.f;
#pragma unroll
        for (int jj = 0; jj < S; jj++) ctr[jj] = m < p.M ? __ldg(p.ball_ctr + m * S + jj) : 0.f;
        if (m < p.M) rad = __ldg(p.ball_rad + m);
#pragma unroll
        for (int t = 0; t < QT; t++) {
            float lo = 0.f, r = 0.f;
            if (m < p.M) {
                if (l2) {
                    float acc = 0.f;
#pragma unroll
                    for (int jj = 0; jj < S; jj++) {
                        const float d = __fsub_rn(qr[t][jj], ctr[jj]);
                        acc = __fmaf_rn(d, d, acc);
                    }
                    const float d = __fsqrt_rn(acc);
                    float a = __fsub_rn(d, rad);
                    if (a < 0.f) a = 0.f;
                    const float b = __fadd_rn(d, rad);
                    lo = __fmul_rn(a, a);
                    r = __fsub_rn(__fmul_rn(b, b), lo);
                } else {
                    float qq = 0.f, qc = 0.f;
#pragma unroll
                    for (int jj = 0; jj < S; jj++) qq = __fmaf_rn(qr[t][jj], qr[t][jj], qq);
#pragma unroll
                    for (int jj = 0; jj < S; jj++) qc = __fmaf_rn(qr[t][jj], ctr[jj], qc);
                    const float w = __fmul_rn(__fsqrt_rn(qq), rad);
                    lo = __fsub_rn(qc, w);
                    r = __fsub_rn(__fadd_rn(qc, w), lo);
                }
            }
            lo_s[t * MP + m] = lo;
            if (r > 0.f) atomicMax(&range_s[t], __float_as_uint(r)); // r >= 0: the bit pattern orders like the value
        }
    }
    __syncthreads();
    // ||q||^2 (canonical order) for the cosine decoder: one warp per query
    for (int t = warp; t < QT; t += nw) {
        float qn = 0.f;
        if (q0 + t < p.nq) {
            const float *gq = p.queries + (int64_t)(q0 + t) * p.dim;
            qn = jv_warp_reduce_pair<false>(gq, gq, p.dim, lane, (p.dim & 3) == 0 && (reinterpret_cast<uintptr_t>(gq) & 15) == 0);
        }
        if (lane == 0) qn_s[t] = qn;
    }
    __syncthreads();
    if (tid < QT) { // (every code-range slice of a tile computes the same scale; slice 0 publishes it)
        const float range = __uint_as_float(range_s[tid]);
        const float inv = range > 0.f ? __fdiv_rn(255.0f, range) : 0.f;
        float base = 0.f;
        for (int m = 0; m < p.M; m++) base = __fadd_rn(base, lo_s[tid * MP + m]);
        inv_s[tid] = inv;
        if (q0 + tid < p.nq && blockIdx.y == 0) p.qparams[q0 + tid] = make_float4(__fdiv_rn(range, 255.0f), base, qn_s[tid], 0.f);
    }
    __syncthreads();

    // ---- main pass: thread = subspace (lane) of block j; chunks of CH codes staged per warp
    uint32_t *obase = reinterpret_cast<uint32_t *>(p.lut + (int64_t)q0 * p.lutb) + lane;
    const uint32_t tstride = (uint32_t)p.lutb >> 2;
    const bool full = q0 + QT <= p.nq;
    for (int j = warp; j < p.NJ; j += nw) {
        const int m = j * 32 + lane;
        float qr[QT][S];
        load_q(m, qr);
        // queries are processed in pairs with FFMA2: same fmaf chain per query, half the instructions
        uint64_t q2[QT / 2][S], inv2[QT / 2], nlo2[QT / 2];
#pragma unroll
        for (int t = 0; t < QT / 2; t++) {
#pragma unroll
            for (int jj = 0; jj < S; jj++) q2[t][jj] = pack2(qr[2 * t][jj], qr[2 * t + 1][jj]);
            // lanes past the last subspace (M not a multiple of 32) get scale 0: their entries round to 0 (NaN -> 0 as well),
            // so the padding bytes of the table need no select at the store
            const float i0 = m < p.M ? inv_s[2 * t] : 0.f, i1 = m < p.M ? inv_s[2 * t + 1] : 0.f;
            inv2[t] = pack2(i0, i1);
            nlo2[t] = pack2(-__fmul_rn(lo_s[(2 * t) * MP + m], i0), -__fmul_rn(lo_s[(2 * t + 1) * MP + m], i1));
        }
        const uint64_t neg1 = pack2(-1.0f, -1.0f);
        // a chunk holds, for each of the 32 subspaces, the codes {q*64 + ch*EP + e : q < 4, e < EP}: the 4 codes that
        // share one table word ((c & 63) fixed, byte = c >> 6) sit in the same chunk
        auto issue = [&](int ch, int buf) {
#pragma unroll
            for (int r = 0; r < PIECES; r++) {
                const int pid = lane + 32 * r;
                const int b = pid / PIECES, piece = pid % PIECES;
                const int quarter = piece / (PIECES / 4), within = piece % (PIECES / 4);
                const int mm = j * 32 + b;
                if (mm < p.M)
                    cp_async<PB>(stage + (size_t)(buf * 32 + b) * ROWB + piece * PB,
                                 reinterpret_cast<const unsigned char *>(p.codebooks + ((int64_t)mm * 256 + quarter * 64 + ch * EP) * S) + within * PB);
            }
            cp_async_commit();
        };
        // small batches: gridDim.y CTAs share a tile, each builds the entries of a range of code chunks (latency: one query's table
        // is the work of NJ warps x gridDim.y CTAs instead of NJ warps)
        const int ch0 = (int)blockIdx.y * NCHUNK / (int)gridDim.y, ch1 = ((int)blockIdx.y + 1) * NCHUNK / (int)gridDim.y;
        issue(ch0, ch0 & 1);
        for (int ch = ch0; ch < ch1; ch++) {
            const int buf = ch & 1;
            if (ch + 1 < ch1) {
                issue(ch + 1, buf ^ 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncwarp();
            const unsigned char *row = stage + (size_t)(buf * 32 + lane) * ROWB;
#pragma unroll
            for (int e = 0; e < EP; e++) {
                int qv[QT][4]; // rounded entries of the 4 codes that share a table word
#pragma unroll
                for (int quarter = 0; quarter < 4; quarter++) {
                    float cv[S];
                    const unsigned char *cp = row + (quarter * EP + e) * S * 4;
                    if (S == 2) {
                        const float2 v = *reinterpret_cast<const float2 *>(cp);
                        cv[0] = v.x, cv[1] = v.y;
                    } else {
#pragma unroll
                        for (int h = 0; h < S / 4; h++) {
                            const float4 v = *reinterpret_cast<const float4 *>(cp + h * 16);
                            cv[4 * h] = v.x, cv[4 * h + 1] = v.y, cv[4 * h + 2] = v.z, cv[4 * h + 3] = v.w;
                        }
                    }
                    uint64_t cv2[S];
#pragma unroll
                    for (int jj = 0; jj < S; jj++) cv2[jj] = pack2(cv[jj], cv[jj]);
#pragma unroll
                    for (int t = 0; t < QT / 2; t++) {
                        uint64_t acc = 0ull; // (+0.f, +0.f)
#pragma unroll
                        for (int jj = 0; jj < S; jj++) {
                            if (l2) {
                                const uint64_t d = ffma2(cv2[jj], neg1, q2[t][jj]); // q - c, rounded once like __fsub_rn
                                acc = ffma2(d, d, acc);
                            } else {
                                acc = ffma2(q2[t][jj], cv2[jj], acc);
                            }
                        }
                        float e0, e1;
                        unpack2(ffma2(acc, inv2[t], nlo2[t]), e0, e1);
                        qv[2 * t][quarter] = __float2int_rn(e0); // NaN -> 0; saturated to a byte below
                        qv[2 * t + 1][quarter] = __float2int_rn(e1);
                    }
                }
                const int c6 = ch * EP + e; // = c & 63
                const uint32_t word = (uint32_t)((j >> 1) * 16384 + c6 * 256 + (j & 1) * 128) / 4u;
                uint32_t *ow = obase + word; // query t of the tile: + t * tstride words (one pointer bump per store)
#pragma unroll
                for (int t = 0; t < QT; t++) {
                    // two cvt.pack.sat.u8.s32: bytes (q0, q1, q2, q3) = clamp(entry, 0, 255)
                    uint32_t hi, w;
                    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(qv[t][3]), "r"(qv[t][2]), "r"(0));
                    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(w) : "r"(qv[t][1]), "r"(qv[t][0]), "r"(hi));
                    if (full || q0 + t < p.nq) *ow = w;
                    ow += tstride;
                }
            }
            __syncwarp(); // the buffer is refilled two iterations from now
        }
    }
}

int q8_lut_bytes(int nj) { return ((nj + 1) / 2) * 16384; }

static size_t lut_q8_smem(int nw, int S, int MP) {
    const int rowb = 128 + (S == 2 ? 8 : 16), qt = S == 2 ? 8 : 32 / S;
    return (size_t)nw * 2 * 32 * rowb + (size_t)qt * MP * 4;
}

int32_t launch_lut_q8(jv_index *ix, cudaStream_t stream, const float *d_queries, int nq, uint8_t *d_lut, float4 *d_qparams) {
    JV_REQUIRE(ix->has_pq && ix->q8_ok, "index has no 8-bit table support (needs K = 256 and a uniform sub-vector size of 2, 4 or 8)");
    LutQ8Params p;
    p.queries = d_queries;
    p.codebooks = ix->codebooks.as<float>();
    p.gcent = ix->gcent.as<float>();
    p.ball_ctr = ix->ball_ctr.as<float>();
    p.ball_rad = ix->ball_rad.as<float>();
    p.lut = d_lut;
    p.qparams = d_qparams;
    p.nq = nq;
    p.dim = ix->dim;
    p.M = ix->pq.M;
    p.NJ = ix->q8_nj;
    p.sim = ix->sim;
    p.lutb = q8_lut_bytes(ix->q8_nj);
    const int S = ix->dim / ix->pq.M;
    const int nw = p.NJ < 8 ? p.NJ : 8;
    const size_t smem = lut_q8_smem(nw, S, p.NJ * 32);
    const int qt = S == 2 ? 8 : 32 / S;
    const int grid = (nq + qt - 1) / qt;
    // code-range slices per tile: fill the machine when the batch alone does not (2 CTAs per SM)
    int slices = 1;
    while (slices < 8 && grid * slices * 2 <= 2 * ix->sm_count) slices *= 2;
    const bool l2 = ix->sim == JV_SIM_EUCLIDEAN;
#define JV_LUT_LAUNCH(SV, LV)                                                                                              \
    do {                                                                                                                   \
        JV_CUDA_TRY(cudaFuncSetAttribute(lut_q8_kernel<SV, LV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        lut_q8_kernel<SV, LV><<<dim3(grid, slices), nw * 32, smem, stream>>>(p);                                                          \
    } while (0)
    if (S == 4) {
        if (l2) JV_LUT_LAUNCH(4, true); else JV_LUT_LAUNCH(4, false);
    } else if (S == 2) {
        if (l2) JV_LUT_LAUNCH(2, true); else JV_LUT_LAUNCH(2, false);
    } else {
        if (l2) JV_LUT_LAUNCH(8, true); else JV_LUT_LAUNCH(8, false);
    }
#undef JV_LUT_LAUNCH
    JV_CUDA_TRY(cudaGetLastError());
    return JV_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// K2: traversal
// ---------------------------------------------------------------------------------------------------------------
// NJ_T > 0: code words per lane known at compile time (registers, all loads of U rows in flight before the first lookup).
// W = warps per CTA (4 row groups each); PROF = per-phase cycle counters (jv_index_debug_counter).
// FILT = accept bits given: the list also holds rejected nodes (they are traversed, never returned), up to Lc entries; only
// entries ranked before the L-th ACCEPTED one can still be expanded (the reference stops when the best candidate is worse than
// the worst of rerankK accepted results), and that entry's key is the admission threshold for new nodes.
// CTAs per SM the register allocation must allow.  The kernel is latency-bound and its throughput follows the queries in
// flight per SM, which shared memory caps at 227 KB / (table + lists + filter): 4 at M = 192 (48 KB tables),
// but 8 / 6 / 5 with the 16 / 24 / 32 KB tables of M <= 64 / 96 / 128 — there the register file must not be the limit
// (65536 / (128 threads * CTAs) registers per thread).
constexpr int q8_min_ctas(int nj, int warps, bool filt) {
    return filt ? 3 : warps != 4 ? 4 : (nj == 1 || nj == 2) ? 8 : nj == 3 ? 6 : nj == 4 ? 5 : 4;
}

// FUSE = K3 as the epilogue (diagnostic, JVGPU_Q8_FUSED): a separate instantiation so that the production kernel does not
// carry the rerank code (the step loop alone is larger than the 32 KB L1.5 instruction cache).
template <int NJ_T, int W, bool PROF, bool FILT, bool FUSE>
__global__ void __launch_bounds__(W * 32, q8_min_ctas(NJ_T, W, FILT)) q8_search_kernel(const Q8Params p) {
    constexpr int kQW = W, kQThreads = W * 32, NG = 4 * W;
    constexpr int U = W == 4 ? 3 : 2; // rows in flight per row group and pass: NG * U rows >= the fresh neighbours of a typical step
    constexpr int NJC = NJ_T > 0 ? NJ_T : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = p.L, E = p.E, H = 1 << p.hash_log2, R = p.R;
    const int Lc = FILT ? p.Lc : L; // list capacity
    const int NJ = NJ_T > 0 ? NJ_T : p.NJ;

    unsigned char *sp = smem_raw;
    const uint8_t *lut = sp;
    sp += p.lutb;
    uint64_t *list0 = reinterpret_cast<uint64_t *>(sp);
    sp += (size_t)Lc * 8;
    uint64_t *list1 = reinterpret_cast<uint64_t *>(sp);
    sp += (size_t)Lc * 8;
    uint64_t *surv = reinterpret_cast<uint64_t *>(sp); // queued survivors as key >> 1 (always unexpanded); 0 = dropped duplicate
    sp += (size_t)p.surv_cap * 8;
    int32_t *pool = reinterpret_cast<int32_t *>(sp); // fresh neighbour ids of the current step
    sp += (size_t)p.surv_cap * 4;
    int32_t *gapcnt = reinterpret_cast<int32_t *>(sp); // [L + 1] survivors per list gap (merge)
    sp += (size_t)((Lc + 2) & ~1) * 4;
    uint32_t *filter = reinterpret_cast<uint32_t *>(sp);

    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_query, s_ns[2], s_nn[2], s_ndup[2], w_sel[kQW][2 * kQMaxE], w_pos[kQW][2 * kQMaxE];
    // filtered queries: (an upper bound of) the accepted entries in the list and a lower bound of the first unexpanded position —
    // while fewer than L entries are accepted there is no admission threshold inside the list and every unexpanded entry is
    // eligible, so the selection scans from the first unexpanded entry and stops at 2E instead of walking the whole list
    __shared__ int s_nacc, s_first;

    const bool tagged = p.n <= ((int64_t)1 << (p.hash_log2 + 15));
    const bool isum_keys = p.sim != JV_SIM_COSINE;
    const bool l2 = p.sim == JV_SIM_EUCLIDEAN;
    // ADC lane geometry: group g = lane / 8 scores one code row, lane sl owns subspaces m = 8t + sl.  At lookup (j, i) the
    // group reads bank quarter (i + g) & 3, so the 32 lanes of a warp always hit 32 different banks.
    const int g = lane >> 3, sl = lane & 7;
    // per code word (4 codes): one mask (c & 63 in every byte) and one shift-mask-or (c >> 6 plus the lane's bank bits in every
    // byte: the byte used at slot i carries the bank quarter (i + g) & 3); then ONE byte permute per lookup assembles the clean
    // offset ((c & 63) << 8) | bank * 4 | (c >> 6) — bytes 2 and 3 come from the replicated sign bit of a (c & 63) byte, i.e. zero
    uint32_t sel[4], lbw = 0u;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t qd = (uint32_t)(i + g) & 3u;
        sel[i] = 0x8800u | (qd << 4) | (4u + qd);
        lbw |= (qd * 32u + (uint32_t)sl * 4u) << (8u * qd);
    }

    if (tid == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t phase = 0;

    // table entries selected by code word j of this lane (4 subspaces); caller reduces over the 8 lanes of the group
    auto lookup4 = [&](uint32_t cw, int j) -> uint32_t {
        const uint32_t lo6 = cw & 0x3F3F3F3Fu, hi2 = ((cw >> 6) & 0x03030303u) | lbw;
        const uint8_t *base = lut + (j >> 1) * 16384 + (j & 1) * 128;
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t off; // (__byte_perm ignores the sign-replication bit of the selector nibbles: PTX prmt in its default mode)
            asm("prmt.b32 %0, %1, %2, %3;" : "=r"(off) : "r"(lo6), "r"(hi2), "r"(sel[i]));
            s += base[off];
        }
        return s;
    };
    auto reduce8 = [&](uint32_t s) -> uint32_t {
        s += __shfl_xor_sync(JV_FULL_MASK, s, 4);
        s += __shfl_xor_sync(JV_FULL_MASK, s, 2);
        s += __shfl_xor_sync(JV_FULL_MASK, s, 1);
        return s;
    };
    auto row_sum = [&](int32_t nb) -> uint32_t { // one row per group, loads interleaved with lookups (entry node, generic NJ)
        uint32_t s = 0;
        if (nb >= 0) {
            const unsigned char *row = p.codes_q8 + (int64_t)nb * (NJ * 32);
            for (int j = 0; j < NJ; j++) s += lookup4(__ldg(reinterpret_cast<const uint32_t *>(row + q8_word_offset(NJ, sl, j))), j);
        }
        return reduce8(s);
    };

    for (;;) {
        __syncthreads(); // everyone is done with the previous query's table, lists and counters
        if (tid == 0) {
            s_query = atomicAdd(p.work_counter, 1);
            s_ns[0] = 0;
            s_nn[0] = 0;
            s_ndup[0] = 0;
        }
        __syncthreads();
        const int qi = s_query;
        if (qi >= p.nq) break;
        long long ck[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; // per-phase cycles (thread 0), see jv_index_debug_counter
        long long t_prev = PROF ? clock64() : 0;
#define JV_PHASE(i)                            \
    if (PROF) {                                \
        const long long t_now = clock64();     \
        ck[i] += t_now - t_prev;               \
        t_prev = t_now;                        \
    }
        if (tid == 0) { // K1 result: one TMA bulk copy HBM/L2 -> shared memory
            mbar_expect_tx(&s_bar, (uint32_t)p.lutb);
            bulk_g2s(smem_raw, p.lut + (int64_t)qi * p.lutb, (uint32_t)p.lutb, &s_bar);
        }
        for (int i = tid; i < H; i += kQThreads) filter[i] = tagged ? 0u : kEmpty;
        for (int i = tid; i <= Lc; i += kQThreads) gapcnt[i] = 0;
        const float4 qp = __ldg(p.qparams + qi);
        const float delta = qp.x, base = qp.y, qnorm = qp.z;
        auto score_of = [&](uint32_t isum, int32_t nb) -> float {
            const float s = __fmaf_rn(delta, (float)isum, base);
            const float nn = p.sim == JV_SIM_COSINE ? __ldg(p.node_norm + nb) : 0.f;
            return adc_finish(p.sim, s, nn, qnorm);
        };
        auto ord_of = [&](uint32_t isum, int32_t nb) -> uint32_t {
            if (isum_keys) return l2 ? ~isum : isum;
            return jv_f2ord(score_of(isum, nb));
        };
        const uint64_t *abits = FILT ? p.accept + (int64_t)qi * p.accept_stride : nullptr;
        auto pack_key = [&](uint32_t isum, int32_t nb) -> uint64_t { // full list key of a freshly scored node (unexpanded)
            if (FILT) {
                const int32_t doc = p.ord_to_doc ? __ldg(p.ord_to_doc + nb) : nb; // a9: accept-bits lambda, JVectorReader.java:157-163
                const bool acc = doc >= 0 && ((__ldg(abits + (doc >> 6)) >> (doc & 63)) & 1ull);
                return qkey_pack_f(ord_of(isum, nb), nb, acc);
            }
            return qkey_pack(ord_of(isum, nb), nb);
        };
        auto node_of = [&](uint64_t k) -> int32_t { return FILT ? qkey_node_f(k) : qkey_node(k); };
        __syncthreads(); // filter cleared
        mbar_wait(&s_bar, phase);
        phase ^= 1u;

        int n = 0, cur = 0, visited = 0, expanded = 0, step = 0;
        if (p.entry >= 0 && p.entry < p.n) {
            if (warp == 0) {
                const uint32_t s = row_sum(g == 0 ? p.entry : -1);
                if (lane == 0) {
                    const uint64_t ek = pack_key(s, p.entry);
                    list0[0] = ek;
                    q_filter_insert(filter, p.hash_log2, tagged, p.entry);
                    if (FILT) {
                        s_nacc = (ek & 2ull) ? 1 : 0;
                        s_first = 0;
                    }
                }
            }
            n = 1;
            visited = 1;
        }
        __syncthreads();
        JV_PHASE(0)

        while (n > 0) {
            uint64_t *list = cur ? list1 : list0, *out = cur ? list0 : list1;
            const int par = step & 1;
            // ---- (a) every warp scans the list flags itself: no serial section, no barrier
            int found = 0;
            uint64_t worst;
            if (!FILT) {
                for (int c0 = 0; c0 < n; c0 += 32) {
                    const int i = c0 + lane;
                    const bool un = i < n && (list[i] & 1ull);
                    const uint32_t ballot = __ballot_sync(JV_FULL_MASK, un);
                    const int rank = found + __popc(ballot & ((1u << lane) - 1u));
                    if (un && rank < 2 * E) { // ranks E..2E-1 are runners-up: their rows are prefetched into L2
                        w_sel[warp][rank] = qkey_node(list[i]);
                        w_pos[warp][rank] = i;
                    }
                    found += __popc(ballot);
                    if (found >= 2 * E) break;
                }
                worst = n >= L ? (list[L - 1] >> 1) : 0ull;
            } else if (s_nacc < L) {
                // fewer than L accepted entries in the whole list: no threshold, every unexpanded entry is eligible
                for (int c0 = s_first & ~31; c0 < n; c0 += 32) {
                    const int i = c0 + lane;
                    const uint64_t k = i < n ? list[i] : 0ull;
                    const bool un = i < n && (k & 1ull);
                    const uint32_t ballot = __ballot_sync(JV_FULL_MASK, un);
                    const int rank = found + __popc(ballot & ((1u << lane) - 1u));
                    if (un && rank < 2 * E) {
                        w_sel[warp][rank] = qkey_node_f(k);
                        w_pos[warp][rank] = i;
                    }
                    found += __popc(ballot);
                    if (found >= 2 * E) break;
                }
                worst = n >= Lc ? (list[Lc - 1] >> 1) : 0ull;
                __syncwarp();
                if (tid == 0 && found > 0) s_first = w_pos[0][0]; // entries in front of it are expanded (they only move down)
            } else {
                // only entries with fewer than L accepted entries in front of them can still be expanded; the L-th accepted
                // entry's key is the admission threshold (the worst of rerankK accepted results)
                int nacc = 0;
                uint64_t thr = 0ull;
                bool have_thr = false;
                for (int c0 = 0; c0 < n && !have_thr; c0 += 32) {
                    const int i = c0 + lane;
                    const uint64_t k = i < n ? list[i] : 0ull;
                    const bool acc = (k & 2ull) != 0ull;
                    const uint32_t aball = __ballot_sync(JV_FULL_MASK, acc);
                    const int before = nacc + __popc(aball & ((1u << lane) - 1u)); // accepted entries strictly in front of me
                    const bool un = i < n && (k & 1ull) && before < L;
                    const uint32_t ballot = __ballot_sync(JV_FULL_MASK, un);
                    const int rank = found + __popc(ballot & ((1u << lane) - 1u));
                    if (un && rank < 2 * E && found < 2 * E) {
                        w_sel[warp][rank] = qkey_node_f(k);
                        w_pos[warp][rank] = i;
                    }
                    found += __popc(ballot);
                    const uint32_t lth = __ballot_sync(JV_FULL_MASK, acc && before == L - 1);
                    if (lth) {
                        thr = __shfl_sync(JV_FULL_MASK, k, __ffs(lth) - 1) >> 1;
                        have_thr = true;
                    }
                    nacc += __popc(aball);
                }
                worst = have_thr ? thr : (n >= Lc ? (list[Lc - 1] >> 1) : 0ull);
            }
            __syncwarp();
            const int nsel = found < E ? found : E;
            if (nsel == 0) break;
            JV_PHASE(2)

            // ---- (b1) warp w expands candidates w, w + 4, ..: adjacency row -> visited filter -> shared pool of fresh ids
            for (int e = warp; e < nsel; e += kQW) {
                const int32_t cand = w_sel[warp][e];
                if (nsel + e < found && nsel + e < 2 * E) { // runner-up row -> L2 for the next step
                    const int lines = (R * 4 + 127) >> 7;
                    if (lane < lines)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(p.adjacency + (int64_t)w_sel[warp][nsel + e] * R) + lane * 128));
                }
                for (int r0 = 0; r0 < R; r0 += 32) {
                    int32_t nb = -1;
                    if (r0 + lane < R) nb = __ldg(p.adjacency + (int64_t)cand * R + r0 + lane);
                    const bool fresh = nb >= 0 && nb < p.n && q_filter_insert(filter, p.hash_log2, tagged, nb);
                    const uint32_t ballot = __ballot_sync(JV_FULL_MASK, fresh);
                    int slot = 0;
                    if (lane == 0 && ballot) slot = atomicAdd(&s_nn[par], __popc(ballot));
                    slot = __shfl_sync(JV_FULL_MASK, slot, 0);
                    if (fresh) {
                        pool[slot + __popc(ballot & ((1u << lane) - 1u))] = nb;
                        // the code row is read after the pool barrier by whichever group the row is dealt to: start the
                        // DRAM -> L2 transfer now (rows are 32-byte aligned and <= 256 bytes here: the lines of the first and
                        // of the last byte cover them; longer rows get their middle lines with the demand loads)
                        const char *row = reinterpret_cast<const char *>(p.codes_q8 + (int64_t)nb * (NJ * 32));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
                        if (NJ * 32 > 128 || (reinterpret_cast<uintptr_t>(row) & 127) + NJ * 32 > 128)
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(row + NJ * 32 - 1));
                    }
                }
            }
            __syncthreads(); // B0: the pool of fresh neighbours is complete
            const int nn = s_nn[par];
            JV_PHASE(3)

            // ---- (b2) ADC of the pooled rows, spread evenly over the 16 row groups of the CTA: all code words of a
            //           group's U rows are in flight before the first table lookup (one DRAM round trip per step)
            {
                const int gid = warp * 4 + g;
                auto offer = [&](uint32_t isum, int32_t node) { // re-scored list members are dropped in (c)
                    const uint64_t a = pack_key(isum, node) >> 1;
                    if (a > worst) surv[atomicAdd(&s_ns[par], 1)] = a;
                };
                if (NJ_T > 0) {
                    // one pass over NU rows per group; NU is warp-uniform (rows are dealt to the warps 4 at a time)
                    auto pass = [&](int i0, auto nu_tag) {
                        constexpr int NU = decltype(nu_tag)::value;
                        uint32_t cw[NU][NJC], s[NU];
                        int32_t nbv[NU];
#pragma unroll
                        for (int u = 0; u < NU; u++) {
                            const int idx = i0 + u * NG + gid;
                            nbv[u] = idx < nn ? pool[idx] : -1;
                            s[u] = 0u;
                            if (nbv[u] >= 0) {
                                q8_load_row<NJC>(p.codes_q8 + (int64_t)nbv[u] * (NJ * 32), sl, cw[u]); // 16-byte vector loads
                            } else {
#pragma unroll
                                for (int j = 0; j < NJC; j++) cw[u][j] = 0u;
                            }
                        }
                        if (PROF) { // sub-phase 8: code words in registers
                            uint32_t acc = 0;
#pragma unroll
                            for (int u = 0; u < NU; u++)
#pragma unroll
                                for (int j = 0; j < NJC; j++) acc |= cw[u][j];
                            asm volatile("" ::"r"(acc));
                            JV_PHASE(8)
                        }
#pragma unroll
                        for (int j = 0; j < NJC; j++) {
#pragma unroll
                            for (int u = 0; u < NU; u++) s[u] += lookup4(cw[u][j], j); // a group without a row looks up code 0: harmless
                        }
#pragma unroll
                        for (int u = 0; u < NU; u++) s[u] = reduce8(s[u]);
                        if (PROF) {
                            uint32_t acc = 0;
#pragma unroll
                            for (int u = 0; u < NU; u++) acc |= s[u];
                            asm volatile("" ::"r"(acc));
                            JV_PHASE(9)
                        }
                        // queue the rows that beat the list's worst entry.  After the butterfly every lane of a group holds the
                        // group's sums: lane sl = u takes row u, so one key, one ballot and one shared-memory atomic serve all
                        // NU rows of the 4 groups
                        uint32_t my_s = s[0];
                        int32_t my_nb = nbv[0];
#pragma unroll
                        for (int u = 1; u < NU; u++) {
                            my_s = sl == u ? s[u] : my_s;
                            my_nb = sl == u ? nbv[u] : my_nb;
                        }
                        const uint64_t ka = (sl < NU && my_nb >= 0) ? (pack_key(my_s, my_nb) >> 1) : 0ull;
                        const uint32_t bal = __ballot_sync(JV_FULL_MASK, ka > worst);
                        if (bal) { // warp-uniform
                            int slot = 0;
                            if (lane == 0) slot = atomicAdd(&s_ns[par], __popc(bal));
                            slot = __shfl_sync(JV_FULL_MASK, slot, 0);
                            if ((bal >> lane) & 1u) surv[slot + __popc(bal & ((1u << lane) - 1u))] = ka;
                        }
                        JV_PHASE(10)
                    };
                    for (int i0 = 0; i0 < nn; i0 += NG * U) {
                        const int left = nn - i0 - warp * 4; // rows of this pass at or after this warp's first group
                        if (U >= 3 && left > 2 * NG)
                            pass(i0, std::integral_constant<int, U >= 3 ? 3 : 1>());
                        else if (left > NG)
                            pass(i0, std::integral_constant<int, 2>());
                        else if (left > 0)
                            pass(i0, std::integral_constant<int, 1>());
                    }
                } else {
                    for (int i0 = 0; i0 < nn; i0 += NG) {
                        const int idx = i0 + gid;
                        const int32_t node = idx < nn ? pool[idx] : -1;
                        const uint32_t sm = row_sum(node);
                        if (sl == 0 && node >= 0) offer(sm, node);
                    }
                }
            }
            __syncthreads(); // B1: all survivors are queued
            const int ns = s_ns[par];
            JV_PHASE(4)
            expanded += nsel;
            visited += nn;

            // ---- (c) merge.  Phase 1: position of every survivor in the list (binary search); a survivor equal to a list
            //          entry is a re-scored member (evicted from the visited filter earlier) and is dropped.
            int my_pos0 = 0, my_pos1 = 0; // survivors tid and tid + 128 (E * R <= 256)
#pragma unroll 1
            for (int c = 0; c < 2; c++) {
                if (c * kQThreads >= ns) break; // (uniform) the second round exists only when a step queues > 128 survivors
                const int t = tid + c * kQThreads;
                if (t < ns) {
                    const uint64_t a = surv[t];
                    int lo = 0, hi = n;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if ((list[mid] >> 1) > a)
                            lo = mid + 1;
                        else
                            hi = mid;
                    }
                    if (lo < n && (list[lo] >> 1) == a) {
                        surv[t] = 0ull;
                        atomicAdd(&s_ndup[par], 1);
                    } else {
                        atomicAdd(&gapcnt[lo], 1); // lo list entries are better than this survivor
                    }
                    if (c == 0) my_pos0 = lo; else my_pos1 = lo;
                }
            }
            __syncthreads(); // B1.5: duplicates are zeroed
            JV_PHASE(1)
            if (tid == 0) { // next step's counters (pushes start after B2)
                s_ns[par ^ 1] = 0;
                s_nn[par ^ 1] = 0;
                s_ndup[par ^ 1] = 0;
            }
            const int ndup = s_ndup[par];
            //          Phase 2: position = own rank + number of better keys on the other side (survivors are distinct:
            //          atomic filter insertion).  Selected entries lose their "unexpanded" flag here.
            //          Two survivors can carry the SAME key: when the visited filter evicts an entry that was inserted in
            //          this very step, the node is scored twice.  Equal keys are ordered by queue index, so every element
            //          still gets its own slot (no holes); the copy is removed when the list is emitted.
            // survivors queued before this warp's 32 count when >= a (earlier queue index wins ties), the others when > a:
            // one loop, threshold a - 1 in front of t0
            auto count_better = [&](int t0, uint64_t a) -> int {
                int c0 = 0, c1 = 0, c2 = 0, c3 = 0, j = 0;
#pragma unroll 1
                for (; j + 4 <= ns; j += 4) {
                    const uint64_t thr = a - (j < t0 ? 1ull : 0ull); // t0 is a multiple of 32: the 4 entries are on one side
                    c0 += (surv[j] > thr) ? 1 : 0;
                    c1 += (surv[j + 1] > thr) ? 1 : 0;
                    c2 += (surv[j + 2] > thr) ? 1 : 0;
                    c3 += (surv[j + 3] > thr) ? 1 : 0;
                }
#pragma unroll 1
                for (; j < ns; j++) c0 += (surv[j] > a - (j < t0 ? 1ull : 0ull)) ? 1 : 0;
                return (c0 + c1) + (c2 + c3);
            };
#pragma unroll 1
            for (int c = 0; c < 2; c++) {
                if (c * kQThreads >= ns) break;
                const int t = tid + c * kQThreads;
                if (t < ns) {
                    const uint64_t a = surv[t];
                    // copies of my key among this warp's 32 survivors (one MATCH instruction); earlier queue index wins ties
                    const uint32_t same = __match_any_sync(__activemask(), a) & ((1u << lane) - 1u);
                    int ins_pos = 0x7fffffff;
                    bool ins_acc = false;
                    if (a != 0ull) {
                        const int t0 = t & ~31; // this warp's survivors start here
                        const int pos = (c == 0 ? my_pos0 : my_pos1) + count_better(t0, a) + __popc(same);
                        if (pos < Lc) {
                            out[pos] = (a << 1) | 1ull;
                            ins_pos = pos;
                            ins_acc = (a & 1ull) != 0ull; // bit 1 of the list key
                        }
                    }
                    if (FILT) { // bookkeeping of the selection's fast path: accepted entries gained, first unexpanded position
                        const uint32_t am = __activemask();
                        const int lo_pos = __reduce_min_sync(am, ins_pos);
                        const int gained = __popc(__ballot_sync(am, ins_acc));
                        if (lane == __ffs(am) - 1) {
                            if (gained) atomicAdd(&s_nacc, gained);
                            if (lo_pos != 0x7fffffff) atomicMin(&s_first, lo_pos);
                        }
                    }
                }
            }
            // list entry t moves down by the number of survivors better than it = survivors in gaps 0..t (inclusive prefix
            // sum of gapcnt): warp-level scan, chunks of 32 entries dealt to the warps from the top (survivors use the low ones)
            int offset = 0, done_cc = 0; // survivors in the gaps [0, done_cc): carried from chunk to chunk (linear in n)
            for (int c0 = (kQW - 1 - warp) * 32; c0 < n; c0 += kQW * 32) {
                for (int cc = done_cc; cc < c0; cc += 32) { // the chunks between this warp's previous chunk and this one
                    int v = gapcnt[cc + lane];
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(JV_FULL_MASK, v, o);
                    offset += v;
                }
                done_cc = c0;
                const int t = c0 + lane;
                int v = t < n ? gapcnt[t] : 0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int up = __shfl_up_sync(JV_FULL_MASK, v, o);
                    if (lane >= o) v += up;
                }
                uint32_t selmask = 0u; // the selected entries of this chunk (warp-uniform: nsel compares per chunk, not per entry)
                for (int e = 0; e < nsel; e++) {
                    const int d = w_pos[warp][e] - c0;
                    if (d >= 0 && d < 32) selmask |= 1u << d;
                }
                bool lost_acc = false;
                if (t < n) {
                    uint64_t k = list[t];
                    if ((selmask >> lane) & 1u) k &= ~1ull;
                    const int pos = t + offset + v;
                    if (pos < Lc)
                        out[pos] = k;
                    else
                        lost_acc = FILT && (k & 2ull) != 0ull; // an accepted entry falls off the end of the list
                }
                if (FILT) {
                    const int lost = __popc(__ballot_sync(JV_FULL_MASK, lost_acc));
                    if (lost && lane == 0) atomicSub(&s_nacc, lost);
                }
            }
            __syncthreads(); // B2
            JV_PHASE(5)
            for (int i = tid; i <= Lc; i += kQThreads) gapcnt[i] = 0; // consumed above; the next increments come after B1
            n = n + ns - ndup < Lc ? n + ns - ndup : Lc;
            cur ^= 1;
            step++;
        }

        // ---- emit the approximate result list, best first, in the (score, ~node) key format of the rerank step
        {
            const uint64_t *list = cur ? list1 : list0;
            uint64_t *akeys = cur ? list0 : list1; // fused rerank: converted keys go to the idle list buffer
            uint64_t *o = FUSE ? akeys : p.approx_keys + (int64_t)qi * L;
            auto key_of = [&](uint64_t k) -> uint64_t {
                const int32_t node = node_of(k);
                const uint32_t ord = (uint32_t)(k >> 32);
                const float sc = isum_keys ? score_of(l2 ? ~ord : ord, node) : jv_ord2f(ord);
                return jv_mk_key(sc, node);
            };
            int dup = 0; // a node scored twice in one step (see the merge) sits in two adjacent slots
            for (int i = tid + 1; i < n; i += kQThreads) dup |= ((list[i] >> 1) == (list[i - 1] >> 1)) ? 1 : 0;
            int cnt = n;
            if (FILT) { // the best L ACCEPTED entries, in order (serial: <= Lc iterations once per query)
                if (tid == 0) {
                    int w = 0;
                    for (int i = 0; i < n && w < L; i++) {
                        const uint64_t k = list[i];
                        if ((k & 2ull) && (i == 0 || (k >> 1) != (list[i - 1] >> 1))) o[w++] = key_of(k);
                    }
                    s_nn[0] = w; // broadcast slot (reset at the top of the next query)
                    for (; !FUSE && w < L; w++) o[w] = 0ull;
                }
                __syncthreads();
                cnt = s_nn[0];
            } else if (!__syncthreads_or(dup)) {
                for (int i = tid; i < (FUSE ? n : L); i += kQThreads) o[i] = i < n ? key_of(list[i]) : 0ull;
            } else { // rare: serial compaction
                if (tid == 0) {
                    int w = 0;
                    for (int i = 0; i < n; i++)
                        if (i == 0 || (list[i] >> 1) != (list[i - 1] >> 1)) o[w++] = key_of(list[i]);
                    s_nn[0] = w; // broadcast slot (reset at the top of the next query)
                    for (; !FUSE && w < L; w++) o[w] = 0ull;
                }
                __syncthreads();
                cnt = s_nn[0];
            }
            int reranked = 0;
            if (FUSE) {
                // ---- K3 as the epilogue: the table is no longer needed, its shared memory holds the fp32 query and the keys
                __syncthreads();
                float *sq = reinterpret_cast<float *>(smem_raw);
                uint64_t *rkeys = reinterpret_cast<uint64_t *>(smem_raw + ((((size_t)p.dim * 4) + 15) & ~(size_t)15));
                const bool vec4 = (p.dim & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.queries) & 15) == 0);
                reranked = rerank_query<kQThreads>(p.vectors, p.vec_norm, p.ord_to_doc, p.dim, p.sim, 1, p.queries + (int64_t)qi * p.dim, vec4,
                                                   p.fuse_k, cnt, p.rerank_floor, akeys, sq, rkeys, p.out_doc + (int64_t)qi * p.fuse_k,
                                                   p.out_score + (int64_t)qi * p.fuse_k, p.out_count + qi);
            }
            if (tid == 0) {
                p.approx_count[qi] = cnt;
                if (p.stats) {
                    jv_query_stats st;
                    st.visited = visited;
                    st.expanded = expanded;
                    st.expanded_base = expanded;
                    st.reranked = reranked;
                    p.stats[qi] = st;
                }
            }
        }
        JV_PHASE(6)
        if (PROF && tid == 0 && p.dbg) { // phases: 0 setup + table wait, 1 merge/dedupe, 2 select, 3 neighbour rows, 4 scoring, 5 merge/rank, 6 emit
            unsigned long long *ph = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(p.dbg) + 64);
            for (int i = 0; i < 7; i++) atomicAdd(ph + i, (unsigned long long)ck[i]);
            atomicAdd(ph + 7, (unsigned long long)step);
            for (int i = 8; i < 12; i++) atomicAdd(ph + i, (unsigned long long)ck[i]); // scoring sub-phases: code words arrive, lookups, offers
        }
#undef JV_PHASE
    }
}

template <int NJ_T, int W, bool PROF, bool FILT = false, bool FUSE = false>
static int32_t launch_q8_typed(jv_index *ix, SearchCtx *ctx, Q8Params &p) {
    constexpr int kQThreads = W * 32;
    auto kern = q8_search_kernel<NJ_T, W, PROF, FILT, FUSE>;
    const int Lc = FILT ? p.Lc : p.L;
    const size_t fixed = (size_t)p.lutb + (size_t)Lc * 16 + (size_t)p.surv_cap * 12 + (size_t)((Lc + 2) & ~1) * 4;
    const size_t sm_total = 228 * 1024;
    int64_t want = (int64_t)p.L * p.R; // words; 2 tags each
    if (want < 1024) want = 1024;
    const int64_t min_words = FILT ? 4096 : 1024; // a filter multiplies the visited nodes by ~1/selectivity
    if (FILT) want = 8192;
    int best_occ = 0, best_log2 = 0;
    for (int occ = 8; occ >= 1; occ--) {
        const int64_t per = (int64_t)(sm_total / occ) - 1024 - 640 - (int64_t)fixed; // 1 KB system + static __shared__
        if (per < min_words * 4) continue;
        int lg = FILT ? 12 : 10;
        while (lg < 15 && ((int64_t)4 << (lg + 1)) <= per && ((int64_t)1 << lg) < want) lg++;
        best_occ = occ;
        best_log2 = lg;
        break;
    }
    if (!best_occ) {
        set_error("search (8-bit table): shared memory budget exceeded (%zu fixed bytes)", fixed);
        return JV_ERR_UNSUPPORTED;
    }
    p.hash_log2 = best_log2;
    const size_t smem = fixed + ((size_t)4 << best_log2);
    JV_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    JV_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kQThreads, smem));
    if (occ < 1) {
        set_error("search (8-bit table): kernel does not fit on an SM (smem %zu)", smem);
        return JV_ERR_UNSUPPORTED;
    }
    if (q8_knobs().occ >= 1 && q8_knobs().occ < occ) occ = q8_knobs().occ; // diagnostics: cap the CTAs per SM
    int grid = ix->sm_count * occ;
    if (grid > p.nq) grid = p.nq;
    JV_CUDA_TRY(cudaMemsetAsync(p.work_counter, 0, sizeof(int), ctx->stream));
    kern<<<grid, kQThreads, smem, ctx->stream>>>(p);
    JV_CUDA_TRY(cudaGetLastError());
    return JV_OK;
}

// list capacity with a filter: rejected nodes stay in the list (they are traversed) next to the rerankK accepted ones
static int q8_filtered_list_cap(int L) { return 8 * L; }

bool q8_search_supported(const jv_index *ix, int L, int R, bool filtered) {
    if (!ix->has_pq || !ix->q8_ok || R > 64) return false; // <= 2 queued survivors per thread and step
    if (filtered && (q8_filtered_list_cap(L) > 1024 || ix->n >= (1ll << 30))) return false; // strict kernel instead
    const int Lc = filtered ? q8_filtered_list_cap(L) : L;
    const size_t fixed = (size_t)q8_lut_bytes(ix->q8_nj) + (size_t)Lc * 16 + (size_t)kQMaxE * ((R + 31) / 32) * 32 * 12 + (size_t)((Lc + 2) & ~1) * 4;
    return fixed + (filtered ? 16384 : 4096) + 2048 <= 227 * 1024;
}

// LUT build + traversal for queries [0, nq) in chunks bounded by the table staging buffer
int32_t launch_search_q8(jv_index *ix, SearchCtx *ctx, const SearchLaunch &a, int *launches, bool *reranked) {
    const int lutb = q8_lut_bytes(ix->q8_nj);
    // staging buffer: <= 512 MB of tables per chunk (10 922 queries at M = 192)
    int chunk = (int)((size_t)512 * 1024 * 1024 / (size_t)lutb);
    if (q8_knobs().chunk > 0 && q8_knobs().chunk < chunk) chunk = q8_knobs().chunk; // test knob: force small chunks
    if (chunk > a.nq) chunk = a.nq;
    if (chunk < 1) chunk = 1;
    JV_TRY(ctx->lut8.ensure((size_t)chunk * lutb));
    JV_TRY(ctx->qparams.ensure((size_t)chunk * sizeof(float4)));
    JV_TRY(ctx->counter.ensure(sizeof(int)));
    // default width: 4 (measured best without a filter); filtered queries need ~1/selectivity more expansions anyway, so a
    // wider step costs few wasted visits and saves steps (measured at 10 % selectivity: 2/4/6/8 -> 265k/453k/514k/460k queries/s;
    // at 50 %: 4 -> 1.13 M, 8 -> 1.08 M)
    // the manager / expander / scorer kernel (M = 161..192, no filter, list <= 64) works two steps deep; measured at cfg2: 3 candidates
    // per step -> 1.19 ms / 1 078 visited / recall 0.9947, 4 -> 1.21 ms / 1 218 visited / recall 0.9962: width 4 everywhere
    const int dflt = a.d_accept != nullptr ? 6 : 4;
    int E = a.expand_width <= 0 ? dflt : (a.expand_width > kQMaxE ? kQMaxE : a.expand_width);
    for (int q0 = 0; q0 < a.nq; q0 += chunk) {
        const int nqc = a.nq - q0 < chunk ? a.nq - q0 : chunk;
        JV_TRY(launch_lut_q8(ix, ctx->stream, a.d_queries + (int64_t)q0 * ix->dim, nqc, ctx->lut8.as<uint8_t>(), ctx->qparams.as<float4>()));
        if (q0 == 0 && ctx->time_lut) {
            JV_CUDA_TRY(cudaEventRecord(ctx->ev[5], ctx->stream));
            ctx->lut_timed = true;
        }
        Q8Params p;
        memset(&p, 0, sizeof(p));
        const bool filt = a.d_accept != nullptr;
        p.adjacency = ix->adjacency.as<int32_t>();
        p.codes_q8 = ix->codes_q8.as<uint8_t>();
        p.node_norm = ix->node_norm.as<float>();
        p.lut = ctx->lut8.as<uint8_t>();
        p.qparams = ctx->qparams.as<float4>();
        p.approx_keys = a.d_approx_keys + (int64_t)q0 * a.rerank_k;
        p.approx_count = a.d_approx_count + q0;
        p.stats = a.d_stats ? a.d_stats + q0 : nullptr;
        p.work_counter = ctx->counter.as<int>();
        p.dbg = ix->dbg.as<int>();
        p.n = ix->n;
        p.nq = nqc;
        p.L = a.rerank_k;
        p.R = ix->R;
        p.entry = a.entry_override >= 0 ? a.entry_override : ix->entry;
        p.sim = ix->sim;
        p.NJ = ix->q8_nj;
        p.lutb = lutb;
        p.E = E;
        p.dim = ix->dim;
        p.ord_to_doc = ix->ord_to_doc.as<int32_t>();
        p.Lc = filt ? q8_filtered_list_cap(a.rerank_k) : a.rerank_k;
        p.accept = a.d_accept;
        p.accept_stride = a.accept_stride_words; // words per query (0 = one bitset for the batch)
        if (filt && a.accept_stride_words) p.accept += (int64_t)q0 * a.accept_stride_words;
        // K3 can run as the traversal's epilogue (JVGPU_Q8_FUSED=1).  Measured at cfg2: 2.67 ms fused vs 2.34 ms with the
        // separate rerank kernel — a 4-warp CTA gathers its 50 rows one DRAM round trip after the other while it holds
        // 1/4 of an SM; the stand-alone K3 keeps 16 CTAs per SM in flight (0.31 ms, 0.76 of the HBM roofline) — so off by default.
        const bool fuse = a.fuse_k > 0 && !ix->vectors_on_host && !ix->has_nvq && (size_t)ix->dim * 4 + 16 + (size_t)a.rerank_k * 8 <= (size_t)lutb &&
                          ix->q8_nj == 6 && !filt && q8_knobs().fused; // instantiated for the headline shape only
        if (fuse) {
            p.fuse_k = a.fuse_k;
            p.rerank_floor = a.rerank_floor;
            p.queries = a.d_queries + (int64_t)q0 * ix->dim;
            p.vectors = ix->vectors_dev;
            p.vec_norm = ix->vec_norm.as<float>();
            p.ord_to_doc = ix->ord_to_doc.as<int32_t>();
            p.out_doc = a.d_out_doc + (int64_t)q0 * a.fuse_k;
            p.out_score = a.d_out_score + (int64_t)q0 * a.fuse_k;
            p.out_count = a.d_out_count + q0;
        }
        if (reranked) *reranked = fuse;
        p.surv_cap = E * ((ix->R + 31) / 32) * 32;
        int32_t st;
        const bool prof = q8_knobs().prof; // per-phase cycle counters (costs registers): diagnostics only
        const int warps = q8_knobs().warps == 8 ? 8 : 4;
        // production: manager / scorer kernel (jv_q8_beam.cu); the round-synchronous kernel below keeps filtered queries, lists
        // longer than 64 entries and steps wider than 4 adjacency chunks (JVGPU_Q8_SYNC=1 forces it: diagnostics)
        if (!q8_knobs().sync && !filt && !fuse && q8_beam_supported(ix, a.rerank_k, ix->R, E)) {
            JV_TRY(launch_q8_beam(ix, ctx, p));
            if (launches) *launches += 2;
            ctx->last_width = E;
            ctx->last_kernel = JV_KERNEL_Q8_BEAM;
            continue;
        }
        // the diagnostic instantiations (8 warps, phase counters) exist for the headline shape (M = 192) only
#define JV_Q8_CASE(NJV)                                                                         \
    case NJV:                                                                                   \
        st = filt ? launch_q8_typed<NJV, 4, false, true>(ix, ctx, p) : launch_q8_typed<NJV, 4, false>(ix, ctx, p); \
        break;
        switch (ix->q8_nj) {
            JV_Q8_CASE(1)
            JV_Q8_CASE(2)
            JV_Q8_CASE(3)
            JV_Q8_CASE(4)
            JV_Q8_CASE(8)
        case 6:
            if (filt)
                st = launch_q8_typed<6, 4, false, true>(ix, ctx, p);
            else
                st = fuse         ? launch_q8_typed<6, 4, false, false, true>(ix, ctx, p)
                     : warps == 8 ? (prof ? launch_q8_typed<6, 8, true>(ix, ctx, p) : launch_q8_typed<6, 8, false>(ix, ctx, p))
                                  : (prof ? launch_q8_typed<6, 4, true>(ix, ctx, p) : launch_q8_typed<6, 4, false>(ix, ctx, p));
            break;
        default:
            st = filt ? launch_q8_typed<0, 4, false, true>(ix, ctx, p) : launch_q8_typed<0, 4, false>(ix, ctx, p);
            break;
        }
#undef JV_Q8_CASE
        JV_TRY(st);
        if (launches) *launches += 2;
        ctx->last_width = E;
        ctx->last_kernel = JV_KERNEL_Q8_SYNC;
    }
    return JV_OK;
}

}  // namespace jv
