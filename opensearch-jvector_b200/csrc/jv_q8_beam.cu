// jv_q8_beam.cu — K2, production kernel of the 8-bit table path: manager warp + expander warp + scorer warps, software-pipelined steps.
//
// Reference loop: GraphSearcher.search over the PQ score function (JVectorReader.java:165-173).
//
// One CTA owns one query; its 8-bit ADC table is staged in shared memory with one TMA bulk copy (like the round-synchronous
// kernel of jv_q8.cu, which stays the path of filtered queries and of lists longer than 64).  The warps are specialised:
//
//   manager warp       owns the list.  The sorted list of the best L visited nodes (L <= 64) lives in its REGISTERS, two keys per
//                      lane (shared memory holds the copy the merge gathers and scatters through); selection and the list merge
//                      are warp-synchronous — no block barrier, no shared-memory atomics on the list, nothing replicated across
//                      warps.  Per step: pick the E best unexpanded entries and hand them to the expander.  Merge, without a loop
//                      over the list: one survivor per lane; its slot = list entries better than it (binary search) + survivors
//                      of the step better than it (counting over 16-byte broadcast reads of the queue); the list entries fill
//                      the remaining slots in order — an occupancy mask of the survivor slots (two REDUX.OR) tells every output
//                      slot which old entry it takes; one gather, one scatter, reload.
//   expander warp      owns the visited filter (plain loads and stores, no atomics).  Per step: read the adjacency rows of the
//                      selected entries (one coalesced 128-byte load each, all in flight together), test-and-insert the filter,
//                      start the fresh code rows towards L2, write the fresh ids to the pool, hand the pool to the scorers.  With
//                      the manager merging step s while the expander prepares step s+1, neither the adjacency round trip nor the
//                      filter chain is on the manager's critical path.
//   scorer warps       score pools: 8 lanes per code row, all 16-byte code loads of a group's rows in flight before the first
//                      table lookup, bank-conflict-free lookups at 3.25 instructions each (layout in jv_q8.cu); sums that beat the
//                      list's worst entry — read from the manager's live copy — are queued (one shared-memory atomic per warp
//                      and pass).
//
// Hand-over is three named barriers per step (bar.arrive / bar.sync, ids by step parity), so a waiting warp costs no issue
// slots; the barrier itself orders the producer's shared-memory writes before the consumer's reads (PTX ISA, bar: producer /
// consumer example), no fence instruction on the chain.  With depth 2 the manager selects and issues step s+1 BEFORE it merges
// the scores of step s: selection, expansion, scoring and merge of neighbouring steps overlap.  The selection then lags one step
// behind the scores — the same relaxation as a wider step — and the result does not depend on timing (survivors are ranked by
// key, whatever their queue order).  expand_width = 1 runs at depth 1: exactly the best-first order of the oracle's 8-bit mode.
//
// Measured alternatives: DESIGN.md section 6 (items 1-17).
#include "jv_q8.cuh"

namespace jv {

__device__ __forceinline__ void nbar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
    const uint32_t lo = __shfl_sync(JV_FULL_MASK, (uint32_t)v, src), hi = __shfl_sync(JV_FULL_MASK, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}

constexpr int kBeamPool = 128; // fresh neighbours per step: E * ceil(R / 32) <= 4 adjacency chunks of 32
constexpr int kBeamList = 64;  // list capacity (two keys per manager lane)
constexpr int kBarPool = 1;    // +parity: expander arrives, scorers sync (the pool of a step is complete)
constexpr int kBarDone = 3;    // +parity: scorers arrive, manager syncs (the survivors of a step are queued)
constexpr int kBarSel = 5;     // +parity: manager arrives, expander syncs (the selection of a step is published)

// PROF: cycles of the manager (0 select, 3 wait for the scorers + survivors in registers, 11 binary search + rank count, 12 occupancy
// mask + gather, 13 scatter + reload, 6 rest of the merge, 4 query setup, 5 emit), of the expander (2 wait for a selection, 1 adjacency + filter + pool) and of scorer warp 0 (8 wait for a pool,
// 9 code words in registers, 10 lookups + queue); 7 = steps, 14 = survivors, 15 = merge rounds that had to drop a copy
template <int NJ, int SW, bool PROF>
__global__ void __launch_bounds__((SW + 2) * 32, 4) q8_beam_kernel(const Q8Params p, const int depth) {
    constexpr int kT = (SW + 2) * 32, NG = SW * 4, kTS = SW * 32; // threads per CTA / row groups / scorer threads
    constexpr int U = SW >= 7 ? 2 : SW == 3 ? 4 : 3; // rows in flight per row group and pass (NG * U >= 48: one pass per typical step)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = p.L, E = p.E, H = 1 << p.hash_log2, R = p.R;
    const uint32_t lt = (1u << lane) - 1u;
    // roles by warp index: scorers 0 .. SW-1, expander SW, manager SW+1.  The SM's warp arbiter prefers the highest warp id among
    // eligible warps: the manager's dependent chain is the critical path and gets the slot whenever it is ready, the scorers'
    // lookups fill the rest.
    const int sw = warp; // scorer index

    unsigned char *sp = smem_raw;
    const uint8_t *lut = sp;
    sp += p.lutb;
    uint64_t *lm = reinterpret_cast<uint64_t *>(sp); // the list as the merge scatters it (registers are the working copy)
    sp += kBeamList * 8;
    uint64_t *survq = reinterpret_cast<uint64_t *>(sp); // [2][kBeamPool] keys >> 1 of the scored nodes that beat the published threshold
    sp += 2 * kBeamPool * 8;
    int32_t *pool = reinterpret_cast<int32_t *>(sp); // [2][kBeamPool] fresh neighbour ids by step parity
    sp += 2 * kBeamPool * 4;
    uint32_t *filter = reinterpret_cast<uint32_t *>(sp);

    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_query, s_nn[2], s_ns[2], s_nsel[2], s_ru[2], w_sel[2][2 * kQMaxE];

    const bool tagged = p.n <= ((int64_t)1 << (p.hash_log2 + 15));
    const bool isum_keys = p.sim != JV_SIM_COSINE;
    const bool l2 = p.sim == JV_SIM_EUCLIDEAN;
    const int g = lane >> 3, sl = lane & 7;
    // table offset of code c in bank column b = 4 * (lane's bank): ((c & 63) << 8) | b | (c >> 6).  Per code word (4 codes): one mask
    // (c & 63 in every byte) and one shift-mask-or (c >> 6 and the lane's bank bits in every byte: the byte used at slot i carries the
    // bank quarter (i + g) & 3), then ONE byte permute per lookup assembles the clean offset: 3.25 instructions per lookup
    uint32_t sel[4], lbw = 0u;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t qd = (uint32_t)(i + g) & 3u;
        // byte 0 <- hi.byte[qd], byte 1 <- lo.byte[qd], bytes 2 and 3 <- the replicated sign bit of lo.byte[0] (always 0: c & 63)
        sel[i] = 0x8800u | (qd << 4) | (4u + qd);
        lbw |= (qd * 32u + (uint32_t)sl * 4u) << (8u * qd);
    }
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

    if (tid == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t phase = 0;

    for (;;) {
        __syncthreads(); // everyone is done with the previous query's table and pools
        if (tid == 0) {
            s_query = atomicAdd(p.work_counter, 1);
            s_ns[0] = 0;
            s_ns[1] = 0;
        }
        __syncthreads();
        const int qi = s_query;
        if (qi >= p.nq) break;
        long long ck[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        long long t_prev = PROF ? clock64() : 0;
#define JV_PHASE(i)                        \
    if (PROF) {                            \
        const long long t_now = clock64(); \
        ck[i] += t_now - t_prev;           \
        t_prev = t_now;                    \
    }
        if (tid == 0) { // K1 result: one TMA bulk copy HBM/L2 -> shared memory
            mbar_expect_tx(&s_bar, (uint32_t)p.lutb);
            bulk_g2s(smem_raw, p.lut + (int64_t)qi * p.lutb, (uint32_t)p.lutb, &s_bar);
        }
        for (int i = tid; i < H; i += kT) filter[i] = tagged ? 0u : kEmpty;
        if (tid < kBeamList) lm[tid] = 0ull;
        for (int i = tid; i < 2 * kBeamPool; i += kT) survq[i] = 0ull;
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
        __syncthreads(); // filter and list cleared
        mbar_wait(&s_bar, phase);
        phase ^= 1u;

        if (warp < SW) {
            // ================================================================================================ scorers
            const int gid = sw * 4 + g;
            for (int step = 0;; step++) {
                const int par = step & 1;
                nbar_sync(kBarPool + par, kTS + 32); // the expander's pool of this step is complete
                const int nn = s_nn[par];
                if (nn < 0) break; // query finished
                // admission threshold: the L-th list entry (0 until the list is full), read from the manager's live copy — it only
                // improves, the manager checks every survivor against the current list again, and whatever is admitted in between
                // is dropped there, so the result does not depend on when this read happens
                const uint64_t worst = lm[L - 1] >> 1;
                if (sw == 0) JV_PHASE(8)
                const int32_t *pl = pool + par * kBeamPool;
                uint64_t *sq = survq + par * kBeamPool;
                auto pass = [&](int i0, auto nu_tag) {
                    constexpr int NU = decltype(nu_tag)::value;
                    uint32_t cw[NU][NJ], s[NU];
                    int32_t nbv[NU];
#pragma unroll
                    for (int u = 0; u < NU; u++) {
                        const int idx = i0 + u * NG + gid;
                        nbv[u] = idx < nn ? pl[idx] : -1;
                        s[u] = 0u;
                        if (nbv[u] >= 0) {
                            q8_load_row<NJ>(p.codes_q8 + (int64_t)nbv[u] * (NJ * 32), sl, cw[u]); // 16-byte vector loads
                        } else {
#pragma unroll
                            for (int j = 0; j < NJ; j++) cw[u][j] = 0u;
                        }
                    }
                    if (PROF && sw == 0) {
                        uint32_t acc = 0;
#pragma unroll
                        for (int u = 0; u < NU; u++)
#pragma unroll
                            for (int j = 0; j < NJ; j++) acc |= cw[u][j];
                        asm volatile("" ::"r"(acc));
                        JV_PHASE(9)
                    }
#pragma unroll
                    for (int j = 0; j < NJ; j++) {
#pragma unroll
                        for (int u = 0; u < NU; u++) s[u] += lookup4(cw[u][j], j); // a group without a row looks up code 0: harmless
                    }
#pragma unroll
                    for (int u = 0; u < NU; u++) s[u] = reduce8(s[u]);
                    // after the butterfly every lane of a group holds the group's sums: lane sl = u takes row u, so one key, one
                    // ballot and one shared-memory atomic serve all NU rows of the 4 groups
                    uint32_t my_s = s[0];
                    int32_t my_nb = nbv[0];
#pragma unroll
                    for (int u = 1; u < NU; u++) {
                        my_s = sl == u ? s[u] : my_s;
                        my_nb = sl == u ? nbv[u] : my_nb;
                    }
                    const uint64_t ka = (sl < NU && my_nb >= 0) ? (qkey_pack(ord_of(my_s, my_nb), my_nb) >> 1) : 0ull;
                    const uint32_t bal = __ballot_sync(JV_FULL_MASK, ka > worst);
                    if (bal) { // warp-uniform
                        int slot = 0;
                        if (lane == 0) slot = atomicAdd(&s_ns[par], __popc(bal));
                        slot = __shfl_sync(JV_FULL_MASK, slot, 0);
                        if ((bal >> lane) & 1u) sq[slot + __popc(bal & lt)] = ka;
                    }
                };
                for (int i0 = 0; i0 < nn; i0 += NG * U) {
                    const int left = nn - i0 - sw * 4; // rows of this pass at or after this warp's first group (warp-uniform)
                    if (U >= 4 && left > 3 * NG)
                        pass(i0, std::integral_constant<int, U >= 4 ? 4 : 1>());
                    else if (U >= 3 && left > 2 * NG)
                        pass(i0, std::integral_constant<int, U >= 3 ? 3 : 1>());
                    else if (left > NG)
                        pass(i0, std::integral_constant<int, 2>());
                    else if (left > 0)
                        pass(i0, std::integral_constant<int, 1>());
                }
                nbar_arrive(kBarDone + par, kTS + 32); // the survivors of this step are queued
                if (sw == 0) JV_PHASE(10)
            }
            if (PROF && sw == 0 && lane == 0 && p.dbg) {
                unsigned long long *ph = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(p.dbg) + 64);
                for (int i = 8; i < 11; i++) atomicAdd(ph + i, (unsigned long long)ck[i]);
            }
            continue;
        }

        const int RC = (R + 31) >> 5; // adjacency chunks of 32 per row (R <= 64 here)
        if (warp == SW) {
            // =============================================================================================== expander
            const int lines = (R * 4 + 127) >> 7;
            for (int step = 0;; step++) {
                const int par = step & 1;
                nbar_sync(kBarSel + par, 64); // the manager's selection of this step is published
                const int nsel = s_nsel[par];
                if (nsel < 0) { // query finished: the scorers leave their loop
                    if (lane == 0) s_nn[par] = -1;
                    nbar_arrive(kBarPool + par, kTS + 32);
                    break;
                }
                JV_PHASE(2)
                const int *ws = w_sel[par];
                const int total = nsel * RC; // <= 4
                int32_t nb[4];
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    nb[t] = -1;
                    if (t < total) {
                        const int e = RC == 1 ? t : (t >> 1), r = (RC == 1 ? 0 : (t & 1) * 32) + lane;
                        if (r < R) nb[t] = __ldg(p.adjacency + (int64_t)ws[e] * R + r);
                    }
                }
                { // runner-up rows -> L2 (lines per row: 1 or 2)
                    const int ru = s_ru[par];
                    const int e = lines == 1 ? lane : (lane >> 1);
                    if (e < ru)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(p.adjacency + (int64_t)ws[nsel + e] * R) +
                                                                      (lines == 1 ? 0 : (lane & 1) * 128)));
                }
                int32_t *pl = pool + par * kBeamPool;
                // visited filter (the expander is its only owner: plain loads and stores).  Hashes of all chunks first
                // (independent), then one short read-test-write per chunk — a chunk must see the entries of the chunks before
                // it (the rows of one step share many neighbours).  Two lanes of one chunk that map to the same set can
                // overwrite each other's tag: the loser may be scored again later and is then dropped by the merge (the result
                // list does not depend on which one survives; the visited counter can differ by a few re-scored nodes).
                uint32_t fset[4], ftag[4];
                bool fresh[4];
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const bool ok = nb[t] >= 0 && nb[t] < p.n;
                    fresh[t] = ok;
                    if (tagged) {
                        const uint32_t x = ((uint32_t)nb[t] * 0x9E3779B1u) & ((1u << (p.hash_log2 + 15)) - 1u);
                        fset[t] = x >> 15;
                        ftag[t] = (x & 0x7fffu) | 0x8000u;
                    } else {
                        fset[t] = ((uint32_t)nb[t] * 2654435761u) >> (32 - p.hash_log2);
                        ftag[t] = (uint32_t)nb[t];
                    }
                }
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    if (t < total) { // warp-uniform
                        if (fresh[t]) {
                            const uint32_t old = filter[fset[t]];
                            if (tagged) {
                                if ((old & 0xffffu) == ftag[t] || (old >> 16) == ftag[t])
                                    fresh[t] = false;
                                else
                                    filter[fset[t]] = (old << 16) | ftag[t];
                            } else {
                                fresh[t] = old != ftag[t];
                                filter[fset[t]] = ftag[t];
                            }
                        }
                        if (fresh[t]) {
                            // the scorers read the code row after the barrier: start DRAM -> L2 as early as possible (rows are
                            // 32-byte aligned and <= 256 bytes: the lines of the first and of the last byte cover them)
                            const char *row = reinterpret_cast<const char *>(p.codes_q8 + (int64_t)nb[t] * (NJ * 32));
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
                            if (NJ * 32 > 128 || (reinterpret_cast<uintptr_t>(row) & 127) + NJ * 32 > 128)
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(row + NJ * 32 - 1));
                        }
                        __syncwarp();
                    }
                }
                int nn = 0;
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    if (t < total) {
                        const uint32_t bal = __ballot_sync(JV_FULL_MASK, fresh[t]);
                        if (fresh[t]) pl[nn + __popc(bal & lt)] = nb[t];
                        nn += __popc(bal);
                    }
                }
                if (lane == 0) s_nn[par] = nn; // an empty pool is a step like any other: the scorers arrive at once
                nbar_arrive(kBarPool + par, kTS + 32);
                JV_PHASE(1)
            }
            if (PROF && lane == 0 && p.dbg) {
                unsigned long long *ph = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(p.dbg) + 64);
                for (int i = 1; i < 3; i++) atomicAdd(ph + i, (unsigned long long)ck[i]);
            }
            continue;
        }

        // ==================================================================================================== manager
        uint64_t k0 = 0ull, k1 = 0ull; // list entries lane and lane + 32, best first; bit 0 = unexpanded; 0 = empty
        int n = 0, visited = 0, expanded = 0;
        if (p.entry >= 0 && p.entry < p.n) {
            uint32_t s = 0;
            if (g == 0) {
                const unsigned char *row = p.codes_q8 + (int64_t)p.entry * (NJ * 32);
#pragma unroll
                for (int j = 0; j < NJ; j++) s += lookup4(__ldg(reinterpret_cast<const uint32_t *>(row + q8_word_offset(NJ, sl, j))), j);
            }
            s = reduce8(s);
            if (lane == 0) {
                k0 = qkey_pack(ord_of(s, p.entry), p.entry);
                lm[0] = k0;
                if (tagged) { // the entry node is visited (the expander reads the filter after the first selection barrier)
                    const uint32_t x = ((uint32_t)p.entry * 0x9E3779B1u) & ((1u << (p.hash_log2 + 15)) - 1u);
                    filter[x >> 15] = (x & 0x7fffu) | 0x8000u;
                } else {
                    filter[((uint32_t)p.entry * 2654435761u) >> (32 - p.hash_log2)] = (uint32_t)p.entry;
                }
            }
            n = 1;
            visited = 1;
        }
        __syncwarp();
        JV_PHASE(4)

        int issued = 0, merged = 0; // steps handed to the expander / merged back
        for (;;) {
            bool can_issue = issued - merged < depth;
            if (can_issue) {
                // ---- select the E best unexpanded entries (runners-up E..2E-1: adjacency rows prefetched into L2 by the expander)
                const bool un0 = lane < n && (k0 & 1ull), un1 = lane + 32 < n && (k1 & 1ull);
                const uint32_t b0 = __ballot_sync(JV_FULL_MASK, un0), b1 = __ballot_sync(JV_FULL_MASK, un1);
                const int c0 = __popc(b0), found = c0 + __popc(b1);
                const int nsel = found < E ? found : E;
                if (nsel == 0) {
                    if (issued == merged) break; // nothing to expand, nothing in flight: done
                    can_issue = false;           // the scores in flight may bring new candidates
                } else {
                    const int par = issued & 1;
                    int *ws = w_sel[par];
                    const int r0 = __popc(b0 & lt), r1 = c0 + __popc(b1 & lt);
                    if (un0 && r0 < 2 * E) ws[r0] = qkey_node(k0);
                    if (un1 && r1 < 2 * E) ws[r1] = qkey_node(k1);
                    if (un0 && r0 < nsel) lm[lane] = (k0 &= ~1ull); // the merge gathers from the shared-memory copy
                    if (un1 && r1 < nsel) lm[lane + 32] = (k1 &= ~1ull);
                    if (lane == 0) {
                        s_nsel[par] = nsel;
                        s_ru[par] = found - nsel < nsel ? found - nsel : nsel;
                    }
                    nbar_arrive(kBarSel + par, 64);
                    issued++;
                    expanded += nsel;
                    JV_PHASE(0)
                    continue;
                }
            }
            // ---- merge the oldest step in flight
            const int par = merged & 1;
            nbar_sync(kBarDone + par, kTS + 32);
            const int ns = s_ns[par];
            visited += s_nn[par];
            uint64_t *sr0 = survq + par * kBeamPool;
            uint64_t a_first = lane < ns ? sr0[lane] : 0ull;
            if (PROF) {
                asm volatile("" ::"l"(a_first));
                JV_PHASE(3)
                ck[14] += ns;
            }
            // rounds of <= 32 survivors (more than one round only while the list is still filling up).  A survivor's slot is
            // (list entries better than it: binary search) + (survivors of the round better than it: counting over broadcast
            // reads); the list entries fill the remaining slots in order (occupancy mask + gather) — no loop over the list.
            for (int r0 = 0; r0 < ns; r0 += 32) {
                const int cnt = ns - r0 < 32 ? ns - r0 : 32;
                uint64_t *sr = sr0 + r0;
                uint64_t a = r0 == 0 ? a_first : (lane < cnt ? sr[lane] : 0ull);
                { // the list may have improved since the scorers read the threshold: such survivors are worse than every survivor
                  // that stays, so they do not disturb the ranks
                    const uint64_t worst = lm[L - 1] >> 1;
                    if (a <= worst) a = 0ull;
                }
                const bool live = a != 0ull;
                const uint64_t A = (a << 1) | 1ull;
                // binary search: list entries better than the survivor; rank count: survivors of the round better than it (the
                // queue is zero beyond its end: 16-byte reads, no tail)
                int lo = 0, cs = 0;
                if (live) {
#pragma unroll
                    for (int st = 32; st >= 1; st >>= 1)
                        if ((lm[lo + st - 1] | 1ull) > A) lo += st;
                }
                {
                    const ulonglong2 *sr2 = reinterpret_cast<const ulonglong2 *>(sr);
#pragma unroll 2
                    for (int j = 0; j < cnt; j += 4) {
                        const ulonglong2 u = sr2[j >> 1], v = sr2[(j >> 1) + 1];
                        cs += (u.x > a ? 1 : 0) + (u.y > a ? 1 : 0) + (v.x > a ? 1 : 0) + (v.y > a ? 1 : 0);
                    }
                }
                bool sv = live && (lm[lo] | 1ull) != A; // equal: a re-scored list member (evicted from the visited filter earlier)
                uint32_t mask = __ballot_sync(JV_FULL_MASK, sv);
                JV_PHASE(11)
                if (!mask) {
                    __syncwarp(); // every lane is done reading the queue
                    if (lane < cnt) sr[lane] = 0ull;
                    continue;
                }
                // rare: a dropped list member was counted by the survivors behind it, or the same node was scored twice in this
                // step (lost or evicted filter entry: equal keys, equal ranks) — drop the copies and count again
                if (__ballot_sync(JV_FULL_MASK, live && !sv) != 0u || __popc(__reduce_or_sync(JV_FULL_MASK, sv ? 1u << cs : 0u)) != __popc(mask)) {
                    if (PROF) ck[15] += 1;
                    if (sv) {
                        const uint32_t same = __match_any_sync(mask, a);
                        if ((same & (0u - same)) != (1u << lane)) sv = false;
                    }
                    mask = __ballot_sync(JV_FULL_MASK, sv);
                    __syncwarp(); // every lane is done reading the queue
                    if (lane < cnt) sr[lane] = sv ? a : 0ull;
                    __syncwarp();
                    cs = 0;
                    for (int j = 0; j < cnt; j++) cs += sr[j] > a ? 1 : 0;
                }
                const int pos = lo + cs;
                const bool in = sv && pos < L;
                const uint32_t occ0 = __reduce_or_sync(JV_FULL_MASK, in && pos < 32 ? 1u << pos : 0u);
                const uint32_t occ1 = __reduce_or_sync(JV_FULL_MASK, in && pos >= 32 ? 1u << (pos - 32) : 0u);
                const bool s0 = (occ0 >> lane) & 1u, s1 = (occ1 >> lane) & 1u; // my two slots are taken by survivors
                const int i0 = lane - __popc(occ0 & lt), i1 = lane + 32 - __popc(occ0) - __popc(occ1 & lt);
                const uint64_t v0 = (!s0 && i0 < n && lane < L) ? lm[i0] : 0ull;
                const uint64_t v1 = (!s1 && i1 < n && lane + 32 < L) ? lm[i1] : 0ull;
                __syncwarp();
                JV_PHASE(12)
                if (in) lm[pos] = A;
                if (!s0 && i0 != lane) lm[lane] = v0;
                if (!s1 && i1 != lane + 32) lm[lane + 32] = v1;
                __syncwarp();
                if (lane < cnt) sr[lane] = 0ull; // the queue stays zero beyond its end
                n += __popc(mask);
                n = n < L ? n : L;
                k0 = s0 ? lm[lane] : v0;
                k1 = s1 ? lm[lane + 32] : v1;
                if (PROF) {
                    asm volatile("" ::"l"(k0 | k1));
                    JV_PHASE(13)
                }
            }
            if (lane == 0) s_ns[par] = 0; // the scorers queue into this buffer again two steps from now (after the next pool barrier)
            merged++;
            JV_PHASE(6)
        }
        // ---- the expander and (through it) the scorers leave their loops
        {
            const int par = issued & 1;
            if (lane == 0) s_nsel[par] = -1;
            nbar_arrive(kBarSel + par, 64);
        }
        // ---- emit the approximate result list, best first, in the (score, ~node) key format of the rerank step
        {
            uint64_t *o = p.approx_keys + (int64_t)qi * L;
            auto key_of = [&](uint64_t k) -> uint64_t {
                const int32_t node = qkey_node(k);
                const uint32_t ord = (uint32_t)(k >> 32);
                const float sc = isum_keys ? score_of(l2 ? ~ord : ord, node) : jv_ord2f(ord);
                return jv_mk_key(sc, node);
            };
            if (lane < L) o[lane] = lane < n ? key_of(k0) : 0ull;
            if (lane + 32 < L) o[lane + 32] = lane + 32 < n ? key_of(k1) : 0ull;
            if (lane == 0) {
                p.approx_count[qi] = n;
                if (p.stats) {
                    jv_query_stats st;
                    st.visited = visited;
                    st.expanded = expanded;
                    st.expanded_base = expanded;
                    st.reranked = 0;
                    p.stats[qi] = st;
                }
            }
        }
        JV_PHASE(5)
        if (PROF && lane == 0 && p.dbg) {
            unsigned long long *ph = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(p.dbg) + 64);
            for (int i = 0; i < 7; i++)
                if (i != 1 && i != 2) atomicAdd(ph + i, (unsigned long long)ck[i]);
            atomicAdd(ph + 7, (unsigned long long)issued);
            for (int i = 11; i < 16; i++) atomicAdd(ph + i, (unsigned long long)ck[i]);
        }
#undef JV_PHASE
    }
}

// The kernel pays off where the table fills the SM (M = 161..192: 48 KB, 4 CTAs per SM either way).  With the smaller tables of
// M <= 128 the round-synchronous kernel runs 5..8 CTAs per SM on 64..95 registers and is faster (measured at M = 48: 0.85 ms
// against 1.26 ms), and the 64 KB table of M = 256 leaves no room for four of these CTAs.
bool q8_beam_supported(const jv_index *ix, int L, int R, int E) {
    if (!ix->has_pq || !ix->q8_ok) return false;
    if (L > kBeamList || R > 64 || E * ((R + 31) / 32) > 4) return false;
    return ix->q8_nj == 6;
}

template <int NJ, int SW, bool PROF>
static int32_t launch_beam_typed(jv_index *ix, SearchCtx *ctx, Q8Params &p, int depth) {
    constexpr int kT = (SW + 2) * 32;
    auto kern = q8_beam_kernel<NJ, SW, PROF>;
    const size_t fixed = (size_t)p.lutb + kBeamList * 8 + 2 * kBeamPool * 8 + 2 * kBeamPool * 4;
    const size_t sm_total = 228 * 1024;
    // visited filter: as large as 4 CTAs per SM allow (more CTAs than that the register file does not hold), 2 tags per word
    int lg = 10;
    const int64_t per = (int64_t)(sm_total / 4) - 1024 - 256 - (int64_t)fixed;
    if (per < 4096) {
        set_error("search (8-bit table, beam kernel): shared memory budget exceeded (%zu fixed bytes)", fixed);
        return JV_ERR_UNSUPPORTED;
    }
    const int64_t want = (int64_t)p.L * p.R; // words; 2 tags each
    while (lg < 13 && ((int64_t)4 << (lg + 1)) <= per && ((int64_t)1 << lg) < want) lg++;
    p.hash_log2 = lg;
    const size_t smem = fixed + ((size_t)4 << lg);
    JV_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    JV_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kT, smem));
    if (occ < 1) {
        set_error("search (8-bit table, beam kernel): kernel does not fit on an SM (smem %zu)", smem);
        return JV_ERR_UNSUPPORTED;
    }
    if (q8_knobs().occ >= 1 && q8_knobs().occ < occ) occ = q8_knobs().occ; // diagnostics: cap the CTAs per SM
    int grid = ix->sm_count * occ;
    if (grid > p.nq) grid = p.nq;
    JV_CUDA_TRY(cudaMemsetAsync(p.work_counter, 0, sizeof(int), ctx->stream));
    kern<<<grid, kT, smem, ctx->stream>>>(p, depth);
    JV_CUDA_TRY(cudaGetLastError());
    return JV_OK;
}

int32_t launch_q8_beam(jv_index *ix, SearchCtx *ctx, Q8Params &p) {
    const Q8Knobs &kn = q8_knobs();
    int depth = kn.depth > 0 ? kn.depth : 2;
    if (depth > 2) depth = 2;
    if (p.E == 1) depth = 1; // explicit width 1 = the reference's best-first order
    const bool prof = kn.prof;
    if (ix->q8_nj == 6) { // diagnostic instantiations: phase counters, 3 / 7 scorer warps
        if (prof) return launch_beam_typed<6, 4, true>(ix, ctx, p, depth);
        if (kn.warps == 8) return launch_beam_typed<6, 7, false>(ix, ctx, p, depth);
        if (kn.warps == 4) return launch_beam_typed<6, 3, false>(ix, ctx, p, depth);
        return launch_beam_typed<6, 4, false>(ix, ctx, p, depth);
    }
    set_error("search (8-bit table, beam kernel): unsupported code row width");
    return JV_ERR_UNSUPPORTED;
}

}  // namespace jv
