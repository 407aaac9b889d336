// jv_q8_pipe.cu — K2, production kernel of the 8-bit table path: software-pipelined beam search.
//
// Reference loop: GraphSearcher.search over the PQ score function (JVectorReader.java:165-173).
//
// One CTA owns one query (its 8-bit ADC table is staged in shared memory with one TMA bulk copy, like the synchronous
// kernel in jv_q8.cu), but the warps of the CTA no longer move in lock step.  Each warp runs its own expansion loop:
//
//     take the turn -> merge the survivors of my previous expansion into the shared sorted list -> pick the best
//     unexpanded entry -> pass the turn on -> adjacency row (one coalesced 128-byte load) -> visited filter -> code rows of
//     the fresh neighbours (16-byte vector loads, all rows in flight before the first table lookup) -> table lookups ->
//     keys that beat the list's worst entry stay in registers (one per lane) until my next turn
//
// The turn is a token that travels round robin over the warps (one mbarrier per warp: arrive = hand over with release
// semantics, try_wait = sleep until it is mine), so list updates are serialised without a lock and without any block
// barrier inside a query; a warp that waits for DRAM never stops the other warps of its query, nor the other three
// queries on the SM.  With E warps in flight the search is a best-first search whose selections lag E-1 expansions behind
// (expand_width = 1: exactly the reference order of the oracle's 8-bit mode).  Which of two concurrently expanding
// warps scores a common neighbour depends on timing, so for E > 1 the visiting order (and, rarely, an id of the
// approximate list) can differ from run to run; recall is gated like every wide-step result (tests/test_gpu_q8.py).
//
// Everything a step of the synchronous kernel needed in shared memory besides the table — second list buffer, survivor
// queue, neighbour pool, gap counters — is gone: survivors live in registers, the merge is in place.
#include "jv_q8.cuh"

namespace jv {

constexpr int q8p_min_ctas(int nj, int warps) {
    const int by_smem = (nj == 1 || nj == 2) ? 8 : nj == 3 ? 6 : nj == 4 ? 5 : 4;
    const int by_regs = 65536 / (warps * 32 * 64); // never ask for fewer than 64 registers per thread
    return by_smem < by_regs ? by_smem : by_regs;
}

// PROF: per-phase cycle counters of every expanding warp (jv_index_debug_counter 8..): turn wait, merge, select + hand-over,
// adjacency row + visited filter, code rows in registers, table lookups + keys; slot 7 = expansions
template <int NJ_T, int W, bool PROF>
__global__ void __launch_bounds__(W * 32, q8p_min_ctas(NJ_T, W)) q8_pipe_kernel(const Q8Params p) {
    constexpr int kThreadsQ = W * 32;
    constexpr int NJ = NJ_T;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = p.L, H = 1 << p.hash_log2, R = p.R;
    const int WA = p.E < W ? p.E : W; // warps that expand (the others only help at query boundaries)

    unsigned char *sp = smem_raw;
    const uint8_t *lut = sp;
    sp += p.lutb;
    uint64_t *list = reinterpret_cast<uint64_t *>(sp); // sorted, best first; bit 0 = unexpanded
    sp += (size_t)(L < 64 ? 64 : ((L + 1) & ~1)) * 8;
    uint64_t *wsurv = reinterpret_cast<uint64_t *>(sp) + warp * 64; // this warp's survivors, best first, per adjacency chunk
    sp += (size_t)W * 512;
    int32_t *wscr = reinterpret_cast<int32_t *>(sp) + warp * 32; // this warp's fresh neighbour ids
    sp += (size_t)W * 128;
    uint32_t *filter = reinterpret_cast<uint32_t *>(sp);

    __shared__ __align__(8) uint64_t s_bar;      // table arrival
    __shared__ __align__(8) uint64_t s_tok[W];   // the turn: one barrier per warp
    __shared__ __align__(8) uint64_t s_worst;    // key >> 1 of the list's last entry once the list is full (admission threshold)
    __shared__ uint32_t s_ctl;                   // n | inflight << 16 | exited << 24 | done << 31 (owned by the turn holder)
    __shared__ int s_query, s_vis, s_exp;

    const bool tagged = p.n <= ((int64_t)1 << (p.hash_log2 + 15));
    const bool isum_keys = p.sim != JV_SIM_COSINE;
    const bool l2 = p.sim == JV_SIM_EUCLIDEAN;
    // ADC lane geometry (header of jv_q8.cu): group g = lane / 8 scores one code row, lane sl owns subspaces m = 8t + sl; at
    // lookup (j, i) the group reads bank quarter (i + g) & 3, so the 32 lanes of a warp always hit 32 different banks
    const int g = lane >> 3, sl = lane & 7;
    uint32_t sel[4], lb[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t qd = (uint32_t)(i + g) & 3u;
        sel[i] = 0x4400u | (qd << 4) | (4u + qd); // byte 0 <- (cw >> 6).byte[qd], byte 1 <- cw.byte[qd]
        lb[i] = qd * 32u + (uint32_t)sl * 4u;
    }

    if (tid == 0) {
        mbar_init(&s_bar, 1);
        for (int w = 0; w < W; w++) mbar_init(&s_tok[w], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t phase = 0, tokpar = 0; // parities of the table barrier and of my turn barrier (they persist across queries)

    auto lookup4 = [&](uint32_t cw, int j) -> uint32_t {
        const uint32_t sh = cw >> 6;
        const uint8_t *base = lut + (j >> 1) * 16384 + (j & 1) * 128;
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t v = __byte_perm(cw, sh, sel[i]);
            s += base[(v & 0x3F03u) | lb[i]];
        }
        return s;
    };
    auto reduce8 = [&](uint32_t s) -> uint32_t {
        s += __shfl_xor_sync(JV_FULL_MASK, s, 4);
        s += __shfl_xor_sync(JV_FULL_MASK, s, 2);
        s += __shfl_xor_sync(JV_FULL_MASK, s, 1);
        return s;
    };

    for (;;) {
        __syncthreads(); // everyone is done with the previous query's table, list and counters
        if (tid == 0) {
            s_query = atomicAdd(p.work_counter, 1);
            s_ctl = 0u;
            s_vis = 0;
            s_exp = 0;
            s_worst = 0ull;
        }
        __syncthreads();
        const int qi = s_query;
        if (qi >= p.nq) break;
        if (tid == 0) { // K1 result: one TMA bulk copy HBM/L2 -> shared memory
            mbar_expect_tx(&s_bar, (uint32_t)p.lutb);
            bulk_g2s(smem_raw, p.lut + (int64_t)qi * p.lutb, (uint32_t)p.lutb, &s_bar);
        }
        for (int i = tid; i < H; i += kThreadsQ) filter[i] = tagged ? 0u : kEmpty;
        const float4 qp = __ldg(p.qparams + qi);
        const float delta = qp.x, base = qp.y, qnorm = qp.z;
        auto score_of = [&](uint32_t isum, int32_t nb) -> float {
            const float s = __fmaf_rn(delta, (float)isum, base);
            const float nn = p.sim == JV_SIM_COSINE ? __ldg(p.node_norm + nb) : 0.f;
            return adc_finish(p.sim, s, nn, qnorm);
        };
        auto pack_key = [&](uint32_t isum, int32_t nb) -> uint64_t { // full list key of a freshly scored node (unexpanded)
            const uint32_t ord = isum_keys ? (l2 ? ~isum : isum) : jv_f2ord(score_of(isum, nb));
            return qkey_pack(ord, nb);
        };
        __syncthreads(); // filter cleared
        mbar_wait(&s_bar, phase);
        phase ^= 1u;

        if (warp == 0 && p.entry >= 0 && p.entry < p.n) { // seed: the entry node (its row is scored by group 0)
            uint32_t s = 0;
            if (g == 0) {
                const unsigned char *row = p.codes_q8 + (int64_t)p.entry * (NJ * 32);
#pragma unroll
                for (int j = 0; j < NJ; j++) s += lookup4(__ldg(reinterpret_cast<const uint32_t *>(row + q8_word_offset(NJ, sl, j))), j);
            }
            s = reduce8(s);
            if (lane == 0) {
                list[0] = pack_key(s, p.entry);
                q_filter_insert(filter, p.hash_log2, tagged, p.entry);
                s_ctl = 1u;
                s_vis = 1;
            }
        }
        __syncthreads();

        if (warp < WA) {
            const int next = warp + 1 == WA ? 0 : warp + 1;
            const int RB = (R + 31) >> 5; // adjacency chunks of 32 neighbours (R <= 64)
            bool first = warp == 0; // warp 0 owns the first turn of every query
            bool busy = false;      // an expansion of mine is waiting to be merged
            int ns0 = 0, ns1 = 0;   // survivors of my expansion per adjacency chunk, sorted best first in wsurv[chunk] ...
            uint64_t sa0 = 0ull, sa1 = 0ull; // ... and in registers: lane j holds the j-th best (key >> 1)
            int my_vis = 0, my_exp = 0;
            long long ck[6] = {0, 0, 0, 0, 0, 0}, ck_ns = 0;
            long long t_prev = PROF ? clock64() : 0;
#define JV_PHASE(i)                        \
    if (PROF) {                            \
        const long long t_now = clock64(); \
        ck[i] += t_now - t_prev;           \
        t_prev = t_now;                    \
    }
            for (;;) {
                // ---------------------------------------------------------------- wait for the turn
                if (!first) {
                    if (!mbar_try_wait(&s_tok[warp], tokpar)) {
                        const long long t0 = clock64();
                        while (!mbar_try_wait(&s_tok[warp], tokpar))
                            if (clock64() - t0 > 4000000000ll) __trap(); // a lost token must not hang the device (never seen)
                    }
                    tokpar ^= 1u;
                }
                first = false;
                JV_PHASE(0)
                // control word (only ever touched inside a turn): n | inflight << 16 | exited << 24 | done << 31
                const uint32_t ctl = *reinterpret_cast<volatile uint32_t *>(&s_ctl);
                int n = (int)(ctl & 0xffffu);
                int infl = (int)((ctl >> 16) & 0xffu) - (busy ? 1 : 0); // expansions whose survivors are not merged yet
                int exited = (int)((ctl >> 24) & 0x7fu);
                bool done = (ctl >> 31) != 0u;
                int cand = -1;
                if (!done && L <= 64) {
                    // ------------------------------------------------------------ short lists (rerankK <= 64, every production
                    // shape): the whole list sits in registers, two entries per lane.  One pass per adjacency chunk computes where
                    // everything goes — survivor j broadcasts its key, two votes count the entries in front of it, every lane counts
                    // the survivors in front of its own entries — then the best unexpanded element is picked from the registers
                    // (one warp-wide min) and list, flags and admission threshold are written once.
                    const int nb = busy ? RB : 1;
#pragma unroll 1
                    for (int b = 0; b < nb; b++) {
                        uint64_t k0 = list[lane], k1 = list[lane + 32];
                        if (lane >= n) k0 = 0ull; // key >> 1 of a missing entry = 0: never in front of a survivor
                        if (lane + 32 >= n) k1 = 0ull;
                        const uint64_t e0 = k0 >> 1, e1 = k1 >> 1;
                        const uint64_t worst = *reinterpret_cast<volatile uint64_t *>(&s_worst);
                        const uint64_t *ws = wsurv + b * 32;
                        uint64_t a = 0ull;
                        int ns = 0;
                        if (busy) {
                            a = b == 0 ? sa0 : sa1;
                            ns = __popc(__ballot_sync(JV_FULL_MASK, a > worst)); // sorted best first: the losers are a suffix
                        }
                        int lo = 0, idx = 0, g0 = 0, g1 = 0, nk = 0;
                        bool keep = false;
                        if (PROF) ck_ns += ns;
#pragma unroll 4
                        for (int j = 0; j < ns; j++) { // branch-free: the iterations pipeline
                            const uint64_t aj = ws[j];
                            const bool gt0 = e0 > aj, gt1 = e1 > aj;
                            const int c = __popc(__ballot_sync(JV_FULL_MASK, gt0)) + __popc(__ballot_sync(JV_FULL_MASK, gt1));
                            lo = lane == j ? c : lo;
                            g0 += gt0 ? 0 : 1;
                            g1 += gt1 ? 0 : 1;
                        }
                        nk = ns;
                        idx = lane;
                        keep = lane < ns;
                        // an entry equal to a survivor is the same node, scored again after the visited filter forgot it (rare):
                        // the entry at position lo is the only one that can be equal
                        {
                            const uint32_t xl0 = __shfl_sync(JV_FULL_MASK, (uint32_t)e0, lo & 31), xh0 = __shfl_sync(JV_FULL_MASK, (uint32_t)(e0 >> 32), lo & 31);
                            const uint32_t xl1 = __shfl_sync(JV_FULL_MASK, (uint32_t)e1, lo & 31), xh1 = __shfl_sync(JV_FULL_MASK, (uint32_t)(e1 >> 32), lo & 31);
                            const uint64_t at = lo < 32 ? (((uint64_t)xh0 << 32) | xl0) : (((uint64_t)xh1 << 32) | xl1);
                            const uint32_t dm = __ballot_sync(JV_FULL_MASK, keep && at == a);
                            if (dm) { // slow path: recount without the duplicates
                                lo = idx = g0 = g1 = nk = 0;
                                keep = false;
                                for (int j = 0; j < ns; j++) {
                                    const uint64_t aj = ws[j];
                                    const bool gt0 = e0 > aj, gt1 = e1 > aj;
                                    const uint32_t v0 = __ballot_sync(JV_FULL_MASK, gt0), v1 = __ballot_sync(JV_FULL_MASK, gt1);
                                    const bool ok = ((dm >> j) & 1u) == 0u;
                                    const bool me = ok && lane == j;
                                    lo = me ? __popc(v0) + __popc(v1) : lo;
                                    idx = me ? nk : idx;
                                    keep = keep || me;
                                    g0 += (ok && !gt0) ? 1 : 0;
                                    g1 += (ok && !gt1) ? 1 : 0;
                                    nk += ok ? 1 : 0;
                                }
                            }
                        }
                        // final positions (>= L: dropped)
                        const int p0 = lane < n ? lane + g0 : 0x7fffffff, p1 = lane + 32 < n ? lane + 32 + g1 : 0x7fffffff;
                        const int ps = keep ? lo + idx : 0x7fffffff;
                        n = n + nk < L ? n + nk : L;
                        uint64_t w0 = k0, w1 = k1, wsv = (a << 1) | 1ull;
                        if (b == nb - 1) { // best unexpanded element = the next candidate
                            int best = 0x7fffffff, which = 0;
                            if ((k0 & 1ull) && p0 < L) best = p0, which = 0;
                            if ((k1 & 1ull) && p1 < L && p1 < best) best = p1, which = 1;
                            if (ps < L && ps < best) best = ps, which = 2;
                            const int bmin = (int)__reduce_min_sync(JV_FULL_MASK, (unsigned)best);
                            if (bmin != 0x7fffffff) {
                                const bool mine = best == bmin;
                                const uint32_t low = which == 0 ? (uint32_t)k0 : which == 1 ? (uint32_t)k1 : (uint32_t)wsv;
                                cand = qkey_node((uint64_t)__shfl_sync(JV_FULL_MASK, low, __ffs(__ballot_sync(JV_FULL_MASK, mine)) - 1));
                                if (mine) {
                                    if (which == 0) w0 &= ~1ull;
                                    else if (which == 1) w1 &= ~1ull;
                                    else wsv &= ~1ull;
                                }
                            }
                        }
                        __syncwarp(); // every lane has read its entries
                        if (p0 < L && (g0 > 0 || w0 != k0)) list[p0] = w0; // moved or flag cleared
                        if (p1 < L && (g1 > 0 || w1 != k1)) list[p1] = w1;
                        if (ps < L) {
                            list[ps] = wsv;
                            if (ps < 2 * WA) { // a newcomer near the top is about to be expanded: bring its adjacency row into L2 now
                                const char *row = reinterpret_cast<const char *>(p.adjacency + (int64_t)qkey_node(wsv) * R);
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
                                if (R > 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 128));
                            }
                        }
                        if (n >= L && nk > 0) { // the element that lands on the last slot is the new admission threshold
                            if (p0 == L - 1) *reinterpret_cast<volatile uint64_t *>(&s_worst) = e0;
                            if (p1 == L - 1) *reinterpret_cast<volatile uint64_t *>(&s_worst) = e1;
                            if (ps == L - 1) *reinterpret_cast<volatile uint64_t *>(&s_worst) = a;
                        }
                        if (b + 1 < nb) __syncwarp();
                    }
                    busy = false;
                    JV_PHASE(1)
                    if (cand >= 0) infl++;
                    else if (infl == 0) done = true; // nothing left and nobody can add anything
                } else if (!done) {
                    // ------------------------------------------------------------ long lists: merge my survivors in place
                    if (busy) {
#pragma unroll 1
                        for (int b = 0; b < RB; b++) {
                            int ns = b == 0 ? ns0 : ns1;
                            if (ns == 0) continue;
                            const uint64_t *ws = wsurv + b * 32;
                            const uint64_t worst = n >= L ? (list[L - 1] >> 1) : 0ull;
                            uint64_t a = lane < ns ? ws[lane] : 0ull;
                            ns = __popc(__ballot_sync(JV_FULL_MASK, a > worst)); // sorted best first: the losers are a suffix
                            if (ns == 0) continue;
                            // Phase A — where every survivor goes: survivor j broadcasts its key (one shared-memory read), the 32
                            // lanes compare it with 64 list entries at a time, two votes give the number of better entries.  No
                            // dependent chain: the iterations over j pipeline.
                            int lo = 0;
                            uint64_t k0 = 0ull, k1 = 0ull;
                            for (int c0 = 0; c0 < n; c0 += 64) {
                                const int t0 = c0 + lane, t1 = t0 + 32;
                                k0 = t0 < n ? list[t0] : 0ull; // key >> 1 of a missing entry = 0: never better than a survivor
                                k1 = t1 < n ? list[t1] : 0ull;
                                const uint64_t e0 = k0 >> 1, e1 = k1 >> 1;
#pragma unroll 4
                                for (int j = 0; j < ns; j++) {
                                    const uint64_t aj = ws[j];
                                    const uint32_t v0 = __ballot_sync(JV_FULL_MASK, e0 > aj), v1 = __ballot_sync(JV_FULL_MASK, e1 > aj);
                                    if (lane == j) lo += __popc(v0) + __popc(v1);
                                }
                            }
                            // an equal entry is the same node, scored again after the visited filter forgot it (rare): drop it and
                            // close the gap (keys and positions of the others stay valid)
                            const bool dup = lane < ns && lo < n && (list[lo] >> 1) == a;
                            if (__any_sync(JV_FULL_MASK, dup)) {
                                const uint32_t keep = __ballot_sync(JV_FULL_MASK, lane < ns && !dup);
                                const int to = __popc(keep & ((1u << lane) - 1u));
                                __syncwarp();
                                if (lane < ns && !dup) {
                                    wsurv[b * 32 + to] = a;
                                    wscr[to] = lo;
                                }
                                __syncwarp();
                                ns = __popc(keep);
                                if (ns == 0) continue;
                                a = lane < ns ? ws[lane] : 0ull;
                                lo = lane < ns ? wscr[lane] : 0;
                                __syncwarp();
                            }
                            const int minlo = __shfl_sync(JV_FULL_MASK, lo, 0);
                            // Phase B — list entry t moves down by the number of survivors better than it; 64 entries per round
                            // from the bottom up, reads before writes (a single round reuses the entries phase A left in registers)
                            for (int c0 = (n - 1) & ~63; c0 >= (minlo & ~63); c0 -= 64) {
                                const int t0 = c0 + lane, t1 = t0 + 32;
                                if (n > 64) {
                                    k0 = t0 < n ? list[t0] : 0ull;
                                    k1 = t1 < n ? list[t1] : 0ull;
                                }
                                const uint64_t e0 = k0 >> 1, e1 = k1 >> 1;
                                int l0 = 0, l1 = 0;
#pragma unroll 4
                                for (int j = 0; j < ns; j++) {
                                    const uint64_t aj = ws[j];
                                    l0 += aj > e0 ? 1 : 0;
                                    l1 += aj > e1 ? 1 : 0;
                                }
                                __syncwarp();
                                if (t0 < n && l0 > 0 && t0 + l0 < L) list[t0 + l0] = k0;
                                if (t1 < n && l1 > 0 && t1 + l1 < L) list[t1 + l1] = k1;
                                __syncwarp();
                            }
                            if (lane < ns && lo + lane < L) {
                                list[lo + lane] = (a << 1) | 1ull;
                                // a newcomer near the top is about to be expanded: bring its adjacency row into L2 now
                                if (lo + lane < 2 * WA) {
                                    const char *row = reinterpret_cast<const char *>(p.adjacency + (int64_t)qkey_node(a << 1) * R);
                                    asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
                                    if (R > 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 128));
                                }
                            }
                            __syncwarp();
                            n = n + ns < L ? n + ns : L;
                        }
                        busy = false;
                    }
                    JV_PHASE(1)
                    // ------------------------------------------------------------ best unexpanded entry
                    for (int c0 = 0; c0 < n; c0 += 32) {
                        const int t = c0 + lane;
                        const uint64_t k = t < n ? list[t] : 0ull;
                        const uint32_t un = __ballot_sync(JV_FULL_MASK, (k & 1ull) != 0ull);
                        if (un) {
                            const int j = __ffs(un) - 1;
                            cand = qkey_node((uint64_t)__shfl_sync(JV_FULL_MASK, (uint32_t)k, j));
                            if (lane == j) list[t] = k & ~1ull;
                            break;
                        }
                    }
                    if (cand >= 0) infl++;
                    else if (infl == 0) done = true; // nothing left and nobody can add anything
                }
                // ---------------------------------------------------------------- hand the turn over
                const bool leave = done;
                if (leave) exited++;
                __syncwarp(); // orders the other lanes' list writes before lane 0's release
                if (lane == 0) {
                    *reinterpret_cast<volatile uint32_t *>(&s_ctl) =
                        (uint32_t)n | ((uint32_t)infl << 16) | ((uint32_t)exited << 24) | (done ? 0x80000000u : 0u);
                    if (!leave && L > 64) *reinterpret_cast<volatile uint64_t *>(&s_worst) = n >= L ? (list[L - 1] >> 1) : 0ull;
                    if (!leave || exited < WA) mbar_arrive(&s_tok[next]); // the last warp out keeps the token: every arrive has had its wait
                }
                JV_PHASE(2)
                if (leave) break;
                if (cand < 0) continue; // nothing to expand right now: other warps are still producing candidates

                // ---------------------------------------------------------------- expand `cand`
                my_exp++;
                ns0 = ns1 = 0;
                sa0 = sa1 = 0ull;
#pragma unroll 1
                for (int r0 = 0, ch = 0; r0 < R; r0 += 32, ch++) {
                    int32_t nb = -1;
                    if (r0 + lane < R) nb = __ldg(p.adjacency + (int64_t)cand * R + r0 + lane);
                    const bool fresh = nb >= 0 && nb < p.n && q_filter_insert(filter, p.hash_log2, tagged, nb);
                    const uint32_t fb = __ballot_sync(JV_FULL_MASK, fresh);
                    const int nf = __popc(fb);
                    JV_PHASE(3)
                    if (nf == 0) continue;
                    if (fresh) wscr[__popc(fb & ((1u << lane) - 1u))] = nb;
                    __syncwarp();
                    my_vis += nf;
                    uint64_t ka = 0ull;
                    // fresh row r goes to group r & 3 as its row r >> 2; its key ends up in lane (r & 3) * 8 + (r >> 2)
                    auto pass_rows = [&](int u0, auto nu_tag) {
                        constexpr int NU = decltype(nu_tag)::value;
                        uint32_t cw[NU][NJ], s[NU];
                        int32_t nbv[NU];
#pragma unroll
                        for (int u = 0; u < NU; u++) {
                            const int idx = (u0 + u) * 4 + g;
                            nbv[u] = idx < nf ? wscr[idx] : -1;
                            s[u] = 0u;
                            if (nbv[u] >= 0) {
                                q8_load_row<NJ>(p.codes_q8 + (int64_t)nbv[u] * (NJ * 32), sl, cw[u]);
                            } else {
#pragma unroll
                                for (int j = 0; j < NJ; j++) cw[u][j] = 0u;
                            }
                        }
                        if (PROF) { // code words in registers
                            uint32_t acc = 0;
#pragma unroll
                            for (int u = 0; u < NU; u++)
#pragma unroll
                                for (int j = 0; j < NJ; j++) acc |= cw[u][j];
                            asm volatile("" ::"r"(acc));
                            JV_PHASE(4)
                        }
#pragma unroll
                        for (int j = 0; j < NJ; j++) {
#pragma unroll
                            for (int u = 0; u < NU; u++) s[u] += lookup4(cw[u][j], j); // a group without a row looks up code 0: harmless
                        }
#pragma unroll
                        for (int u = 0; u < NU; u++) s[u] = reduce8(s[u]);
                        // after the butterfly every lane of a group holds the group's sums: lane sl = u0 + u keeps row u
                        uint32_t my_s = 0u;
                        int32_t my_nb = -1;
#pragma unroll
                        for (int u = 0; u < NU; u++) {
                            my_s = sl == u0 + u ? s[u] : my_s;
                            my_nb = sl == u0 + u ? nbv[u] : my_nb;
                        }
                        if (my_nb >= 0) ka = pack_key(my_s, my_nb) >> 1;
                    };
                    for (int u0 = 0; u0 * 4 < nf; u0 += 3) {
                        const int left = nf - u0 * 4;
                        if (left > 8)
                            pass_rows(u0, std::integral_constant<int, 3>());
                        else if (left > 4)
                            pass_rows(u0, std::integral_constant<int, 2>());
                        else
                            pass_rows(u0, std::integral_constant<int, 1>());
                    }
                    // survivors (the threshold may be stale: re-checked at the merge), ranked among themselves and parked in
                    // shared memory best first — all of it outside the turn
                    const uint64_t worst = *reinterpret_cast<volatile uint64_t *>(&s_worst);
                    if (ka <= worst) ka = 0ull;
                    const uint32_t sm = __ballot_sync(JV_FULL_MASK, ka != 0ull);
                    if (sm) {
                        int rank = 0; // survivors better than mine (keys are distinct: the node is part of the key)
                        uint32_t m = sm;
                        while (m) { // four broadcasts in flight per round
                            const int j0 = __ffs(m) - 1;
                            m &= m - 1u;
                            const int j1 = m ? __ffs(m) - 1 : j0;
                            m &= m - 1u;
                            const int j2 = m ? __ffs(m) - 1 : j0;
                            m &= m - 1u;
                            const int j3 = m ? __ffs(m) - 1 : j0;
                            m &= m - 1u;
                            const uint64_t o0 = __shfl_sync(JV_FULL_MASK, ka, j0), o1 = __shfl_sync(JV_FULL_MASK, ka, j1);
                            const uint64_t o2 = __shfl_sync(JV_FULL_MASK, ka, j2), o3 = __shfl_sync(JV_FULL_MASK, ka, j3);
                            rank += (o0 > ka ? 1 : 0) + (j1 != j0 && o1 > ka ? 1 : 0) + (j2 != j0 && o2 > ka ? 1 : 0) + (j3 != j0 && o3 > ka ? 1 : 0);
                        }
                        if (ka != 0ull) wsurv[ch * 32 + rank] = ka;
                        __syncwarp();
                        const int nsc = __popc(sm);
                        const uint64_t sorted = lane < nsc ? wsurv[ch * 32 + lane] : 0ull; // lane j holds the j-th best
                        if (ch == 0) ns0 = nsc, sa0 = sorted;
                        else ns1 = nsc, sa1 = sorted;
                    }
                    __syncwarp(); // wscr is rewritten by the next chunk; wsurv is read at the merge
                    JV_PHASE(5)
                }
                busy = true;
            }
            if (lane == 0) {
                atomicAdd(&s_vis, my_vis);
                atomicAdd(&s_exp, my_exp);
                if (PROF && p.dbg) {
                    unsigned long long *ph = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(p.dbg) + 64);
                    for (int i = 0; i < 6; i++) atomicAdd(ph + i, (unsigned long long)ck[i]);
                    atomicAdd(ph + 6, (unsigned long long)ck_ns); // survivors that reached a merge
                    atomicAdd(ph + 7, (unsigned long long)my_exp);
                }
            }
#undef JV_PHASE
        }
        __syncthreads(); // the list is final

        // ---- emit the approximate result list, best first, in the (score, ~node) key format of the rerank step
        {
            const int n = (int)(s_ctl & 0xffffu);
            uint64_t *o = p.approx_keys + (int64_t)qi * L;
            for (int i = tid; i < L; i += kThreadsQ) {
                uint64_t out = 0ull;
                if (i < n) {
                    const uint64_t k = list[i];
                    const int32_t node = qkey_node(k);
                    const uint32_t ord = (uint32_t)(k >> 32);
                    const float sc = isum_keys ? score_of(l2 ? ~ord : ord, node) : jv_ord2f(ord);
                    out = jv_mk_key(sc, node);
                }
                o[i] = out;
            }
            if (tid == 0) {
                p.approx_count[qi] = n;
                if (p.stats) {
                    jv_query_stats st;
                    st.visited = s_vis;
                    st.expanded = s_exp;
                    st.expanded_base = s_exp;
                    st.reranked = 0;
                    p.stats[qi] = st;
                }
            }
        }
    }
}

template <int NJ_T, int W, bool PROF = false> static int32_t launch_q8_pipe_typed(jv_index *ix, SearchCtx *ctx, Q8Params &p) {
    constexpr int kThreadsQ = W * 32;
    auto kern = q8_pipe_kernel<NJ_T, W, PROF>;
    const size_t fixed = (size_t)p.lutb + (size_t)(p.L < 64 ? 64 : ((p.L + 1) & ~1)) * 8 + (size_t)W * 640;
    const size_t sm_total = 228 * 1024;
    int64_t want = (int64_t)p.L * p.R; // words; 2 tags each
    if (want < 1024) want = 1024;
    int best_occ = 0, best_log2 = 0;
    for (int occ = 8; occ >= 1; occ--) {
        const int64_t per = (int64_t)(sm_total / occ) - 1024 - 256 - (int64_t)fixed; // 1 KB system + static __shared__
        if (per < 4096) continue;
        int lg = 10;
        while (lg < 15 && ((int64_t)4 << (lg + 1)) <= per && ((int64_t)1 << lg) < want) lg++;
        best_occ = occ;
        best_log2 = lg;
        break;
    }
    if (!best_occ) {
        set_error("search (8-bit table, pipelined): shared memory budget exceeded (%zu fixed bytes)", fixed);
        return JV_ERR_UNSUPPORTED;
    }
    p.hash_log2 = best_log2;
    const size_t smem = fixed + ((size_t)4 << best_log2);
    JV_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    JV_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreadsQ, smem));
    if (occ < 1) {
        set_error("search (8-bit table, pipelined): kernel does not fit on an SM (smem %zu)", smem);
        return JV_ERR_UNSUPPORTED;
    }
    if (q8_knobs().occ >= 1 && q8_knobs().occ < occ) occ = q8_knobs().occ; // diagnostics: cap the CTAs per SM
    int grid = ix->sm_count * occ;
    if (grid > p.nq) grid = p.nq;
    JV_CUDA_TRY(cudaMemsetAsync(p.work_counter, 0, sizeof(int), ctx->stream));
    kern<<<grid, kThreadsQ, smem, ctx->stream>>>(p);
    JV_CUDA_TRY(cudaGetLastError());
    return JV_OK;
}

bool q8_pipe_supported(const jv_index *ix, int L, int R) {
    const int nj = ix->q8_nj;
    if (!(nj == 1 || nj == 2 || nj == 3 || nj == 4 || nj == 6 || nj == 8) || R > 64 || L > 65535) return false;
    const size_t fixed = (size_t)q8_lut_bytes(nj) + (size_t)(L < 64 ? 64 : ((L + 1) & ~1)) * 8 + 8 * 640;
    return fixed + 4096 + 2048 <= 227 * 1024;
}

int32_t launch_q8_pipe(jv_index *ix, SearchCtx *ctx, Q8Params &p, int warps) {
#define JV_PIPE_CASE(NJV)                                                                                        \
    case NJV:                                                                                                    \
        return warps == 8 ? launch_q8_pipe_typed<NJV, 8>(ix, ctx, p) : launch_q8_pipe_typed<NJV, 4>(ix, ctx, p);
    if (q8_knobs().prof && ix->q8_nj == 6) return launch_q8_pipe_typed<6, 4, true>(ix, ctx, p); // diagnostics: headline shape only
    switch (ix->q8_nj) {
        JV_PIPE_CASE(1)
        JV_PIPE_CASE(2)
        JV_PIPE_CASE(3)
        JV_PIPE_CASE(4)
        JV_PIPE_CASE(6)
        JV_PIPE_CASE(8)
    default:
        set_error("search (8-bit table, pipelined): unsupported code row width");
        return JV_ERR_UNSUPPORTED;
    }
#undef JV_PIPE_CASE
}

}  // namespace jv
