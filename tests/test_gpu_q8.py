"""The production traversal with the 8-bit quantised ADC table (JV_INDEX_FLAG_LUT_U8, csrc/jv_q8.cu).

* K1: the batched table kernel must reproduce the oracle's quantised tables and (delta, base) bit for bit.
* K2: integer sums are order-independent, so at expand_width = 1 the traversal is the oracle's (adc_order = -8) best-first
  order up to score ties at the list boundary (integer sums tie more often than fp32 ones): ids identical for almost
  every query; for wider expansion the north-star gate applies (recall@10 within 0.005 at bench scale; unit-test
  fixtures of 64..100 queries have a sampling error of ~0.01).
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import clustered, embedded, make_fixture, recall

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fx_dot():  # sub-dim 4, M = 16 (half a bank block)
    base, q = clustered(8000, 64, 100, seed=41, normalize=True)
    return make_fixture(O.SIM_DOT, base, q, max_degree=32, pq_m=16)


@pytest.fixture(scope="module")
def fx_l2():  # sub-dim 2, M = 48 (1.5 bank blocks), centred codebooks
    base, q = clustered(6000, 96, 64, seed=42)
    return make_fixture(O.SIM_EUCLIDEAN, base, q, max_degree=16, pq_m=48)


@pytest.fixture(scope="module")
def fx_cos8():  # sub-dim 8
    base, q = clustered(4000, 128, 64, seed=43)
    return make_fixture(O.SIM_COSINE, base, q, max_degree=16, pq_m=16)


@pytest.fixture(scope="module")
def fx_dot192():  # the cfg-2 shape at test scale: 768-d, M = 192 (6 bank blocks), R = 32
    base, q = embedded(5000, 768, 64, seed=44)
    return make_fixture(O.SIM_DOT, base, q, max_degree=32, pq_m=192)


ALL = ["fx_dot", "fx_l2", "fx_cos8", "fx_dot192"]


@pytest.mark.parametrize("name", ALL)
def test_quantised_tables_match_oracle_bit_for_bit(jv, request, name):
    fx = request.getfixturevalue(name)
    ora = fx.oracle_index(adc_order=-8)
    with fx.gpu_index(jv, flags=jv.native.FLAG_LUT_U8) as gi:
        nq = min(len(fx.queries), 37)  # not a multiple of the queries-per-CTA tile
        q8, prm = gi.pq_lut_q8(fx.queries[:nq])
        w8, wprm = ora.lut_q8(fx.queries[:nq])
        np.testing.assert_array_equal(prm.view(np.uint32), wprm.view(np.uint32))
        np.testing.assert_array_equal(q8, w8)
        assert q8.max() > 128  # the shared scale is actually used


@pytest.mark.parametrize("name", ALL)
def test_width1_follows_the_oracle_order(jv, request, name):
    fx = request.getfixturevalue(name)
    ora = fx.oracle_index(adc_order=-8)
    with fx.gpu_index(jv, flags=jv.native.FLAG_LUT_U8) as gi:
        gt, _, _ = gi.exact_topk(fx.queries, 10)
        for k, rk in ((10, 50), (1, 1), (20, 200)):
            r = gi.search(fx.queries, k, rk, expand_width=1)
            wd, ws, wc, wst = ora.search(fx.queries, k, rk)
            np.testing.assert_array_equal(r.counts, wc)
            same = np.mean([np.array_equal(a, b) for a, b in zip(r.docs, wd)])
            assert same >= 0.9, same
            # returned scores are exact-rerank scores: identical wherever the same doc is returned
            for i in range(len(fx.queries)):
                ref = {int(d): s for d, s in zip(wd[i], ws[i]) if d >= 0}
                for d, s in zip(r.docs[i], r.scores[i]):
                    if int(d) in ref:
                        assert s == ref[int(d)]
            if k == 10:
                assert recall(r.docs, gt) >= recall(wd, gt) - 0.01
                # ties at the list boundary change the expansion count by a few nodes at most
                assert abs(r.stats[:, 1].mean() - wst[:, 1].mean()) <= 0.05 * wst[:, 1].mean() + 1
                assert r.stats[:, 0].mean() <= 1.2 * wst[:, 0].mean() + 4  # re-scored nodes (filter evictions) count as visits


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("width", [0, 2, 4])
def test_wide_expansion_recall_parity(jv, request, name, width):
    fx = request.getfixturevalue(name)
    ora = fx.oracle_index()  # the reference path: fp32 table
    with fx.gpu_index(jv, flags=jv.native.FLAG_LUT_U8) as gi:
        gt, _, _ = gi.exact_topk(fx.queries, 10)
        r = gi.search(fx.queries, 10, 50, expand_width=width)
        wd, ws, wc, wst = ora.search(fx.queries, 10, 50)
        assert recall(r.docs, gt) >= recall(wd, gt) - 0.015
        np.testing.assert_array_equal(r.counts, wc)
        assert (r.stats[:, 3] == 50).all()  # everything in the approximate list is reranked (floor 0)
        r2 = gi.search(fx.queries, 10, 50, expand_width=width)
        np.testing.assert_array_equal(r.docs, r2.docs)  # integer sums: run-to-run deterministic results
        np.testing.assert_array_equal(r.scores, r2.scores)


def test_large_rerank_k_and_chunked_batches(jv, fx_l2, monkeypatch):
    fx = fx_l2
    ora = fx.oracle_index()
    with fx.gpu_index(jv, flags=jv.native.FLAG_LUT_U8) as gi:
        gt, _, _ = gi.exact_topk(fx.queries, 100)
        r = gi.search(fx.queries, 100, 500)  # cfg 5 flavour: k = 100, rerankK = 500
        wd = ora.search(fx.queries, 100, 500)[0]
        assert recall(r.docs, gt) >= recall(wd, gt) - 0.005
        assert (r.counts == 100).all()
        a = gi.search(fx.queries, 10, 50)
        monkeypatch.setenv("JVGPU_Q8_CHUNK", "7")  # table staging buffer of 7 queries: 10 chunks for 64 queries
        gi.refresh_knobs()
        b = gi.search(fx.queries, 10, 50)
        monkeypatch.delenv("JVGPU_Q8_CHUNK")
        gi.refresh_knobs()
        np.testing.assert_array_equal(a.docs, b.docs)
        np.testing.assert_array_equal(a.scores, b.scores)
        np.testing.assert_array_equal(a.stats[:, 1:], b.stats[:, 1:])


def test_thresholds_and_unsupported_shapes_fall_back(jv, fx_l2):
    fx = fx_l2
    ora = fx.oracle_index(adc_order=32)
    bits = O.make_accept_bits(np.random.default_rng(1).random(fx.base.shape[0]) < 0.3)
    with fx.gpu_index(jv, flags=jv.native.FLAG_LUT_U8) as gi:  # explicit strict order: reference kernel, fp32 table
        r = gi.search(fx.queries, 10, 50, accept_bits=bits, expand_width=-1)
        wd, ws, wc, _ = ora.search(fx.queries, 10, 50, accept_bits=bits)
        np.testing.assert_array_equal(r.docs, wd)
        np.testing.assert_array_equal(r.counts, wc)
        r = gi.search(fx.queries, 10, 50, threshold=0.05)       # range threshold: strict kernel
        wd, ws, wc, _ = ora.search(fx.queries, 10, 50, threshold=0.05)
        np.testing.assert_array_equal(r.docs, wd)
    # non-uniform sub-vectors (dim 30, M = 4 -> sizes 8,8,7,7): no 8-bit path, the flag is ignored
    base, q = clustered(3000, 30, 32, seed=5)
    odd = make_fixture(O.SIM_EUCLIDEAN, base, q, max_degree=16, pq_m=4)
    with odd.gpu_index(jv, flags=jv.native.FLAG_LUT_U8) as gi:
        r = gi.search(q, 10, 50, expand_width=1)
        wd = odd.oracle_index(adc_order=1).search(q, 10, 50)[0]
        np.testing.assert_array_equal(r.docs, wd)
        with pytest.raises(ValueError):
            gi.pq_lut_q8(q)


@pytest.mark.parametrize("name", ["fx_dot", "fx_l2", "fx_cos8"])
@pytest.mark.parametrize("selectivity", [0.1, 0.5])
def test_filtered_queries_on_the_8bit_path(jv, request, name, selectivity):
    """Accept bits (a9): rejected nodes are traversed but never returned.  The production path keeps accepted and rejected
    nodes in one list of 8 * rerankK entries; gate = recall parity with the reference loop at equal rerankK."""
    fx = request.getfixturevalue(name)
    n = fx.base.shape[0]
    rng = np.random.default_rng(7)
    mask = rng.random(n) < selectivity
    bits = O.make_accept_bits(mask)
    ora = fx.oracle_index()
    with fx.gpu_index(jv, flags=jv.native.FLAG_LUT_U8) as gi:
        gt, _, gc = gi.exact_topk(fx.queries, 10, accept_bits=bits)
        wd, ws, wc, wst = ora.search(fx.queries, 10, 50, accept_bits=bits)
        for width in (1, 0):
            r = gi.search(fx.queries, 10, 50, accept_bits=bits, expand_width=width)
            assert mask[r.docs[r.docs >= 0]].all()                       # never a rejected doc
            np.testing.assert_array_equal(r.counts, wc)
            assert recall(r.docs, gt) >= recall(wd, gt) - 0.02
            for i in range(len(fx.queries)):                             # exact-rerank scores wherever the doc matches
                ref = {int(d): s for d, s in zip(wd[i], ws[i]) if d >= 0}
                for d, s in zip(r.docs[i], r.scores[i]):
                    if int(d) in ref:
                        assert s == ref[int(d)]
        # per-query bitsets: query i keeps only docs with doc % 3 == i % 3
        per = np.stack([O.make_accept_bits((np.arange(n) % 3) == (i % 3)) for i in range(len(fx.queries))])
        r = gi.search(fx.queries, 10, 50, accept_bits=per)
        for i in range(len(fx.queries)):
            d = r.docs[i][r.docs[i] >= 0]
            assert (d % 3 == i % 3).all() and len(d) == 10


def test_filtered_queries_with_deleted_docs_and_empty_filters(jv):
    base, q = clustered(3000, 64, 32, seed=21, normalize=True)
    n = base.shape[0]
    o2d = np.arange(n, dtype=np.int32) * 2                               # ordinal != docId
    o2d[::7] = -1                                                        # deleted / no-vector ordinals
    fx = make_fixture(O.SIM_DOT, base, q, max_degree=16, pq_m=16, ord_to_doc=o2d, max_doc=2 * n)
    mask = np.zeros(2 * n, bool)
    mask[::4] = True                                                     # docs 0, 4, 8, .. = ordinals 0, 2, 4, ..
    bits = O.make_accept_bits(mask)
    with fx.gpu_index(jv, flags=jv.native.FLAG_LUT_U8) as gi:
        r = gi.search(q, 10, 50, accept_bits=bits)
        wd, ws, wc, _ = fx.oracle_index().search(q, 10, 50, accept_bits=bits)
        got = r.docs[r.docs >= 0]
        assert (got % 4 == 0).all() and not np.isin(got // 2, np.arange(0, n, 7)).any()
        np.testing.assert_array_equal(r.counts, wc)
        gt, _, _ = gi.exact_topk(q, 10, accept_bits=bits)
        assert recall(r.docs, gt) >= recall(wd, gt) - 0.02
        none = gi.search(q, 10, 50, accept_bits=O.make_accept_bits(np.zeros(2 * n, bool)))
        assert (none.counts == 0).all() and (none.docs == -1).all()


def test_single_query_and_tiny_graph(jv, fx_dot):
    with fx_dot.gpu_index(jv, flags=jv.native.FLAG_LUT_U8) as gi:
        one = gi.search(fx_dot.queries[:1], 10, 50)
        many = gi.search(fx_dot.queries, 10, 50)
        np.testing.assert_array_equal(one.docs[0], many.docs[0])
    base, q = clustered(300, 16, 4, seed=3)  # K = 256 needs n >= 256
    small = make_fixture(O.SIM_EUCLIDEAN, base, q, max_degree=8, pq_m=8)
    with small.gpu_index(jv, flags=jv.native.FLAG_LUT_U8) as gi:
        r = gi.search(q, 10, 400)  # rerankK larger than the graph: every node ends up in the list
        wd, ws, wc, _ = small.oracle_index(adc_order=-8).search(q, 10, 400)
        np.testing.assert_array_equal(r.docs, wd)
        np.testing.assert_array_equal(r.counts, wc)


def test_large_host_batches_are_pipelined_in_chunks(jv, fx_dot):
    """jv_search_batch splits staged batches of >= 4096 queries into chunks whose H2D copy overlaps the previous chunk's
    kernels; results must be those of the small batch, query by query."""
    fx = fx_dot
    reps = 4200 // len(fx.queries) + 1
    big = np.tile(fx.queries, (reps, 1))[:4200]
    for flags in (jv.native.FLAG_LUT_U8, 0):
        with fx.gpu_index(jv, flags=flags) as gi:
            small = gi.search(fx.queries, 10, 50)
            r = gi.search(big, 10, 50)
            nq = len(fx.queries)
            for t in range(0, 4200, nq):
                m = min(nq, 4200 - t)
                np.testing.assert_array_equal(r.docs[t:t + m], small.docs[:m])
                np.testing.assert_array_equal(r.scores[t:t + m], small.scores[:m])
                np.testing.assert_array_equal(r.counts[t:t + m], small.counts[:m])
                np.testing.assert_array_equal(r.stats[t:t + m, 1:], small.stats[:m, 1:])


def test_vectors_in_pinned_host_memory_and_wide_graphs(jv, fx_dot, monkeypatch):
    """cfg 5 flavour: fp32 rerank vectors stay in pinned host memory (K3 gathers them over PCIe) while the 8-bit table
    traversal runs from HBM; and a graph with R = 64 (two adjacency chunks per candidate, <= 256 queued survivors)."""
    fx = fx_dot
    with fx.gpu_index(jv, flags=jv.native.FLAG_LUT_U8) as a, \
            fx.gpu_index(jv, flags=jv.native.FLAG_LUT_U8 | jv.native.FLAG_NO_VECTORS_ON_DEVICE) as b:
        ra, rb = a.search(fx.queries, 10, 50), b.search(fx.queries, 10, 50)
        np.testing.assert_array_equal(ra.docs, rb.docs)
        np.testing.assert_array_equal(ra.scores, rb.scores)
        # large batches gather the rows they need ONCE from host memory into HBM (bitmap over the ordinals, popcount ranks, dense
        # staging array, row map in the rerank kernel): forced here for a small batch, same ids / score bits / counters; repeated
        # queries make most rows duplicates
        many = np.tile(fx.queries, (3, 1))
        for knob in ("1", "0"):
            monkeypatch.setenv("JVGPU_RERANK_DEDUPE", knob)
            b.refresh_knobs()
            rk = b.search(many, 10, 50)
            for t in range(3):
                sl = slice(t * len(fx.queries), (t + 1) * len(fx.queries))
                np.testing.assert_array_equal(rk.docs[sl], ra.docs)
                np.testing.assert_array_equal(rk.scores[sl], ra.scores)
                np.testing.assert_array_equal(rk.stats[sl, 1:], ra.stats[:, 1:])
            rt = b.search(fx.queries[:7], 20, 200, rerank_floor=float(np.median(ra.scores)))  # rows below the floor are marked but not read
            monkeypatch.setenv("JVGPU_RERANK_DEDUPE", "0")
            b.refresh_knobs()
            r0 = b.search(fx.queries[:7], 20, 200, rerank_floor=float(np.median(ra.scores)))
            np.testing.assert_array_equal(rt.docs, r0.docs)
            np.testing.assert_array_equal(rt.scores, r0.scores)
        monkeypatch.delenv("JVGPU_RERANK_DEDUPE")
        b.refresh_knobs()
    base, q = clustered(4000, 64, 40, seed=12, normalize=True)
    wide = make_fixture(O.SIM_DOT, base, q, max_degree=64, pq_m=16)
    with wide.gpu_index(jv, flags=jv.native.FLAG_LUT_U8) as gi:
        gt, _, _ = gi.exact_topk(q, 10)
        r = gi.search(q, 10, 50)
        wd = wide.oracle_index().search(q, 10, 50)[0]
        assert recall(r.docs, gt) >= recall(wd, gt) - 0.015
        r1 = gi.search(q, 10, 50, expand_width=1)
        w1 = wide.oracle_index(adc_order=-8).search(q, 10, 50)[0]
        assert np.mean([np.array_equal(x, y) for x, y in zip(r1.docs, w1)]) >= 0.9


@pytest.mark.parametrize("m,sub", [(96, 2), (128, 2), (160, 2), (256, 2), (64, 4), (96, 8)])
def test_every_table_size_instantiation(jv, m, sub):
    """One shape per traversal instantiation that the named fixtures do not reach — NJ = ceil(M / 32) in {3, 4, 5 (generic
    code path), 8} and the M <= 64 shape the full-size workloads use — each compiled for its own CTAs-per-SM budget
    (q8_min_ctas): tables bit-exact, width 1 follows the oracle's order, default width keeps the recall."""
    sim = O.SIM_DOT if m != 128 else O.SIM_EUCLIDEAN
    base, q = clustered(2500, m * sub, 48, seed=100 + m, normalize=sim == O.SIM_DOT)
    fx = make_fixture(sim, base, q, max_degree=16, pq_m=m)
    ora = fx.oracle_index(adc_order=-8)
    with fx.gpu_index(jv, flags=jv.native.FLAG_LUT_U8) as gi:
        q8, prm = gi.pq_lut_q8(q[:9])
        w8, wprm = ora.lut_q8(q[:9])
        np.testing.assert_array_equal(q8, w8)
        np.testing.assert_array_equal(prm.view(np.uint32), wprm.view(np.uint32))
        gt, _, _ = gi.exact_topk(q, 10)
        wd = ora.search(q, 10, 50)[0]
        r1 = gi.search(q, 10, 50, expand_width=1)
        assert np.mean([np.array_equal(a, b) for a, b in zip(r1.docs, wd)]) >= 0.9
        r4 = gi.search(q, 10, 50)
        assert recall(r4.docs, gt) >= recall(wd, gt) - 0.01
        bits = jv.make_accept_bits(np.random.default_rng(m).random(2500) < 0.5)      # FILT instantiation of the same shape
        rf = gi.search(q, 10, 50, accept_bits=bits)
        gf, _, _ = gi.exact_topk(q, 10, bits)
        assert recall(rf.docs, gf) >= 0.95


@pytest.fixture(scope="module")
def fx_dot192_wide():  # M = 192, sub-dim 2, R = 64 (two adjacency chunks per row): the manager / expander / scorer kernel at its limits
    base, q = embedded(12000, 384, 48, seed=45)
    return make_fixture(O.SIM_DOT, base, q, max_degree=64, pq_m=192)


def test_beam_kernel_full_list_wide_rows_and_filter_evictions(jv, fx_dot192_wide):
    """rerankK = 64 fills the register-held list, R = 64 makes every step read two adjacency chunks per candidate, and ~3 000 visits
    against the 2 048-tag visited filter of the M = 192 shape evict entries all the time: re-scored list members and nodes scored
    twice in one step go through the merge's second counting pass.  The result must stay a set (no document twice), follow the
    oracle's 8-bit order at width 1, keep the recall gate at the default width and be identical from run to run."""
    fx = fx_dot192_wide
    ora = fx.oracle_index(adc_order=-8)
    with fx.gpu_index(jv, flags=jv.native.FLAG_LUT_U8) as gi:
        gt, _, _ = gi.exact_topk(fx.queries, 10)
        wd, ws, wc, wst = ora.search(fx.queries, 10, 64)
        for width in (1, 0, 2):
            r = gi.search(fx.queries, 10, 64, expand_width=width)
            # two adjacency chunks per candidate: steps of <= 2 candidates run on q8_beam_kernel, the default width on the synchronous kernel
            assert r.timing["traversal_kernel"] == (3 if width else 2)
            np.testing.assert_array_equal(r.counts, wc)
            for row in r.docs:
                live = row[row >= 0]
                assert len(set(live.tolist())) == len(live)
            assert (r.stats[:, 3] == 64).all()
            assert recall(r.docs, gt) >= recall(wd, gt) - 0.015
            if width == 1:
                same = np.mean([np.array_equal(a, b) for a, b in zip(r.docs, wd)])
                assert same >= 0.9, same
                assert r.stats[:, 0].mean() >= wst[:, 0].mean()  # evictions only ever add visits
            r2 = gi.search(fx.queries, 10, 64, expand_width=width)
            np.testing.assert_array_equal(r.docs, r2.docs)
            np.testing.assert_array_equal(r.scores, r2.scores)
            # expansions follow the list, which is deterministic; the visited counter also counts re-scored nodes, and which of two
            # same-set tags written by one adjacency chunk survives in the visited filter is up to the hardware
            np.testing.assert_array_equal(r.stats[:, 1:], r2.stats[:, 1:])
            assert np.abs(r.stats[:, 0].astype(np.int64) - r2.stats[:, 0]).max() <= 0.02 * r.stats[:, 0].max() + 8


def test_small_batches_take_the_latency_shapes_and_return_the_same_results(jv, fx_dot192):
    """One query at a time (table build split over 8 CTAs, 16-warp rerank, one packed result copy) against the same queries inside a
    batch large enough for the throughput shapes: ids, score bits and counters must be identical."""
    fx = fx_dot192
    with fx.gpu_index(jv, flags=jv.native.FLAG_LUT_U8) as gi:
        big = np.tile(fx.queries, (5, 1))  # 320 queries: more than one per SM
        rb = gi.search(big, 10, 50)
        for nq in (1, 3, len(fx.queries)):
            rs = gi.search(fx.queries[:nq], 10, 50)
            np.testing.assert_array_equal(rs.docs, rb.docs[:nq])
            np.testing.assert_array_equal(rs.scores, rb.scores[:nq])
            np.testing.assert_array_equal(rs.counts, rb.counts[:nq])
            np.testing.assert_array_equal(rs.stats[:, 1:], rb.stats[:nq, 1:])
        q8_one, prm_one = gi.pq_lut_q8(fx.queries[:1])  # 8 code-range slices
        q8_all, prm_all = gi.pq_lut_q8(big)             # 2 slices
        np.testing.assert_array_equal(q8_one[0], q8_all[0])
        np.testing.assert_array_equal(prm_one[0].view(np.uint32), prm_all[0].view(np.uint32))
