"""Parity of the CUDA path against the CPU oracle, through the C-ABI (libjvgpu.so).  -m gpu.

Gates (north star / BASELINE.md §4):
  * PQ codes                      bit-exact
  * ADC table (LUT)               bit-exact (same fmaf order); ADC scores bit-exact vs oracle order "warp32"
  * brute-force / rerank top-k    ids bit-exact (ties -> lower docId), scores <= 1e-5 relative
  * graph search                  fp32 paths: ids, scores AND visited/expanded counters identical to the oracle;
                                  fp16-table mode: recall@k within 0.005 of the oracle's
"""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import clustered, lucene_score, make_fixture, recall

pytestmark = pytest.mark.gpu

SCORE_RTOL = 1e-5  # north star: "scores within 1e-5 relative"
STRICT = -1        # jv_search_params.expand_width: the strict (reference-order) kernel


@pytest.fixture(scope="module")
def fx_pq_dot():
    base, q = clustered(6000, 64, 64, seed=11, normalize=True)
    return make_fixture(O.SIM_DOT, base, q, max_degree=16, pq_m=16)


@pytest.fixture(scope="module")
def fx_pq_l2():
    base, q = clustered(5000, 96, 48, seed=12)
    return make_fixture(O.SIM_EUCLIDEAN, base, q, max_degree=32, pq_m=48)


@pytest.fixture(scope="module")
def fx_pq_cos():
    base, q = clustered(4000, 128, 48, seed=13)
    return make_fixture(O.SIM_COSINE, base, q, max_degree=16, pq_m=32)


@pytest.fixture(scope="module")
def fx_exact_cos():
    # config 1 shape (10k x 128 U[0,1) cosine, M=16, beamWidth=100) at reference seeds 42/43, 200 queries
    base = O.java_random_vectors(10000, 128, 42)
    q = O.java_random_vectors(200, 128, 43)
    return make_fixture(O.SIM_COSINE, base, q, max_degree=16)


def assert_same_results(gpu, ora_docs, ora_scores, ora_counts):
    np.testing.assert_array_equal(gpu.counts, ora_counts)
    np.testing.assert_array_equal(gpu.docs, ora_docs)
    np.testing.assert_allclose(gpu.scores, ora_scores, rtol=SCORE_RTOL, atol=0)


# ---- K6 ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,m,k,center", [(64, 16, 256, False), (96, 48, 256, True), (128, 64, 256, False), (768, 192, 256, False),
                                            (16, 16, 64, True), (100, 7, 33, True), (24, 24, 256, False), (1536, 192, 256, False)])
def test_pq_encode_bit_exact(jv, dim, m, k, center):
    rng = np.random.default_rng(dim * 1000 + m)
    n = 3000 if dim <= 128 else 700
    x = rng.standard_normal((n, dim)).astype(np.float32)
    cb, g = O.pq_train(x[:600], m, k, center, iters=2, seed=3)
    want = O.pq_encode(x, m, k, cb, g)
    got = jv.pq_encode(x, m, k, cb, g)
    np.testing.assert_array_equal(got, want)


def test_pq_encode_ties_and_edges(jv):
    cb = np.array([[0.0, 0.0], [1.0, 1.0], [1.0, 1.0], [5.0, 5.0]], np.float32).reshape(-1)
    x = np.array([[1.0, 1.0], [0.4, 0.4], [9.0, 9.0], [0.5, 0.5]], np.float32)
    np.testing.assert_array_equal(jv.pq_encode(x, 1, 4, cb), O.pq_encode(x, 1, 4, cb))
    assert jv.pq_encode(np.zeros((0, 2), np.float32), 1, 4, cb).shape == (0, 1)  # empty flush
    one = jv.pq_encode(x[:1], 1, 4, cb)
    assert one.shape == (1, 1) and one[0, 0] == 1


def test_pq_encode_idempotent_on_centroids(jv):
    # size-independent property: encoding a centroid returns (the first copy of) itself
    rng = np.random.default_rng(1)
    dim, m, k = 32, 8, 256
    cb = rng.standard_normal(k * dim).astype(np.float32)
    cbs = cb.reshape(m, k, dim // m)
    x = np.concatenate([cbs[j][np.arange(k)] for j in range(m)], axis=1)  # row c = centroid c in every subspace
    codes = jv.pq_encode(x, m, k, cb)
    np.testing.assert_array_equal(codes, np.tile(np.arange(k, dtype=np.uint8)[:, None], (1, m)))


# ---- K1 / a4 ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["fx_pq_dot", "fx_pq_l2", "fx_pq_cos"])
def test_lut_and_adc_bit_exact(jv, request, name):
    fx = request.getfixturevalue(name)
    q = fx.queries[:8]
    with fx.gpu_index(jv) as gi:
        lut = gi.pq_lut(q)
        want = O.pq_lut(fx.sim, fx.base.shape[1], fx.pq_m, fx.pq_k, fx.codebooks, fx.gcent, q)
        np.testing.assert_array_equal(lut, want)
        nodes = np.tile(np.arange(0, 400, dtype=np.int32), (8, 1))
        got = gi.adc_scores(q, nodes)
        ora = fx.oracle_index(adc_order=32).adc_scores(q, nodes)
        np.testing.assert_array_equal(got, ora)
        np.testing.assert_allclose(got, fx.oracle_index(adc_order=0).adc_scores(q, nodes), rtol=SCORE_RTOL)


# ---- K5 ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["fx_pq_dot", "fx_pq_l2", "fx_pq_cos", "fx_exact_cos"])
def test_exact_topk_bit_exact_ids(jv, request, name):
    fx = request.getfixturevalue(name)
    q = fx.queries[:40]
    ora = fx.oracle_index()
    with fx.gpu_index(jv) as gi:
        for k in (1, 10, 100):
            docs, scores, counts = gi.exact_topk(q, k)
            wd, ws, wc = ora.exact_topk(q, k)
            np.testing.assert_array_equal(docs, wd)
            np.testing.assert_array_equal(counts, wc)
            np.testing.assert_allclose(scores, ws, rtol=SCORE_RTOL, atol=0)
            np.testing.assert_array_equal(scores, ws)  # canonical reduction: identical bits


def test_exact_topk_ties_filter_deleted_mip(jv):
    base = np.array([[1.0, 0.0], [0.0, 1.0], [1.0, 0.0], [0.5, 0.5], [1.0, 0.0], [2.0, 0.0]], np.float32)
    o2d = np.array([0, 1, 2, 3, 4, -1], np.int32)  # best vector is deleted
    adj = np.full((6, 2), -1, np.int32)
    with jv.GpuIndex(O.SIM_MIP, base, adj, 0, ord_to_doc=o2d, max_doc=5) as gi:
        docs, scores, counts = gi.exact_topk(np.array([[1.0, 0.0]], np.float32), 4)
        assert list(docs[0]) == [0, 2, 4, 3] and counts[0] == 4
        assert scores[0][0] == 2.0  # MIP: 1 + dot (JVectorVectorScorer.java:43-50)
        mask = np.array([False, True, True, True, False])
        docs, scores, counts = gi.exact_topk(np.array([[1.0, 0.0]], np.float32), 4, jv.make_accept_bits(mask))
        assert list(docs[0]) == [2, 3, 1, -1] and counts[0] == 3


def test_exact_topk_odd_dimension(jv):
    rng = np.random.default_rng(9)
    base = rng.standard_normal((900, 7)).astype(np.float32)
    q = rng.standard_normal((5, 7)).astype(np.float32)
    adj = np.full((900, 2), -1, np.int32)
    for sim in (O.SIM_EUCLIDEAN, O.SIM_DOT, O.SIM_COSINE):
        ora = O.OracleIndex(sim, base, adj, 0)
        with jv.GpuIndex(sim, base, adj, 0) as gi:
            d, s, c = gi.exact_topk(q, 10)
            wd, ws, wc = ora.exact_topk(q, 10)
            np.testing.assert_array_equal(d, wd)
            np.testing.assert_array_equal(s, ws)


# ---- K7 ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("g,nq,k", [(2, 7, 10), (8, 33, 100), (8, 5, 1), (70, 3, 100)])
def test_merge_topk(jv, g, nq, k):
    rng = np.random.default_rng(g * 100 + k)
    docs = rng.permutation(g * nq * k * 2)[: g * nq * k].reshape(g, nq, k).astype(np.int32)
    scores = rng.integers(0, 50, (g, nq, k)).astype(np.float32) / 50.0  # many ties
    order = np.argsort(-scores, axis=2, kind="stable")
    docs, scores = np.take_along_axis(docs, order, 2), np.take_along_axis(scores, order, 2)
    docs[0, 0, k // 2:] = -1  # ragged list
    d, s, c = jv.merge_topk(docs, scores, k)
    wd, ws, wc = O.merge_topk(docs, scores, k)
    np.testing.assert_array_equal(d, wd)
    np.testing.assert_array_equal(s, ws)
    np.testing.assert_array_equal(c, wc)


# ---- K4: exact traversal (config 1) ---------------------------------------------------------------------------
def test_search_exact_identical_to_oracle(jv, fx_exact_cos):
    fx = fx_exact_cos
    ora = fx.oracle_index()
    with fx.gpu_index(jv) as gi:
        for k, rk in ((10, 50), (1, 1), (100, 100)):
            r = gi.search(fx.queries, k, rk, expand_width=STRICT)
            wd, ws, wc, wst = ora.search(fx.queries, k, rk)
            assert_same_results(r, wd, ws, wc)
            np.testing.assert_array_equal(r.stats, wst)  # visited / expanded / reranked
        gt, _, _ = gi.exact_topk(fx.queries, 10)
        r = gi.search(fx.queries, 10, 50, expand_width=STRICT)
        assert abs(recall(r.docs, gt) - recall(ora.search(fx.queries, 10, 50)[0], gt)) <= 0.005


@pytest.mark.parametrize("sim", [O.SIM_EUCLIDEAN, O.SIM_DOT, O.SIM_MIP])
def test_search_exact_other_similarities(jv, sim):
    base, q = clustered(3000, 48, 32, seed=21 + sim, normalize=(sim != O.SIM_EUCLIDEAN))
    fx = make_fixture(sim, base, q, max_degree=16)
    ora = fx.oracle_index()
    with fx.gpu_index(jv) as gi:
        r = gi.search(q, 10, 50, expand_width=STRICT)
        wd, ws, wc, wst = ora.search(q, 10, 50)
        assert_same_results(r, wd, ws, wc)
        np.testing.assert_array_equal(r.stats, wst)
        if sim == O.SIM_MIP:  # un-quantised MIP scores are Lucene's 1 + dot
            assert abs(r.scores[0, 0] - lucene_score(O.SIM_MIP, q[0], base[r.docs[0, 0]])) < 1e-4


# ---- K1+K2+K3: PQ traversal + rerank --------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["fx_pq_dot", "fx_pq_l2", "fx_pq_cos"])
def test_search_pq_identical_to_oracle(jv, request, name):
    fx = request.getfixturevalue(name)
    ora = fx.oracle_index(adc_order=32)
    with fx.gpu_index(jv) as gi:
        for k, rk in ((10, 50), (5, 5), (20, 200)):
            r = gi.search(fx.queries, k, rk, expand_width=STRICT)
            wd, ws, wc, wst = ora.search(fx.queries, k, rk)
            assert_same_results(r, wd, ws, wc)
            np.testing.assert_array_equal(r.stats, wst)
        # recall gate against exact ground truth, vs the oracle in its default summation order
        gt, _, _ = gi.exact_topk(fx.queries, 10)
        r = gi.search(fx.queries, 10, 50, expand_width=STRICT)
        ora0 = fx.oracle_index(adc_order=0)
        assert abs(recall(r.docs, gt) - recall(ora0.search(fx.queries, 10, 50)[0], gt)) <= 0.005
        assert gi.visited_overflows() == 0


@pytest.mark.parametrize("name", ["fx_pq_dot", "fx_pq_l2", "fx_pq_cos"])
def test_search_pq_fp16_table_recall_parity(jv, request, name):
    fx = request.getfixturevalue(name)
    ora = fx.oracle_index()
    with fx.gpu_index(jv, flags=jv.native.FLAG_LUT_F16) as gi:
        gt, _, _ = gi.exact_topk(fx.queries, 10)
        r = gi.search(fx.queries, 10, 50)
        wd, ws, wc, _ = ora.search(fx.queries, 10, 50)
        assert abs(recall(r.docs, gt) - recall(wd, gt)) <= 0.005
        # final scores come from the exact rerank: wherever the same doc is returned the score is identical
        for i in range(len(fx.queries)):
            common = {int(d): s for d, s in zip(wd[i], ws[i]) if d >= 0}
            for d, s in zip(r.docs[i], r.scores[i]):
                if int(d) in common:
                    assert s == common[int(d)]


def test_search_pq_filter_threshold_floor(jv, fx_pq_l2):
    fx = fx_pq_l2
    ora = fx.oracle_index(adc_order=32)
    rng = np.random.default_rng(77)
    mask = rng.random(fx.base.shape[0]) < 0.10  # 10 % selectivity (config 4 flavour)
    bits = O.make_accept_bits(mask)
    per_query = np.stack([O.make_accept_bits(rng.random(fx.base.shape[0]) < 0.5) for _ in range(len(fx.queries))])
    with fx.gpu_index(jv) as gi:
        r = gi.search(fx.queries, 10, 50, accept_bits=bits)
        wd, ws, wc, wst = ora.search(fx.queries, 10, 50, accept_bits=bits)
        assert_same_results(r, wd, ws, wc)
        # with a filter the device keeps a BOUNDED candidate array (8 * rerankK entries) where the reference heap is
        # unbounded: results are identical here, but hopeless candidates are dropped instead of expanded
        assert (r.stats[:, 1] <= wst[:, 1]).all() and (r.stats[:, 3] == wst[:, 3]).all()
        assert mask[r.docs[r.docs >= 0]].all()
        r = gi.search(fx.queries, 10, 50, accept_bits=per_query)  # one bitset per query
        wd, ws, wc, wst = ora.search(fx.queries, 10, 50, accept_bits=per_query)
        assert_same_results(r, wd, ws, wc)
        # threshold + rerank floor
        thr = float(np.median(ws[:, 0])) * 0.9
        r = gi.search(fx.queries, 10, 50, threshold=thr, rerank_floor=thr)
        wd, ws, wc, wst = ora.search(fx.queries, 10, 50, threshold=thr, rerank_floor=thr)
        assert_same_results(r, wd, ws, wc)
        # like a filter, a threshold can keep the result heap from filling: bounded candidates, fewer expansions
        assert (r.stats[:, 1] <= wst[:, 1]).all() and (r.stats[:, 3] == wst[:, 3]).all()


def test_search_ordinal_map_deleted_and_batch_of_one(jv):
    base, q = clustered(2500, 32, 16, seed=31)
    rng = np.random.default_rng(5)
    o2d = rng.permutation(4000)[:2500].astype(np.int32)
    o2d[rng.random(2500) < 0.05] = -1
    fx = make_fixture(O.SIM_EUCLIDEAN, base, q, max_degree=16, pq_m=16, ord_to_doc=o2d, max_doc=4000)
    ora = fx.oracle_index(adc_order=32)
    with fx.gpu_index(jv) as gi:
        r = gi.search(q, 10, 50, expand_width=STRICT)
        wd, ws, wc, wst = ora.search(q, 10, 50)
        assert_same_results(r, wd, ws, wc)
        one = gi.search(q[3], 10, 50, expand_width=STRICT)  # batch of 1 stays legal (reference API is one query per call)
        np.testing.assert_array_equal(one.docs[0], wd[3])
        with pytest.raises(ValueError):
            gi.search(q, 10, 5)  # rerankK < topK
        with pytest.raises(ValueError):
            gi.search(np.zeros((1, 31), np.float32), 10, 50)


def test_reference_analytic_cases_through_reader(jv):
    """KNNJVectorTests.java:73-130, 212-273, 1218-1271 replayed through the reader mirror on the GPU."""
    V = jv.VectorSimilarityFunction

    def run(sim, base, target, k, accept=None):
        adj, entry = O.graph_build(base, sim.jvector_ord, 32, 100)
        seg = jv.Segment(max_doc=len(base))
        seg.fields["test_field"] = jv.FieldData(sim, base, adj, entry, jv.GraphNodeIdToDocMap(np.arange(len(base)), len(base)))
        reader = jv.JVectorReader(seg)
        col = jv.JVectorKnnCollector(jv.TopKnnCollector(k), 0.0, 0.0, 5)
        reader.search("test_field", target, col, accept)
        td = col.top_docs()
        reader.close()
        return td, col

    base = np.array([[0.0, 1.0 / i] for i in range(1, 11)], np.float32)
    td, col = run(V.EUCLIDEAN, base, [0.0, 0.0], 3)
    assert [sd.doc for sd in td] == [9, 8, 7] and col.visited_count() > 0
    for sd in td:
        assert abs(sd.score - lucene_score(O.SIM_EUCLIDEAN, [0.0, 0.0], base[sd.doc])) < 1e-3
    base = np.array([[1.0 / i, 0.0] for i in range(1, 11)], np.float32)
    td, _ = run(V.MAXIMUM_INNER_PRODUCT, base, [1.0, 0.0], 3, accept=np.array([(i % 2 == 0) for i in range(1, 11)]))
    assert [sd.doc for sd in td] == [1, 3, 5]
    for sd in td:
        assert abs(sd.score - lucene_score(O.SIM_MIP, [1.0, 0.0], base[sd.doc])) < 1e-3
    base = np.array([[1.0 + i, 2.0 * i] for i in range(1, 11)], np.float32)
    td, _ = run(V.COSINE, base, [1.0, 1.0], 3)
    assert [sd.doc for sd in td] == [0, 1, 2]


def test_concurrent_queries_share_one_index(jv, fx_pq_dot):
    """KNNJVectorTests.java:982-1028: 10 threads x 100 queries on one reader."""
    import threading
    fx = fx_pq_dot
    want = fx.oracle_index(adc_order=32).search(fx.queries, 10, 50)[0]
    errors = []
    with fx.gpu_index(jv) as gi:
        def worker(t):
            try:
                for it in range(25):
                    i = (t * 7 + it) % len(fx.queries)
                    r = gi.search(fx.queries[i], 10, 50, expand_width=STRICT)
                    if not np.array_equal(r.docs[0], want[i]):
                        errors.append((t, i))
            except Exception as e:  # noqa
                errors.append(repr(e))
        ths = [threading.Thread(target=worker, args=(t,)) for t in range(10)]
        [t.start() for t in ths]
        [t.join() for t in ths]
    assert not errors


# ---- "next" rows: device-side PQ training (8f-2) and Vamana construction (8f-3) ---------------------------------
@pytest.mark.parametrize("n,dim,m,k,center", [(3000, 32, 8, 256, True), (2000, 24, 24, 64, False), (1500, 100, 7, 32, True),
                                              (5000, 64, 16, 256, False)])
def test_pq_train_matches_oracle(jv, n, dim, m, k, center):
    rng = np.random.default_rng(n + dim)
    x = (rng.standard_normal((n, dim)) + rng.integers(0, 4, (n, 1))).astype(np.float32)
    cb, g = jv.pq_train(x, m, k, center, iters=6, seed=5)
    wcb, wg = O.pq_train(x, m, k, center, iters=6, seed=5)
    if center:
        np.testing.assert_array_equal(g, wg)
    np.testing.assert_array_equal(cb, wcb)


@pytest.mark.parametrize("sim,n,dim,R", [(O.SIM_EUCLIDEAN, 3000, 32, 16), (O.SIM_COSINE, 2000, 48, 32), (O.SIM_DOT, 2500, 64, 16),
                                         (O.SIM_EUCLIDEAN, 300, 128, 32), (O.SIM_EUCLIDEAN, 1, 8, 4), (O.SIM_EUCLIDEAN, 2, 8, 4),
                                         # batches of 400 nodes: multi-chunk edge sort + contended back-link targets
                                         (O.SIM_EUCLIDEAN, 20000, 16, 16), (O.SIM_DOT, 12000, 24, 32)])
def test_graph_build_matches_oracle(jv, sim, n, dim, R):
    base, _ = clustered(n, dim, 1, seed=100 + n, normalize=(sim != O.SIM_EUCLIDEAN))
    adj, entry = jv.graph_build(base, sim, R, 100, 1.2, 1.2)
    wadj, wentry = O.graph_build(base, sim, R, 100, 1.2, 1.2)
    assert entry == wentry
    np.testing.assert_array_equal(adj, wadj)


def test_writer_reader_pq_recall_reference_seed(jv):
    """KNNJVectorTests.java:1358-1403 through the writer/reader mirrors, everything on the GPU:
    1024 x 16 Random(1) vectors, EUCLIDEAN, PQ (n >= minBatch 1024), k = 50, overquery 5 -> recall 1.0 +- 0.05."""
    V = jv.VectorSimilarityFunction
    dim, n, k = 16, 1024, 50
    vectors = O.java_random_vectors(n, dim, 1)
    w = jv.JVectorWriter(max_conn=32, beam_width=100, min_batch_size_for_quantization=1024)
    w.add_field("test_field", V.EUCLIDEAN)
    for i in range(n):
        w.add_value("test_field", i, vectors[i])
    seg = w.flush(n)
    assert seg.fields["test_field"].pq_codes is not None and seg.fields["test_field"].pq_m == 16
    reader = jv.JVectorReader(seg)
    target = np.zeros(dim, np.float32)
    col = jv.JVectorKnnCollector(jv.TopKnnCollector(k), 0.0, 0.0, 5)
    reader.search("test_field", target, col)
    got = {sd.doc for sd in col.top_docs()}
    gt = {sd.doc for sd in reader.exact_search("test_field", target, k)}
    assert len(got) == k and len(got & gt) / k >= 0.95
    # below the quantisation threshold the segment stays full precision (JVectorWriter.java:267-279)
    w2 = jv.JVectorWriter(min_batch_size_for_quantization=1024)
    w2.add_field("f", V.EUCLIDEAN)
    for i in range(100):
        w2.add_value("f", i, vectors[i])
    seg2 = w2.flush(100)
    assert seg2.fields["f"].pq_codes is None
    r2 = jv.JVectorReader(seg2)
    q = jv.JVectorKnnFloatVectorQuery("f", vectors[5], 10)
    top = q.search([r2])
    assert top[0].doc == 5 and abs(top[0].score - 1.0) < 1e-6
    reader.close()
    r2.close()
