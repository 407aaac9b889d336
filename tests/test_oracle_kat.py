"""Pins the CPU oracle against every known-answer test the reference holds for the hot path
(SURVEY 8c).  These are the ONLY links back to the real reference (jVector 4.0.0-rc.9 cannot run
here): analytic top-k ids + scores, score formulas, seeded recall floors, the Java LCG, the PQ
subspace table and the doc-map semantics.  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import clustered, lucene_score, make_fixture, recall


def _search_ids(fx, q, k, over=5, accept=None):
    ix = fx.oracle_index()
    bits = None if accept is None else O.make_accept_bits(accept)
    docs, scores, counts, stats = ix.search(np.asarray(q, np.float32)[None, :], k, k * over, accept_bits=bits)
    return docs[0], scores[0], counts[0], stats[0]


# ---- java.util.Random: TestUtils.java:108-124 ---------------------------------------------------------
def test_java_random_known_values():
    # documented outputs of new Random(1).nextFloat() / new Random(42).nextFloat()
    np.testing.assert_allclose(O.java_random_floats(1, 3), [0.7308782, 0.100473166, 0.4100808], rtol=0, atol=1e-7)
    np.testing.assert_allclose(O.java_random_floats(42, 1), [0.7275636], rtol=0, atol=1e-6)
    v = O.java_random_vectors(4, 5, 7)
    assert v.shape == (4, 5) and (v >= 0).all() and (v < 1).all()


# ---- KNNJVectorTests.java:73-130 ----------------------------------------------------------------------
def test_simple_case_euclidean():
    base = np.array([[0.0, 1.0 / i] for i in range(1, 11)], np.float32)
    fx = make_fixture(O.SIM_EUCLIDEAN, base, base[:1], max_degree=32)
    target = [0.0, 0.0]
    docs, scores, cnt, _ = _search_ids(fx, target, 3)
    assert cnt == 3 and list(docs) == [9, 8, 7]
    for d, s in zip(docs, scores):
        assert abs(s - lucene_score(O.SIM_EUCLIDEAN, target, base[d])) < 1e-3


# ---- KNNJVectorTests.java:141-201 (MIP: Lucene score 1+dot on the un-quantised path) -------------------
def test_simple_case_max_inner_product():
    base = np.array([[1.0 / i, 0.0] for i in range(1, 11)], np.float32)
    fx = make_fixture(O.SIM_MIP, base, base[:1], max_degree=32)
    target = [1.0, 0.0]
    docs, scores, cnt, _ = _search_ids(fx, target, 3)
    assert cnt == 3 and list(docs) == [0, 1, 2]
    for d, s in zip(docs, scores):
        assert abs(s - lucene_score(O.SIM_MIP, target, base[d])) < 1e-3


# ---- KNNJVectorTests.java:212-273 (MIP + filter "even" -> docs 1,3,5) ----------------------------------
def test_filter_max_inner_product():
    base = np.array([[1.0 / i, 0.0] for i in range(1, 11)], np.float32)
    fx = make_fixture(O.SIM_MIP, base, base[:1], max_degree=32)
    accept = np.array([(i % 2 == 0) for i in range(1, 11)])  # doc d holds i = d+1
    docs, scores, cnt, _ = _search_ids(fx, [1.0, 0.0], 3, accept=accept)
    assert cnt == 3 and list(docs) == [1, 3, 5]
    for d, s in zip(docs, scores):
        assert abs(s - lucene_score(O.SIM_MIP, [1.0, 0.0], base[d])) < 1e-3


# ---- KNNJVectorTests.java:1218-1271 (cosine) -------------------------------------------------------------
def test_cosine():
    base = np.array([[1.0 + i, 2.0 * i] for i in range(1, 11)], np.float32)
    fx = make_fixture(O.SIM_COSINE, base, base[:1], max_degree=32)
    target = [1.0, 1.0]
    docs, scores, cnt, _ = _search_ids(fx, target, 3)
    assert cnt == 3 and list(docs) == [0, 1, 2]
    for d, s in zip(docs, scores):
        assert abs(s - lucene_score(O.SIM_COSINE, target, base[d])) < 1e-3


# ---- KNNJVectorTests.java:1301-1352 (L2 + filter "odd" -> docs 9,7,5) -----------------------------------
def test_filter_euclidean():
    base = np.array([[0.0, 1.0 / i] for i in range(1, 11)], np.float32)
    fx = make_fixture(O.SIM_EUCLIDEAN, base, base[:1], max_degree=32)
    accept = np.array([(i % 2 == 0) for i in range(1, 11)])  # "even" i -> docs 1,3,5,7,9
    docs, _, cnt, _ = _search_ids(fx, [0.0, 0.0], 3, accept=accept)
    assert cnt == 3 and list(docs) == [9, 7, 5]


# ---- index sort / missing docs: ordinal != docId (KNNJVectorTests.java:279-336, 343-417) ---------------
def test_ordinal_to_doc_mapping_and_deleted():
    base = np.array([[0.0, 1.0 / i] for i in range(1, 11)], np.float32)
    o2d = np.array([19, 17, 15, 13, 11, 9, 7, 5, 3, -1], np.int32)  # sparse docIds, last ordinal deleted
    fx = make_fixture(O.SIM_EUCLIDEAN, base, base[:1], max_degree=32, ord_to_doc=o2d, max_doc=20)
    docs, _, cnt, _ = _search_ids(fx, [0.0, 0.0], 3)
    assert cnt == 3 and list(docs) == [3, 5, 7]  # ordinal 9 (closest) has no doc and is never returned


# ---- score formulas: JVectorEngineIT.java:417-438 via CommonTestUtils.java:84-93 ------------------------
@pytest.mark.parametrize("sim", [O.SIM_EUCLIDEAN, O.SIM_DOT, O.SIM_COSINE])
def test_score_formulas(sim):
    test_vectors = np.array([[1.0, 1.0, 1.0], [2.0, 2.0, 2.0], [3.0, 3.0, 3.0]], np.float32) * 0.1
    queries = np.array([[1.0, 1.0, 1.0], [2.0, 2.0, 2.0], [3.0, 3.0, 3.0]], np.float32)
    for q in queries:
        for x in test_vectors:
            assert abs(O.exact_score(sim, q, x) - lucene_score(sim, q, x)) < 1e-4


def test_canonical_reduction_matches_float64():
    rng = np.random.default_rng(3)
    for dim in (2, 16, 96, 128, 768, 1536, 1000):
        a = rng.standard_normal(dim).astype(np.float32)
        b = rng.standard_normal(dim).astype(np.float32)
        for sim in (O.SIM_EUCLIDEAN, O.SIM_DOT, O.SIM_COSINE):
            ref = lucene_score(sim, a, b)
            assert abs(O.exact_score(sim, a, b) - ref) <= 1e-5 * max(1.0, abs(ref))


# ---- brute force: ties -> lower docId --------------------------------------------------------------------
def test_exact_topk_tie_break():
    base = np.array([[1.0, 0.0], [0.0, 1.0], [1.0, 0.0], [0.5, 0.5], [1.0, 0.0]], np.float32)
    fx = make_fixture(O.SIM_DOT, base, base[:1], max_degree=4)
    docs, scores, counts = fx.oracle_index().exact_topk(np.array([[1.0, 0.0]], np.float32), 4)
    assert list(docs[0]) == [0, 2, 4, 3] and counts[0] == 4
    assert scores[0][0] == scores[0][1] == scores[0][2] == 1.0


def test_rerank_k_must_cover_k():
    base = np.array([[0.0, 1.0 / i] for i in range(1, 11)], np.float32)
    fx = make_fixture(O.SIM_EUCLIDEAN, base, base[:1], max_degree=8)
    with pytest.raises(ValueError):
        fx.oracle_index().search(base[:1], 5, 3)


# ---- PQ recall floor: KNNJVectorTests.java:1358-1403 (1024 x 16, L2, k=50, overquery 5 -> 1.0 +- 0.05) ---
def test_pq_recall_floor_reference_seed():
    dim, n, k = 16, 1024, 50
    base = O.java_random_vectors(n, dim, 1)          # TestUtils.generateRandomVectors -> seed 1
    target = np.zeros((1, dim), np.float32)          # generateZerosVectorWithLastValue(dim, 0)
    m = O.default_num_subspaces(dim)                 # 16 -> 16 subspaces
    fx = make_fixture(O.SIM_EUCLIDEAN, base, target, max_degree=32, pq_m=m)
    ix = fx.oracle_index()
    docs, _, counts, stats = ix.search(target, k, k * 5)
    gt, _, _ = ix.exact_topk(target, k)
    assert counts[0] == k
    assert recall(docs, gt) >= 0.95
    assert stats[0][3] > 0  # reranked


# ---- recall with overquery 1 <= overquery 5: KNNJVectorTests.java:1409-1464 ------------------------------
def test_overquery_monotone():
    dim, n, k = 16, 2048, 10
    base = O.java_random_vectors(n, dim, 1)
    queries = O.java_random_vectors(20, dim, n + 1)
    fx = make_fixture(O.SIM_EUCLIDEAN, base, queries, max_degree=32, pq_m=8)
    ix = fx.oracle_index()
    gt, _, _ = ix.exact_topk(queries, k)
    r1 = recall(ix.search(queries, k, k * 1)[0], gt)
    r5 = recall(ix.search(queries, k, k * 5)[0], gt)
    assert r5 >= r1 and r5 >= 0.9


# ---- merge scenarios: JVectorWriterMergeTests.java:244-262 (dim 128, seeds 42/43, k=10, overquery 5) -----
@pytest.mark.parametrize("n", [100, 300, 601])
def test_merge_scenario_recall_no_pq(n):
    base = O.java_random_vectors(n, 128, 42)
    queries = O.java_random_vectors(10, 128, 43)
    fx = make_fixture(O.SIM_EUCLIDEAN, base, queries, max_degree=32)
    ix = fx.oracle_index()
    docs, _, counts, _ = ix.search(queries, 10, 50)
    gt, _, _ = ix.exact_topk(queries, 10)
    assert (counts == 10).all()
    assert recall(docs, gt) >= 0.98


# ---- deletions: simpleDeletionTest, JVectorWriterMergeTests.java:293-301 (100 docs, 20..69 deleted) -------
def test_deleted_docs_never_returned():
    base = O.java_random_vectors(100, 128, 42)
    queries = O.java_random_vectors(10, 128, 43)
    live = np.ones(100, bool)
    live[20:70] = False
    fx = make_fixture(O.SIM_EUCLIDEAN, base, queries, max_degree=32)
    ix = fx.oracle_index()
    bits = O.make_accept_bits(live)
    docs, _, counts, _ = ix.search(queries, 10, 50, accept_bits=bits)
    gt, _, _ = ix.exact_topk(queries, 10, accept_bits=bits)
    assert (counts == 10).all()
    assert not np.isin(docs, np.arange(20, 70)).any()
    assert recall(docs, gt) == 1.0


# ---- JVectorIndexQuantization.java:428-446 ----------------------------------------------------------------
def test_default_num_subspaces_table():
    expect = {16: 16, 32: 32, 48: 32, 64: 32, 96: 48, 128: 64, 200: 100, 256: 100, 400: 100, 768: 192, 1024: 192,
              1536: 192, 2048: 256, 4096: 512}
    for d, m in expect.items():
        assert O.default_num_subspaces(d) == m


def test_pq_subspace_split():
    sizes, offs = O.pq_subspaces(10, 4)  # base 2, first 10%4=2 get 3
    assert list(sizes) == [3, 3, 2, 2] and list(offs) == [0, 3, 6, 8]
    sizes, offs = O.pq_subspaces(768, 192)
    assert (sizes == 4).all() and offs[-1] == 764


def test_pq_encode_first_min_wins():
    # two identical centroids: strict '<' keeps the lower index
    cb = np.array([[0.0, 0.0], [1.0, 1.0], [1.0, 1.0], [5.0, 5.0]], np.float32).reshape(-1)
    x = np.array([[1.0, 1.0], [0.4, 0.4], [9.0, 9.0]], np.float32)
    codes = O.pq_encode(x, 1, 4, cb)
    assert list(codes[:, 0]) == [1, 0, 3]


def test_pq_lut_and_adc_consistent():
    rng = np.random.default_rng(5)
    base = rng.standard_normal((600, 24)).astype(np.float32)
    q = rng.standard_normal((3, 24)).astype(np.float32)
    for sim in (O.SIM_EUCLIDEAN, O.SIM_DOT, O.SIM_COSINE):
        fx = make_fixture(sim, base, q, max_degree=8, pq_m=6, pq_k=32)
        ix = fx.oracle_index()
        nodes = np.tile(np.arange(50, dtype=np.int32), (3, 1))
        adc = ix.adc_scores(q, nodes)
        adc1 = fx.oracle_index(adc_order=32).adc_scores(q, nodes)
        np.testing.assert_allclose(adc, adc1, rtol=1e-5)
        # ADC score == exact score against the decoded vector
        sizes, offs = O.pq_subspaces(24, 6)
        cbs = fx.codebooks.reshape(6, 32, 4)
        dec = np.concatenate([cbs[m][fx.codes[:50, m]] for m in range(6)], axis=1)
        if fx.gcent is not None:
            dec = dec + fx.gcent
        for i in range(3):
            for j in range(50):
                assert abs(adc[i, j] - lucene_score(sim, q[i], dec[j])) < 2e-5


def test_merge_topk_ties_prefer_lower_doc():
    docs = np.array([[[5, 9, -1]], [[2, 7, 8]]], np.int32)
    scores = np.array([[[0.9, 0.5, 0.0]], [[0.9, 0.5, 0.1]]], np.float32)
    d, s, c = O.merge_topk(docs, scores, 3)
    assert list(d[0]) == [2, 5, 7] and c[0] == 3


def test_quantised_adc_table_bounds_and_recall():
    """adc_order = -8 (the 8-bit table the GPU production traversal uses): every entry lies inside its ball bound, the
    integer sum reproduces the fp32 ADC sum within M * delta / 2, and recall@10 stays within the north-star gate."""
    base, q = clustered(3000, 64, 60, seed=17, normalize=True)
    for sim, m in ((O.SIM_DOT, 16), (O.SIM_EUCLIDEAN, 32), (O.SIM_COSINE, 8)):
        fx = make_fixture(sim, base, q, max_degree=16, pq_m=m)
        ix8, ix = fx.oracle_index(adc_order=-8), fx.oracle_index()
        q8, prm = ix8.lut_q8(q)
        lut = O.pq_lut(sim, 64, m, 256, fx.codebooks, fx.gcent, q)
        delta, lo_sum = prm[:, 0], prm[:, 1]
        assert (delta > 0).all() and q8.max() <= 255
        codes = fx.codes[:200]
        isum = q8[:, np.arange(m)[None, :], codes].astype(np.int64).sum(axis=2)          # [nq, 200]
        fsum = lut[:, np.arange(m)[None, :], codes].astype(np.float64).sum(axis=2)
        approx = delta[:, None].astype(np.float64) * isum + lo_sum[:, None]
        assert (np.abs(approx - fsum) <= m * delta[:, None] * 0.5 + 1e-4).all()
        gt = ix.exact_topk(q, 10)[0]
        r8 = ix8.search(q, 10, 50)[0]
        r32 = ix.search(q, 10, 50)[0]
        assert recall(r8, gt) >= recall(r32, gt) - 0.01


# ---- NVQ-inline vectors (nvq+pq segments): decoder pinned to the in-tree Java code ------------------------------------
def _java_logistic_nqt(value, alpha, x0):
    """JVectorIndexQuantization.java:344-351 in numpy float32 (Math.fma evaluated in float64: exact for float operands)."""
    f = np.float32
    temp = f(np.float64(value) * np.float64(alpha) + np.float64(f(-(f(alpha) * f(x0)))))
    p = int(np.floor(np.float64(f(temp + f(0.5))) + 0.5))                      # Math.round
    m = f(np.float64(f(temp - f(p))) * 0.5 + 1.0).view(np.int32)
    t = np.int32(int(m) + (p << 23)).view(np.float32)
    return f(t / f(t + f(1)))


def _java_logit_nqt(scaled, inv_alpha, x0):
    """JVectorIndexQuantization.java:354-361."""
    f = np.float32
    z = f(scaled / f(f(1) - scaled))
    temp = int(z.view(np.int32))
    e = temp & 0x7F800000
    p = f((e >> 23) - 128)
    m = np.int32((temp & 0x007FFFFF) + 0x3F800000).view(np.float32)
    return f(f(f(m + p) * inv_alpha) + x0)


def _java_nvq_dequantize(bytes_row, params_row, gmean, sizes, offs):
    """nvqDequantize, JVectorIndexQuantization.java:316-341."""
    f = np.float32
    out = np.zeros(len(bytes_row), np.float32)
    for m, (size, off) in enumerate(zip(sizes, offs)):
        growth, mid, lo, hi = (f(x) for x in params_row[m])
        delta = f(hi - lo)
        sgr = f(growth / delta)
        smid = f(mid * delta)
        bias = _java_logistic_nqt(lo, sgr, smid)
        scale = f(f(_java_logistic_nqt(hi, sgr, smid) - bias) / f(255.0))
        inv = f(f(1.0) / sgr)
        for d in range(size):
            b = f(int(bytes_row[off + d]))
            sv = f(np.float64(b) * np.float64(scale) + np.float64(bias))       # Math.fma
            out[off + d] = _java_logit_nqt(sv, inv, smid)
    return (out + gmean).astype(np.float32)


def test_nvq_decoder_matches_the_in_tree_java_formulas():
    x = O.java_random_vectors(40, 30, 73)                                      # dim 30, 4 sub-vectors: sizes 8, 8, 7, 7
    b, prm, g = O.nvq_encode(x, 4, growth_rate=2.5, midpoint=0.1)
    deq = O.nvq_dequantize(b, prm, g)
    sizes, offs = O.pq_subspaces(30, 4)
    for i in range(40):
        want = _java_nvq_dequantize(b[i], prm[i], g, sizes, offs)
        np.testing.assert_array_equal(deq[i].view(np.uint32), want.view(np.uint32))
    assert np.abs(deq - x).max() < 1.0 / 128                                   # 8 bits over a unit range
    # the fast logistic is the base-2 sigmoid up to its piecewise-linear mantissa approximation
    for v, a, x0 in ((0.3, 2.0, 0.1), (-1.5, 0.7, 0.0), (4.0, 1.0, 2.0)):
        t = v * a - a * x0
        assert abs(float(_java_logistic_nqt(np.float32(v), np.float32(a), np.float32(x0))) - 2.0 ** t / (1 + 2.0 ** t)) < 0.02


def test_nvq_rerank_recall_floor_at_the_reference_seed():
    """JVectorNVQTests (dimension 128, 2 sub-vectors, seed 73, overquery 10): recall >= 0.85 with NVQ-inline vectors;
    here the traversal uses the auxiliary PQ codes and the reranker the dequantised vectors (nvq+pq)."""
    base = O.java_random_vectors(2000, 128, 73)
    q = O.java_random_vectors(30, 128, 74)
    fx = make_fixture(O.SIM_EUCLIDEAN, base, q, max_degree=16, pq_m=O.default_num_subspaces(128))
    b, prm, g = O.nvq_encode(base, 2)
    ix = O.OracleIndex(O.SIM_EUCLIDEAN, base, fx.adjacency, fx.entry, pq_m=fx.pq_m, pq_k=fx.pq_k, pq_codebooks=fx.codebooks,
                       pq_global_centroid=fx.gcent, pq_codes=fx.codes, nvq_m=2, nvq_bytes=b, nvq_params=prm, nvq_global_mean=g)
    gt = fx.oracle_index().exact_topk(q, 10)[0]
    docs, scores, _, st = ix.search(q, 10, 100)
    assert recall(docs, gt) >= 0.85
    assert (st[:, 3] == 100).all()
    # scores are the exact similarity of the DEQUANTISED vectors
    deq = O.nvq_dequantize(b, prm, g)
    for i in range(5):
        for d, s in zip(docs[i], scores[i]):
            assert s == np.float32(O.exact_score(O.SIM_EUCLIDEAN, q[i], deq[d]))


# ---- quantised flush builds its graph with PQ build scores (JVectorWriter.java:238-244): same recall floor as above through
#      the PQ-scored builder (testJVectorKnnIndex_simpleCase_withQuantization, KNNJVectorTests.java:1358-1403) ----------------
def test_pq_scored_build_recall_floor_reference_seed():
    dim, n, k = 16, 1024, 50
    base = O.java_random_vectors(n, dim, 1)
    target = np.zeros((1, dim), np.float32)
    m = O.default_num_subspaces(dim)
    cb, g = O.pq_train(base, m, 256, center=True, iters=6, seed=7)
    codes = O.pq_encode(base, m, 256, cb, g)
    dec = O.pq_decode(codes, dim, 256, cb, g)
    assert np.abs(dec - base).max() < np.abs(base).max()            # reconstructions, not the vectors
    adj, entry = O.graph_build_pq(codes, dim, 256, cb, g, O.SIM_EUCLIDEAN, 32, 100)
    ix = O.OracleIndex(O.SIM_EUCLIDEAN, base, adj, entry, pq_m=m, pq_k=256, pq_codebooks=cb, pq_global_centroid=g, pq_codes=codes)
    docs, _, counts, _ = ix.search(target, k, k * 5)
    gt, _, _ = ix.exact_topk(target, k)
    assert counts[0] == k and recall(docs, gt) >= 0.95


# ---- testJVectorKnnIndex_simpleCase_withQuantization_rerank (KNNJVectorTests.java:1409-1464): vectors (0, .., 0, i), i = 1..1024,
#      k = 1; "recall" = the top document's score reaches the score of (0, .., 0, k); over-query 1 must not beat over-query 5 ----
def test_quantization_rerank_reference_fixture():
    dim, n, k = 16, 1024, 1
    base = np.zeros((n, dim), np.float32)
    base[:, -1] = np.arange(1, n + 1, dtype=np.float32)
    target = np.zeros((1, dim), np.float32)
    m = O.default_num_subspaces(dim)
    cb, g = O.pq_train(base, m, 256, center=True, iters=6, seed=7)
    codes = O.pq_encode(base, m, 256, cb, g)
    adj, entry = O.graph_build_pq(codes, dim, 256, cb, g, O.SIM_EUCLIDEAN, 32, 100)
    ix = O.OracleIndex(O.SIM_EUCLIDEAN, base, adj, entry, pq_m=m, pq_k=256, pq_codebooks=cb, pq_global_centroid=g, pq_codes=codes)
    expected_min = 1.0 / (1.0 + float(k) ** 2)                      # EUCLIDEAN.compare(target, (0, .., 0, k))

    def score_recall(over):
        _, scores, counts, _ = ix.search(target, k, k * over)
        assert counts[0] == k
        return float(np.mean(scores[0, :k] >= expected_min - 1e-6))

    low, high = score_recall(1), score_recall(5)
    assert low <= high and high == 1.0


# ---- testJVectorKnnIndex_happyCase_withQuantization_multipleSegments (KNNJVectorTests.java:1471-1530): two flushes of exactly the
#      minimum batch, force-merged: leading-segment merge (exact scores, :1290) + mergePQ with the leading codebooks (:1072-1124) --
def test_quantised_two_segment_merge_recall_floor():
    dim, per, k = 16, 1024, 50
    vectors = O.java_random_vectors(2 * per, dim, 1)
    target = np.zeros((1, dim), np.float32)
    m = O.default_num_subspaces(dim)
    lead = vectors[:per]
    cb, g = O.pq_train(lead, m, 256, center=True, iters=6, seed=7)
    lead_codes = O.pq_encode(lead, m, 256, cb, g)
    lead_adj, lead_entry = O.graph_build_pq(lead_codes, dim, 256, cb, g, O.SIM_EUCLIDEAN, 32, 100)     # flush of segment 0
    adj = O.graph_extend(vectors, lead_adj, lead_entry, O.SIM_EUCLIDEAN, 100)                          # merge: insert segment 1
    codes = O.pq_encode(vectors, m, 256, cb, g)                                                        # mergePQ: leading codebooks
    ix = O.OracleIndex(O.SIM_EUCLIDEAN, vectors, adj, lead_entry, pq_m=m, pq_k=256, pq_codebooks=cb, pq_global_centroid=g, pq_codes=codes)
    docs, _, counts, _ = ix.search(target, k, k * 5)
    gt, _, _ = ix.exact_topk(target, k)
    assert counts[0] == k and recall(docs, gt) >= 0.95
