"""NVQ-inline segments ("nvq+pq"): the traversal uses the auxiliary PQ codes, the reranker scores the
dequantised 8-bit vectors (JVectorReader.java:352-358, decoder JVectorIndexQuantization.java:316-361).  The device decoder
repeats the oracle's operations one by one, so ids and scores must be bit-identical."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import clustered, make_fixture, recall

pytestmark = pytest.mark.gpu


def _nvq_pair(jv, fx, nvq_m, flags=0, keep_vectors=False, adc_order=32):
    b, prm, g = O.nvq_encode(fx.base, nvq_m)
    ora = O.OracleIndex(fx.sim, fx.base, fx.adjacency, fx.entry, fx.ord_to_doc, fx.max_doc, fx.pq_m, fx.pq_k, fx.codebooks, fx.gcent,
                        fx.codes, adc_order=adc_order, nvq_m=nvq_m, nvq_bytes=b, nvq_params=prm, nvq_global_mean=g)
    gi = jv.GpuIndex(fx.sim, fx.base if keep_vectors else None, fx.adjacency, fx.entry, ord_to_doc=fx.ord_to_doc, max_doc=fx.max_doc,
                     pq_m=fx.pq_m, pq_k=fx.pq_k, pq_codebooks=fx.codebooks, pq_global_centroid=fx.gcent, pq_codes=fx.codes, flags=flags,
                     nvq_m=nvq_m, nvq_bytes=b, nvq_params=prm, nvq_global_mean=g)
    return ora, gi, (b, prm, g)


@pytest.mark.parametrize("sim,dim,pq_m,nvq_m", [(O.SIM_EUCLIDEAN, 128, 64, 2), (O.SIM_DOT, 64, 16, 4), (O.SIM_COSINE, 30, 6, 4)])
def test_nvq_rerank_bit_exact(jv, sim, dim, pq_m, nvq_m):
    base, q = clustered(3000, dim, 48, seed=70 + dim, normalize=(sim == O.SIM_DOT))
    fx = make_fixture(sim, base, q, max_degree=16, pq_m=pq_m)
    ora, gi, _ = _nvq_pair(jv, fx, nvq_m)
    with gi:
        for k, rk in ((10, 100), (5, 5)):
            r = gi.search(q, k, rk, expand_width=-1)                    # reference-order traversal + NVQ rerank
            wd, ws, wc, wst = ora.search(q, k, rk)
            np.testing.assert_array_equal(r.docs, wd)
            np.testing.assert_array_equal(r.scores.view(np.uint32), ws.view(np.uint32))
            np.testing.assert_array_equal(r.counts, wc)
            np.testing.assert_array_equal(r.stats, wst)
        with pytest.raises(NotImplementedError):                        # JVectorQuantizedNvqVectorValues.java:33-36
            gi.exact_topk(q, 10)


def test_nvq_with_production_traversal_and_reference_recall_floor(jv):
    """JVectorNVQTests: dimension 128, 2 sub-vectors, seed 73, overquery 10, recall >= 0.85."""
    base = O.java_random_vectors(3000, 128, 73)
    q = O.java_random_vectors(64, 128, 74)
    fx = make_fixture(O.SIM_EUCLIDEAN, base, q, max_degree=16, pq_m=O.default_num_subspaces(128))
    ora, gi, (b, prm, g) = _nvq_pair(jv, fx, 2, flags=jv.native.FLAG_LUT_U8, keep_vectors=True, adc_order=0)
    with gi:
        gt, _, _ = gi.exact_topk(q, 10)                                 # fp32 vectors were kept: ground truth on the device
        r = gi.search(q, 10, 100)
        wd = ora.search(q, 10, 100)[0]
        assert recall(r.docs, gt) >= 0.85
        assert recall(r.docs, gt) >= recall(wd, gt) - 0.015
        deq = O.nvq_dequantize(b, prm, g)                               # NVQ wins over the fp32 vectors when both are given
        for i in range(8):
            for d, s in zip(r.docs[i], r.scores[i]):
                assert s == np.float32(O.exact_score(O.SIM_EUCLIDEAN, q[i], deq[d]))


@pytest.mark.parametrize("sim", [O.SIM_EUCLIDEAN, O.SIM_DOT, O.SIM_COSINE, O.SIM_MIP])
def test_nvq_only_segment_is_traversed_with_the_nvq_reranker(jv, sim):
    """No auxiliary PQ blob (JVectorReader.java:357-358): DefaultSearchScoreProvider(view.rerankerFor(q, sim)) — the traversal is
    scored exactly against the DEQUANTISED inline vectors and the score is NOT MIP-wrapped.  Same as an un-quantised segment whose
    vectors are the dequantised ones (for MIP: with the plain dot-product score)."""
    base, q = clustered(1500, 24, 40, seed=17, normalize=sim in (O.SIM_DOT, O.SIM_MIP))
    b, prm, g = O.nvq_encode(base, 2)
    deq = O.nvq_dequantize(b, prm, g)
    fx = make_fixture(sim, base, q, max_degree=12)
    inner = O.SIM_DOT if sim == O.SIM_MIP else sim                        # un-wrapped: MIP scores like DOT here
    want = O.OracleIndex(inner, deq, fx.adjacency, fx.entry)
    wd, ws, wc, wst = want.search(q, 10, 50)
    with jv.GpuIndex(sim, None, fx.adjacency, fx.entry, nvq_m=2, nvq_bytes=b, nvq_params=prm, nvq_global_mean=g) as gi:
        r = gi.search(q, 10, 50, expand_width=-1)                         # strict reference-order kernel
        np.testing.assert_array_equal(r.docs, wd)
        np.testing.assert_array_equal(r.scores.view(np.uint32), ws.view(np.uint32))
        np.testing.assert_array_equal(r.stats[:, :2], wst[:, :2])
        rf = gi.search(q, 10, 50)                                         # production (wide-step) kernel
        gt, _, _ = want.exact_topk(q, 10)
        assert recall(rf.docs, gt) >= recall(wd, gt) - 0.02
        with pytest.raises(NotImplementedError):                          # brute force: JVectorQuantizedNvqVectorValues.java:33-36
            gi.exact_topk(q, 10)
