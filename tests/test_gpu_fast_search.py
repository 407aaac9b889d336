"""The production (fast) traversal kernel: unified sorted list, expand_width E, direct-mapped visited filter,
optional fp16 ADC table.  Gate = north star: recall@10 within 0.005 of the reference path at equal rerankK;
for E = 1 and the fp32 table the traversal is the reference's best-first order, so ids/scores must be identical."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import clustered, make_fixture, recall

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fx_dot():
    base, q = clustered(8000, 64, 100, seed=41, normalize=True)
    return make_fixture(O.SIM_DOT, base, q, max_degree=32, pq_m=16)


@pytest.fixture(scope="module")
def fx_l2():
    base, q = clustered(6000, 96, 64, seed=42)
    return make_fixture(O.SIM_EUCLIDEAN, base, q, max_degree=16, pq_m=48)


@pytest.fixture(scope="module")
def fx_cos8():
    base, q = clustered(4000, 128, 64, seed=43)
    return make_fixture(O.SIM_COSINE, base, q, max_degree=16, pq_m=16)  # sub-dim 8


@pytest.fixture(scope="module")
def fx_exact():
    base = O.java_random_vectors(10000, 128, 42)
    q = O.java_random_vectors(200, 128, 43)
    return make_fixture(O.SIM_COSINE, base, q, max_degree=16)


def adc_lanes(m: int) -> int:
    """lanes per code row in the fast kernel (jv_search_fast.cu adc_lanes_for): <= 4 code words per lane."""
    nwords, lanes = (m + 3) // 4, 1
    while lanes < 32 and lanes * 4 < nwords:
        lanes *= 2
    return lanes


@pytest.mark.parametrize("name", ["fx_dot", "fx_l2", "fx_cos8", "fx_exact"])
def test_width1_is_reference_order(jv, request, name):
    fx = request.getfixturevalue(name)
    ora = fx.oracle_index(adc_order=adc_lanes(fx.pq_m) if fx.pq_m else 0)
    with fx.gpu_index(jv) as gi:
        for k, rk in ((10, 50), (1, 1), (20, 200)):
            r = gi.search(fx.queries, k, rk, expand_width=1)
            wd, ws, wc, wst = ora.search(fx.queries, k, rk)
            np.testing.assert_array_equal(r.counts, wc)
            np.testing.assert_array_equal(r.docs, wd)
            np.testing.assert_allclose(r.scores, ws, rtol=1e-5, atol=0)
            np.testing.assert_array_equal(r.stats[:, 1], wst[:, 1])   # expansions
            np.testing.assert_array_equal(r.stats[:, 3], wst[:, 3])   # reranked
            assert (r.stats[:, 0] >= wst[:, 0]).all()                 # re-scored nodes count as visits
            assert r.stats[:, 0].mean() <= 1.15 * wst[:, 0].mean() + 4  # the 2-way tagged filter rarely evicts


@pytest.mark.parametrize("name", ["fx_dot", "fx_l2", "fx_cos8", "fx_exact"])
@pytest.mark.parametrize("width", [0, 2, 4, 8])
def test_wide_expansion_recall_parity(jv, request, name, width):
    fx = request.getfixturevalue(name)
    ora = fx.oracle_index()
    with fx.gpu_index(jv) as gi:
        gt, _, _ = gi.exact_topk(fx.queries, 10)
        r = gi.search(fx.queries, 10, 50, expand_width=width)
        wd, ws, wc, wst = ora.search(fx.queries, 10, 50)
        rec_gpu, rec_ref = recall(r.docs, gt), recall(wd, gt)
        # north star: within 0.005 at benchmark scale (10k queries, checked in bench.py); a 64..200-query fixture has
        # a sampling error of ~0.01 on recall, so the unit test allows 0.015
        assert rec_gpu >= rec_ref - 0.015
        np.testing.assert_array_equal(r.counts, wc)
        same = np.mean([np.array_equal(a, b) for a, b in zip(r.docs, wd)])
        if rec_ref > 0.9:  # on hard fixtures (recall ~0.6) the per-query lists legitimately differ with the width
            assert same >= 0.9
        # final scores are exact-rerank scores: identical wherever the same doc is returned
        for i in range(len(fx.queries)):
            ref = {int(d): s for d, s in zip(wd[i], ws[i]) if d >= 0}
            for d, s in zip(r.docs[i], r.scores[i]):
                if int(d) in ref:
                    assert s == ref[int(d)]
        # speculative width costs extra expansions (the last step always expands up to `width` entries)
        assert r.stats[:, 1].mean() <= 2.0 * wst[:, 1].mean() + 2 * max(width, 4), (r.stats[:, 1].mean(), wst[:, 1].mean())


@pytest.mark.parametrize("name", ["fx_dot", "fx_l2", "fx_cos8"])
def test_fp16_table_recall_parity(jv, request, name):
    fx = request.getfixturevalue(name)
    ora = fx.oracle_index()
    with fx.gpu_index(jv, flags=jv.native.FLAG_LUT_F16) as gi:
        gt, _, _ = gi.exact_topk(fx.queries, 10)
        wd = ora.search(fx.queries, 10, 50)[0]
        for width in (1, 4):
            r = gi.search(fx.queries, 10, 50, expand_width=width)
            assert recall(r.docs, gt) >= recall(wd, gt) - 0.015  # 64..100 queries: sampling error ~0.01 (see above)


def test_large_rerank_k_and_small_graph(jv, fx_l2):
    fx = fx_l2
    ora = fx.oracle_index()
    with fx.gpu_index(jv) as gi:
        gt, _, _ = gi.exact_topk(fx.queries, 100)
        r = gi.search(fx.queries, 100, 500)  # cfg 5 flavour: k = 100, rerankK = 500
        wd = ora.search(fx.queries, 100, 500)[0]
        assert recall(r.docs, gt) >= recall(wd, gt) - 0.005
        assert (r.counts == 100).all()
    # a graph smaller than rerankK: every node ends up in the list
    base, q = clustered(30, 16, 4, seed=3)
    small = make_fixture(O.SIM_EUCLIDEAN, base, q, max_degree=8)
    with small.gpu_index(jv) as gi:
        r = gi.search(q, 10, 50)
        wd, ws, wc, _ = small.oracle_index().search(q, 10, 50)
        np.testing.assert_array_equal(r.docs, wd)
        np.testing.assert_array_equal(r.counts, wc)


def test_filtered_and_threshold_queries_use_strict_kernel(jv, fx_l2):
    fx = fx_l2
    ora = fx.oracle_index(adc_order=32)
    rng = np.random.default_rng(1)
    bits = O.make_accept_bits(rng.random(fx.base.shape[0]) < 0.3)
    with fx.gpu_index(jv) as gi:
        r = gi.search(fx.queries, 10, 50, accept_bits=bits, expand_width=4)
        wd, ws, wc, _ = ora.search(fx.queries, 10, 50, accept_bits=bits)
        np.testing.assert_array_equal(r.docs, wd)
        np.testing.assert_array_equal(r.counts, wc)


def test_pinned_host_queries_are_read_in_place(jv, fx_dot):
    """jv_search_batch reads page-locked host queries zero-copy (no staged H2D); results must not change."""
    torch = pytest.importorskip("torch")
    fx = fx_dot
    pinned = torch.from_numpy(fx.queries.copy()).pin_memory()
    with fx.gpu_index(jv) as gi:
        a = gi.search(fx.queries, 10, 50)                 # pageable numpy -> staged copy
        b = gi.search(pinned.numpy(), 10, 50)             # pinned -> kernels read host memory directly
        np.testing.assert_array_equal(a.docs, b.docs)
        np.testing.assert_array_equal(a.scores, b.scores)
        # visited counts include re-scored nodes, which depend on the (timing-dependent) eviction order of the filter
        np.testing.assert_array_equal(a.stats[:, 1:], b.stats[:, 1:])
        c = gi.search(pinned.numpy(), 10, 50, expand_width=-1)
        d = gi.search(fx.queries, 10, 50, expand_width=-1)
        np.testing.assert_array_equal(c.docs, d.docs)
