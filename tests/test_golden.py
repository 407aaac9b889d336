"""The committed golden fixture (tests/golden/golden_small.npz, made by tests/golden/make_golden.py) pins the oracle: every
stage of the hot path must reproduce the stored outputs bit for bit from the seeds alone."""
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O

GOLD = Path(__file__).resolve().parent / "golden" / "golden_small.npz"
SIMS = (("l2", O.SIM_EUCLIDEAN), ("dot", O.SIM_DOT), ("cos", O.SIM_COSINE))


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


def _inputs(g):
    return O.java_random_vectors(int(g["n"]), int(g["dim"]), int(g["seed_base"])), \
        O.java_random_vectors(int(g["nq"]), int(g["dim"]), int(g["seed_query"]))


def test_fixture_is_small_and_complete(gold):
    assert GOLD.stat().st_size < 1 << 20
    for name, _ in SIMS:
        for key in ("adj", "cb", "codes", "docs", "scores", "exact_docs", "lut", "q8", "docs8"):
            assert f"{name}_{key}" in gold


@pytest.mark.parametrize("name,sim", SIMS)
def test_oracle_reproduces_the_golden_outputs(gold, name, sim):
    base, q = _inputs(gold)
    m, dim, r = int(gold["m"]), int(gold["dim"]), int(gold["r"])
    adj, entry = O.graph_build(base, sim, r, 100)
    np.testing.assert_array_equal(adj, gold[f"{name}_adj"])
    assert entry == int(gold[f"{name}_entry"])
    cb, g = O.pq_train(base, m, 256, center=(sim == O.SIM_EUCLIDEAN), iters=6, seed=7)
    np.testing.assert_array_equal(cb.view(np.uint32), gold[f"{name}_cb"].view(np.uint32))
    codes = O.pq_encode(base, m, 256, cb, g)
    np.testing.assert_array_equal(codes, gold[f"{name}_codes"])
    ix = O.OracleIndex(sim, base, adj, entry, pq_m=m, pq_k=256, pq_codebooks=cb, pq_global_centroid=g, pq_codes=codes, adc_order=32)
    docs, scores, counts, stats = ix.search(q, 10, 50)
    np.testing.assert_array_equal(docs, gold[f"{name}_docs"])
    np.testing.assert_array_equal(scores.view(np.uint32), gold[f"{name}_scores"].view(np.uint32))
    np.testing.assert_array_equal(stats, gold[f"{name}_stats"])
    ed, es, _ = ix.exact_topk(q, 10)
    np.testing.assert_array_equal(ed, gold[f"{name}_exact_docs"])
    np.testing.assert_array_equal(es.view(np.uint32), gold[f"{name}_exact_scores"].view(np.uint32))
    np.testing.assert_array_equal(O.pq_lut(sim, dim, m, 256, cb, g, q[:2]).view(np.uint32), gold[f"{name}_lut"].view(np.uint32))
    ix8 = O.OracleIndex(sim, base, adj, entry, pq_m=m, pq_k=256, pq_codebooks=cb, pq_global_centroid=g, pq_codes=codes, adc_order=-8)
    q8, q8p = ix8.lut_q8(q[:2])
    np.testing.assert_array_equal(q8, gold[f"{name}_q8"])
    np.testing.assert_array_equal(q8p.view(np.uint32), gold[f"{name}_q8p"].view(np.uint32))
    np.testing.assert_array_equal(ix8.search(q, 10, 50)[0], gold[f"{name}_docs8"])


def test_nvq_decoder_reproduces_the_golden_vectors(gold):
    deq = O.nvq_dequantize(gold["nvq_bytes"], gold["nvq_params"], gold["nvq_gmean"])
    np.testing.assert_array_equal(deq.view(np.uint32), gold["nvq_deq"].view(np.uint32))


@pytest.mark.parametrize("name,sim", (("l2", O.SIM_EUCLIDEAN), ("cos", O.SIM_COSINE)))
def test_oracle_reproduces_the_golden_merge(gold, name, sim):
    """Leading-segment merge: seeded build over the first 1000 ordinals + the rest, then delete consolidation."""
    base, _ = _inputs(gold)
    r = int(gold["r"])
    seed_adj, seed_entry = O.graph_build(base[:1000], sim, r, 100)
    assert seed_entry == int(gold[f"{name}_merge_seed_entry"])
    ext = O.graph_extend(base, seed_adj, seed_entry, sim)
    np.testing.assert_array_equal(ext, gold[f"{name}_merge_ext"])
    cons, e = O.graph_remove_deleted(base, ext, seed_entry, gold[f"{name}_merge_dead"], sim)
    np.testing.assert_array_equal(cons, gold[f"{name}_merge_cons"])
    assert e == int(gold[f"{name}_merge_cons_entry"])
