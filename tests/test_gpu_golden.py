"""Device path against the COMMITTED golden fixture (tests/golden/golden_small.npz): inputs regenerated from the seeds with a
numpy restatement of java.util.Random, everything else read from the file — no oracle call on the GPU box for the targets."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = Path(__file__).resolve().parent / "golden" / "golden_small.npz"
SIMS = (("l2", 0), ("dot", 1), ("cos", 2))


def java_random_vectors(n, dim, seed):
    """new java.util.Random(seed).nextFloat() x n*dim (48-bit LCG, 24 high bits), TestUtils.java:108-124."""
    mult, add, mask = 0x5DEECE66D, 0xB, (1 << 48) - 1
    s = (seed ^ mult) & mask
    out = np.empty(n * dim, np.float32)
    for i in range(n * dim):
        s = (s * mult + add) & mask
        out[i] = np.float32((s >> 24) / float(1 << 24))
    return out.reshape(n, dim)


@pytest.fixture(scope="module")
def gold():
    g = dict(np.load(GOLD))
    g["base"] = java_random_vectors(int(g["n"]), int(g["dim"]), int(g["seed_base"]))
    g["q"] = java_random_vectors(int(g["nq"]), int(g["dim"]), int(g["seed_query"]))
    return g


@pytest.mark.parametrize("name,sim", SIMS)
def test_device_reproduces_the_golden_outputs(jv, gold, name, sim):
    base, q, m, r = gold["base"], gold["q"], int(gold["m"]), int(gold["r"])
    gc = gold.get(f"{name}_gcent")
    # K6 / f-2 / f-3: codebooks, codes and graph built on the device equal the stored ones bit for bit
    cb, g = jv.pq_train(base, m, 256, sim == 0, 6, 7)
    np.testing.assert_array_equal(cb.view(np.uint32), gold[f"{name}_cb"].view(np.uint32))
    np.testing.assert_array_equal(jv.pq_encode(base, m, 256, gold[f"{name}_cb"], gc), gold[f"{name}_codes"])
    adj, entry = jv.graph_build(base, sim, r, 100, 1.2, 1.2)
    np.testing.assert_array_equal(adj, gold[f"{name}_adj"])
    assert entry == int(gold[f"{name}_entry"])
    with jv.GpuIndex(sim, base, gold[f"{name}_adj"], int(gold[f"{name}_entry"]), pq_m=m, pq_k=256, pq_codebooks=gold[f"{name}_cb"],
                     pq_global_centroid=gc, pq_codes=gold[f"{name}_codes"], flags=jv.native.FLAG_LUT_U8) as gi:
        r_strict = gi.search(q, 10, 50, expand_width=-1)                        # K1+K2 (reference order) + K3
        np.testing.assert_array_equal(r_strict.docs, gold[f"{name}_docs"])
        np.testing.assert_array_equal(r_strict.scores.view(np.uint32), gold[f"{name}_scores"].view(np.uint32))
        np.testing.assert_array_equal(r_strict.stats, gold[f"{name}_stats"])
        ed, es, _ = gi.exact_topk(q, 10)                                        # K5
        np.testing.assert_array_equal(ed, gold[f"{name}_exact_docs"])
        np.testing.assert_array_equal(es.view(np.uint32), gold[f"{name}_exact_scores"].view(np.uint32))
        np.testing.assert_array_equal(gi.pq_lut(q[:2]).view(np.uint32), gold[f"{name}_lut"].view(np.uint32))   # K1 fp32
        q8, q8p = gi.pq_lut_q8(q[:2])                                           # K1 8-bit
        np.testing.assert_array_equal(q8, gold[f"{name}_q8"])
        np.testing.assert_array_equal(q8p.view(np.uint32), gold[f"{name}_q8p"].view(np.uint32))
        r1 = gi.search(q, 10, 50, expand_width=1)                               # production kernel, best-first order
        assert np.mean([np.array_equal(a, b) for a, b in zip(r1.docs, gold[f"{name}_docs8"])]) >= 0.9


def test_device_nvq_rerank_scores_the_golden_decoded_vectors(jv, gold):
    """NVQ-inline rerank on the first 64 vectors: returned scores = exact similarity of the stored dequantised vectors."""
    n = gold["nvq_bytes"].shape[0]
    base, q, m = gold["base"][:n], gold["q"], int(gold["m"])
    adj = np.full((n, 8), -1, np.int32)
    for i in range(n):                                                          # a ring: every node reachable
        adj[i, 0], adj[i, 1] = (i + 1) % n, (i + 7) % n
    cb = gold["dot_cb"]
    codes = jv.pq_encode(base, m, 256, cb)
    with jv.GpuIndex(1, None, adj, 0, pq_m=m, pq_k=256, pq_codebooks=cb, pq_codes=codes, nvq_m=2, nvq_bytes=gold["nvq_bytes"],
                     nvq_params=gold["nvq_params"], nvq_global_mean=gold["nvq_gmean"]) as gi:
        r = gi.search(q, 5, n, expand_width=-1)                                 # rerankK = n: everything is reranked
        deq = gold["nvq_deq"]
        for i in range(len(q)):
            want = (1.0 + deq @ q[i]) / 2.0                                      # DOT score of the decoded vectors
            top = np.argsort(-want, kind="stable")[:5]
            assert set(r.docs[i].tolist()) == set(top.tolist())
            np.testing.assert_allclose(r.scores[i], np.sort(want)[::-1][:5], rtol=1e-5)


@pytest.mark.parametrize("name,sim", (("l2", 0), ("cos", 2)))
def test_device_reproduces_the_golden_merge(jv, gold, name, sim):
    """Merge path against the committed file (no oracle call): leading graph over the first 1000 ordinals extended by the rest
    (jv_graph_extend), then delete consolidation (jv_graph_remove_deleted)."""
    base, r = gold["base"], int(gold["r"])
    seed_adj, seed_entry = jv.graph_build(base[:1000], sim, r, 100, 1.2, 1.2)
    assert seed_entry == int(gold[f"{name}_merge_seed_entry"])
    ext = jv.graph_extend(base, seed_adj, seed_entry, sim, 100)
    np.testing.assert_array_equal(ext, gold[f"{name}_merge_ext"])
    cons, e = jv.graph_remove_deleted(base, gold[f"{name}_merge_ext"], seed_entry, gold[f"{name}_merge_dead"], sim)
    np.testing.assert_array_equal(cons, gold[f"{name}_merge_cons"])
    assert e == int(gold[f"{name}_merge_cons_entry"])
