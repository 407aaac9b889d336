"""Leading-segment merge, insert-only part (JVectorWriter.tryLeadingSegmentMerge, JVectorWriter.java:1166-1341): the leading
segment's graph is kept and the other segments' vectors are inserted (jv_graph_extend).  CPU: the oracle's seeded builder keeps
the recall floors of JVectorWriterMergeTests (0.98-1.0 on seeded data, JVectorWriterMergeTests.java:244-462).  GPU: the device
builder reproduces the oracle's merged graph bit for bit, and the codec-level merge answers queries like a fresh build."""
import numpy as np
import pytest

from tests.helpers import clustered, recall


def _recall_of(oracle, sim, base, adj, entry, queries, k=10, rk=50):
    ix = oracle.OracleIndex(sim, base, adj, entry)
    d, _, _, _ = ix.search(queries, k, rk)
    gd, _, _ = ix.exact_topk(queries, k)
    return recall(d, gd)


@pytest.mark.parametrize("sim_name", ["SIM_EUCLIDEAN", "SIM_DOT", "SIM_COSINE"])
def test_oracle_extend_keeps_merge_recall_floor(oracle, sim_name):
    sim = getattr(oracle, sim_name)
    base = oracle.java_random_vectors(601, 128, 42)             # the merge tests' seeded fixtures: dim 128, seeds 42 / 43,
    queries = oracle.java_random_vectors(10, 128, 43)           # up to 601 documents, k = 10, overquery 5
    n0 = 300
    adj0, e0 = oracle.graph_build(base[:n0], sim, 32, 100)
    adj = oracle.graph_extend(base, adj0, e0, sim)
    assert adj.shape == (601, 32) and adj.max() < 601
    assert ((adj >= 0).sum(1) > 0).all()                          # every inserted node is connected
    assert (adj[n0:] >= 0).any(1).all() and (adj[:n0] >= n0).any()  # and the old nodes link to the new ones
    full, ef = oracle.graph_build(base, sim, 32, 100)
    r_ext = _recall_of(oracle, sim, base, adj, e0, queries)
    r_full = _recall_of(oracle, sim, base, full, ef, queries)
    assert r_ext >= 0.98 and r_ext >= r_full - 0.02, (r_ext, r_full)
    with pytest.raises(ValueError):
        oracle.graph_extend(base, adj0, n0 + 5, sim)               # entry outside the seed


@pytest.mark.gpu
@pytest.mark.parametrize("sim_name", ["SIM_EUCLIDEAN", "SIM_DOT", "SIM_COSINE"])
def test_gpu_extend_matches_oracle_bit_for_bit(jv, oracle, sim_name):
    sim = getattr(oracle, sim_name)
    base, _ = clustered(5000, 48, 8, seed=21, normalize=sim_name != "SIM_EUCLIDEAN")
    n0 = 3000
    adj0, e0 = oracle.graph_build(base[:n0], sim, 16, 100)
    want = oracle.graph_extend(base, adj0, e0, sim)
    got = jv.graph_extend(base, adj0, e0, sim, 100)
    assert np.array_equal(got, want)
    # seed of one node, and a seed that is the whole input (nothing to insert)
    one = np.full((1, 16), -1, np.int32)
    assert np.array_equal(jv.graph_extend(base[:600], one, 0, sim), oracle.graph_extend(base[:600], one, 0, sim))
    assert np.array_equal(jv.graph_extend(base[:n0], adj0, e0, sim), adj0)
    bad = adj0.copy()
    bad[5, 0] = n0 + 7
    with pytest.raises(ValueError, match="outside"):
        jv.graph_extend(base, bad, e0, sim)


@pytest.mark.gpu
def test_gpu_codec_merge_uses_leading_graph(jv, oracle):
    V = jv.VectorSimilarityFunction
    base, queries = clustered(4000, 32, 64, seed=31)
    w = jv.JVectorWriter(max_conn=16, min_batch_size_for_quantization=1024)
    segs = []
    for lo, hi in ((0, 2500), (2500, 3400), (3400, 4000)):
        w.add_field("vec", V.EUCLIDEAN)
        for d, i in enumerate(range(lo, hi)):
            w.add_value("vec", d, base[i])
        segs.append(w.flush(hi - lo))
    merged = w.merge(segs)
    fd = merged.fields["vec"]
    assert merged.max_doc == 4000 and fd.vectors.shape == (4000, 32) and np.array_equal(fd.vectors, base)
    assert np.array_equal(fd.doc_map.graph_node_ids_to_doc_ids, np.arange(4000))
    lead = segs[0].fields["vec"]
    assert fd.entry_node == lead.entry_node                       # the leading graph was extended, not rebuilt
    assert np.array_equal(fd.adjacency, oracle.graph_extend(base, lead.adjacency, lead.entry_node, oracle.SIM_EUCLIDEAN))
    assert fd.pq_codes is not None and fd.pq_codes.shape[0] == 4000
    r = jv.JVectorReader(merged)
    try:
        ix = r.field_index("vec")
        res = ix.search(queries, 10, 50)
        gd, _, _ = ix.exact_topk(queries, 10)
        assert recall(res.docs, gd) >= 0.95
    finally:
        r.close()
    # a leading segment with deleted documents: rebuilt from the live vectors (the reference's fallback)
    live0 = np.ones(2500, bool)
    live0[[3, 77]] = False
    m2 = w.merge(segs, live_docs=[live0, None, None])
    f2 = m2.fields["vec"]
    assert m2.max_doc == 3998 and f2.vectors.shape[0] == 3998
    keep = np.ones(4000, bool)
    keep[[3, 77]] = False
    assert np.array_equal(f2.vectors, base[keep])
    want_adj, want_entry = oracle.graph_build(base[keep], oracle.SIM_EUCLIDEAN, 16, 100)
    assert f2.entry_node == want_entry and np.array_equal(f2.adjacency, want_adj)
