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
    assert np.array_equal(fd.pq_codebooks, lead.pq_codebooks)      # mergePQ keeps the leading codebooks and re-encodes everything
    assert np.array_equal(fd.pq_codes, oracle.pq_encode(base, fd.pq_m, fd.pq_k, lead.pq_codebooks, lead.pq_global_centroid))
    assert np.array_equal(fd.pq_codes[:2500], lead.pq_codes)
    r = jv.JVectorReader(merged)
    try:
        ix = r.field_index("vec")
        res = ix.search(queries, 10, 50)
        gd, _, _ = ix.exact_topk(queries, 10)
        assert recall(res.docs, gd) >= 0.95
    finally:
        r.close()
    # a leading segment with deleted documents: extend, then markNodeDeleted + cleanup, then compact the ordinals
    live0 = np.ones(2500, bool)
    live0[[3, 77, lead.entry_node]] = False
    m2 = w.merge(segs, live_docs=[live0, None, None])
    f2 = m2.fields["vec"]
    keep = np.ones(4000, bool)
    keep[[3, 77, lead.entry_node]] = False
    assert m2.max_doc == 3997 and np.array_equal(f2.vectors, base[keep])
    assert np.array_equal(f2.doc_map.graph_node_ids_to_doc_ids, np.arange(3997))
    ext = oracle.graph_extend(base, lead.adjacency, lead.entry_node, oracle.SIM_EUCLIDEAN)
    cons, e2 = oracle.graph_remove_deleted(base, ext, lead.entry_node, ~keep, oracle.SIM_EUCLIDEAN)
    to_final = np.full(4000, -1)
    to_final[keep] = np.arange(3997)
    want = np.where(cons[keep] >= 0, to_final[np.maximum(cons[keep], 0)], -1)
    assert keep[e2] and f2.entry_node == to_final[e2] and np.array_equal(f2.adjacency, want)
    r = jv.JVectorReader(m2)
    try:
        ix = r.field_index("vec")
        res = ix.search(queries, 10, 50)
        gd, _, _ = ix.exact_topk(queries, 10)
        assert recall(res.docs, gd) >= 0.95
    finally:
        r.close()


def _compact(adj, entry, dead):
    live = ~dead
    to_final = np.full(len(dead), -1)
    to_final[live] = np.arange(int(live.sum()))
    out = np.where(adj[live] >= 0, to_final[np.maximum(adj[live], 0)], -1).astype(np.int32)
    order = np.argsort(out < 0, axis=1, kind="stable")
    return np.take_along_axis(out, order, 1), int(to_final[entry])


@pytest.mark.parametrize("frac", [0.02, 0.3])
def test_oracle_remove_deleted_keeps_graph_searchable(oracle, frac):
    """simpleDeletionTest / merge-with-deletes scenarios (JVectorWriterMergeTests.java:293-462): after consolidation no live row
    points at a deleted node, deleted rows are empty, and the compacted graph keeps the recall floor."""
    base, queries = clustered(3000, 32, 50, seed=9)
    adj, entry = oracle.graph_build(base, oracle.SIM_EUCLIDEAN, 16, 100)
    dead = np.random.default_rng(2).random(3000) < frac
    dead[entry] = True                                             # the entry node goes too
    out, e2 = oracle.graph_remove_deleted(base, adj, entry, dead, oracle.SIM_EUCLIDEAN)
    assert (out[dead] == -1).all() and not dead[e2]
    rows = out[~dead]
    assert not dead[rows[rows >= 0]].any()
    untouched = ~dead & ~np.isin(adj, np.nonzero(dead)[0]).any(1)
    assert np.array_equal(out[untouched], adj[untouched])          # rows without a deleted neighbour are not rewritten
    cadj, ce = _compact(out, e2, dead)
    ix = oracle.OracleIndex(oracle.SIM_EUCLIDEAN, base[~dead], cadj, ce)
    d, _, _, _ = ix.search(queries, 10, 50)
    gd, _, _ = ix.exact_topk(queries, 10)
    assert recall(d, gd) >= 0.98
    # nothing deleted: identity; everything deleted: empty graph, entry -1
    same, e3 = oracle.graph_remove_deleted(base, adj, entry, np.zeros(3000, bool), oracle.SIM_EUCLIDEAN)
    assert np.array_equal(same, adj) and e3 == entry
    none, e4 = oracle.graph_remove_deleted(base, adj, entry, np.ones(3000, bool), oracle.SIM_EUCLIDEAN)
    assert (none == -1).all() and e4 == -1


@pytest.mark.gpu
@pytest.mark.parametrize("sim_name,frac", [("SIM_EUCLIDEAN", 0.02), ("SIM_DOT", 0.3), ("SIM_COSINE", 0.6)])
def test_gpu_remove_deleted_matches_oracle_bit_for_bit(jv, oracle, sim_name, frac):
    sim = getattr(oracle, sim_name)
    base, _ = clustered(4000, 40, 8, seed=13, normalize=sim_name != "SIM_EUCLIDEAN")
    adj, entry = oracle.graph_build(base, sim, 16, 100)
    dead = np.random.default_rng(4).random(4000) < frac
    dead[entry] = True
    want, we = oracle.graph_remove_deleted(base, adj, entry, dead, sim)
    got, ge = jv.graph_remove_deleted(base, adj, entry, dead, sim)
    assert ge == we and np.array_equal(got, want)
    g0, e0 = jv.graph_remove_deleted(base, adj, entry, np.zeros(4000, bool), sim)
    assert e0 == entry and np.array_equal(g0, adj)
    g1, e1 = jv.graph_remove_deleted(base, adj, entry, np.ones(4000, bool), sim)
    assert e1 == -1 and (g1 == -1).all()
    only = np.ones(4000, bool)
    only[[17, 3000]] = False                                       # two live nodes far from the entry
    g2, e2 = jv.graph_remove_deleted(base, adj, entry, only, sim)
    w2, we2 = oracle.graph_remove_deleted(base, adj, entry, only, sim)
    assert e2 == we2 and np.array_equal(g2, w2)
