"""Graph build with PQ build scores (JVectorWriter.java:238-244 at flush, :1143-1151 when a merge rebuilds):
PQ decode and the PQ-scored builder on the GPU against the oracle, and the reference's quantised-flush recall floor through the
writer mirror (KNNJVectorTests.java:1358-1403)."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import clustered, recall

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sim,dim,m", [(O.SIM_EUCLIDEAN, 16, 8), (O.SIM_DOT, 64, 16), (O.SIM_COSINE, 26, 8)])
def test_pq_decode_and_pq_scored_build_match_the_oracle(jv, sim, dim, m):
    base, _ = clustered(3000, dim, 4, seed=dim, normalize=sim == O.SIM_DOT)
    center = sim == O.SIM_EUCLIDEAN
    cb, g = O.pq_train(base, m, 256, center=center, iters=4, seed=3)
    codes = O.pq_encode(base, m, 256, cb, g)
    want = O.pq_decode(codes, dim, 256, cb, g)
    got = jv.pq_decode(codes, dim, 256, cb, g)
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
    wadj, wentry = O.graph_build_pq(codes, dim, 256, cb, g, sim, 16, 100)
    gadj, gentry = jv.graph_build_pq(codes, dim, 256, cb, g, sim, 16, 100)
    assert gentry == wentry
    np.testing.assert_array_equal(gadj, wadj)
    eadj, _ = O.graph_build(base, sim, 16, 100)          # the exact-score graph is a different graph
    assert not np.array_equal(wadj, eadj)


def test_quantised_flush_builds_with_pq_scores_and_keeps_the_recall_floor(jv):
    """testJVectorKnnIndex_simpleCase_withQuantization (KNNJVectorTests.java:1358-1403): 1024 uniform 16-d vectors = exactly the
    minimum batch for quantisation, EUCLIDEAN, one flush, k = 50, recall 1.0 +- 0.05."""
    V = jv.VectorSimilarityFunction
    rng = np.random.default_rng(42)
    n, dim, k = 1024, 16, 50
    vectors = rng.random((n, dim), dtype=np.float32)
    target = np.zeros(dim, np.float32)
    w = jv.JVectorWriter()
    w.add_field("vec", V.EUCLIDEAN)
    for i in range(n):
        w.add_value("vec", i, vectors[i])
    seg = w.flush(n)
    fd = seg.fields["vec"]
    assert fd.pq_codes is not None
    # the flushed graph is the PQ-scored one (JVectorWriter.java:238-244), not the exact-score one
    padj, pentry = jv.graph_build_pq(fd.pq_codes, dim, fd.pq_k, fd.pq_codebooks, fd.pq_global_centroid, V.EUCLIDEAN.jvector_ord, w.max_conn, w.beam_width)
    np.testing.assert_array_equal(fd.adjacency, padj)
    assert fd.entry_node == pentry
    truth = np.argsort(((vectors - target) ** 2).sum(1), kind="stable")[:k]
    reader = jv.JVectorReader(seg)
    col = jv.JVectorKnnCollector(jv.TopKnnCollector(k), 0.0, 0.0, 5)
    reader.search("vec", target, col)
    found = np.array([sd.doc for sd in col.top_docs()], np.int64)
    reader.close()
    assert len(found) == k
    assert recall(found[None, :], truth[None, :]) >= 0.95


def test_two_quantised_flushes_force_merged_keep_the_recall_floor(jv):
    """testJVectorKnnIndex_happyCase_withQuantization_multipleSegments (KNNJVectorTests.java:1471-1530) through the writer mirror:
    two flushes of exactly the minimum batch (each quantised, PQ-scored graph), force-merged (leading-segment merge + mergePQ with
    the leading codebooks), k = 50, recall 1.0 +- 0.05."""
    V = jv.VectorSimilarityFunction
    dim, per, k = 16, 1024, 50
    vectors = O.java_random_vectors(2 * per, dim, 1)
    target = np.zeros(dim, np.float32)
    segs = []
    for s in range(2):
        w = jv.JVectorWriter()
        w.add_field("test_field", V.EUCLIDEAN)
        for i in range(per):
            w.add_value("test_field", i, vectors[s * per + i])
        segs.append(w.flush(per))
        assert segs[-1].fields["test_field"].pq_codes is not None
    merged = jv.JVectorWriter().merge(segs)
    fd = merged.fields["test_field"]
    assert fd.vectors.shape[0] == 2 * per and fd.pq_codes is not None
    # equal live counts: the LATER segment leads (`>=`, JVectorWriter.java:805-808); its codebooks are kept ("not refining PQ codes on merge")
    np.testing.assert_array_equal(fd.pq_codebooks, segs[1].fields["test_field"].pq_codebooks)
    # docIds are re-based in the original segment order whichever segment leads: doc d = vectors[d]
    truth = np.argsort(((vectors - target) ** 2).sum(1), kind="stable")[:k]
    reader = jv.JVectorReader(merged)
    col = jv.JVectorKnnCollector(jv.TopKnnCollector(k), 0.0, 0.0, 5)
    reader.search("test_field", target, col)
    found = np.array([sd.doc for sd in col.top_docs()], np.int64)
    reader.close()
    assert len(found) == k and recall(found[None, :], truth[None, :]) >= 0.95


def test_merge_picks_the_segment_with_the_most_live_vectors_as_leading_and_keeps_every_field(jv):
    """JVectorWriter.java:784-848: the leading reader is the one with the most live vectors (ties -> the later one), swapped to
    position 0; docIds are re-based in the original segment order.  Fields missing from the first segment are not dropped."""
    V = jv.VectorSimilarityFunction
    rng = np.random.default_rng(12)
    small, big = rng.random((60, 8), dtype=np.float32), rng.random((500, 8), dtype=np.float32)
    extra = rng.random((40, 8), dtype=np.float32)

    def flush(fields, n):
        w = jv.JVectorWriter()
        for name, vecs in fields.items():
            w.add_field(name, V.EUCLIDEAN)
            for i, v in enumerate(vecs):
                w.add_value(name, i, v)
        return w.flush(n)

    s0, s1 = flush({"a": small}, 60), flush({"a": big, "b": extra}, 500)
    merged = jv.JVectorWriter().merge([s0, s1])
    assert set(merged.fields) == {"a", "b"} and merged.max_doc == 560
    a = merged.fields["a"]
    # leading = segment 1 (500 live vectors): its vectors come first in the merged ordinal space, with docIds 60..559
    np.testing.assert_array_equal(a.vectors[:500], big)
    np.testing.assert_array_equal(a.doc_map.graph_node_ids_to_doc_ids[:500], np.arange(60, 560))
    np.testing.assert_array_equal(a.doc_map.graph_node_ids_to_doc_ids[500:], np.arange(0, 60))
    # and the leading graph was kept: the first 500 rows still contain the leading segment's edges among themselves
    lead_adj = s1.fields["a"].adjacency
    kept = np.mean([len(set(lead_adj[i][lead_adj[i] >= 0]) & set(a.adjacency[i][a.adjacency[i] >= 0])) / max(1, (lead_adj[i] >= 0).sum()) for i in range(0, 500, 7)])
    assert kept > 0.7
    b = merged.fields["b"]
    np.testing.assert_array_equal(b.doc_map.graph_node_ids_to_doc_ids, np.arange(60, 100))
    reader = jv.JVectorReader(merged)
    col = jv.JVectorKnnCollector(jv.TopKnnCollector(5), 0.0, 0.0, 5)
    reader.search("a", small[7], col)
    assert col.top_docs()[0].doc == 7
    reader.close()


def test_a_merge_can_start_from_segment_files(jv, tmp_path):
    """Flush two segments, persist them in the reference's file layout, read them back with the loader and merge from the FILES:
    same merged segment as merging the in-memory segments (the neighbours-score cache is recomputed on the device, not read)."""
    V = jv.VectorSimilarityFunction
    rng = np.random.default_rng(5)
    vecs = rng.random((1500, 12), dtype=np.float32)
    segs, from_files = [], []
    for s_i, (a, b) in enumerate([(0, 1100), (1100, 1500)]):
        w = jv.JVectorWriter()
        w.add_field("vec", V.EUCLIDEAN)
        for i in range(a, b):
            w.add_value("vec", i - a, vecs[i])
        seg = w.flush(b - a)
        segs.append(seg)
        d = tmp_path / f"seg{s_i}"
        d.mkdir()
        jv.JVectorWriter.write(seg, d, field_numbers={"vec": 4})
        from_files.append(jv.Segment.from_files(d, {4: "vec"}, max_doc=b - a))
        np.testing.assert_array_equal(from_files[-1].fields["vec"].adjacency, seg.fields["vec"].adjacency)
    live = [np.ones(1100, bool), np.ones(400, bool)]
    live[0][::9] = False                                   # deleted documents in the leading segment
    want = jv.JVectorWriter().merge(segs, live)
    got = jv.JVectorWriter().merge(from_files, live)
    for f in ("vectors", "adjacency", "pq_codes", "pq_codebooks"):
        np.testing.assert_array_equal(getattr(got.fields["vec"], f), getattr(want.fields["vec"], f))
    assert got.fields["vec"].entry_node == want.fields["vec"].entry_node and got.max_doc == want.max_doc
