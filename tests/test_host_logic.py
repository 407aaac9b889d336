"""Host-side mirror of the reference's codec glue (collectors, doc map, defaults).  CPU only."""
import numpy as np
import pytest


def test_default_num_subspaces_matches_reference_table(jv, oracle):
    for d in list(range(1, 70)) + [96, 100, 128, 200, 201, 256, 384, 400, 401, 512, 768, 769, 960, 1024, 1536, 1537, 2048, 4096]:
        assert jv.default_num_subspaces(d) == oracle.default_num_subspaces(d)
    assert jv.default_num_subspaces(768) == 192 and jv.default_num_subspaces(96) == 48
    assert jv.default_num_subspaces(1536) == 192 and jv.default_num_subspaces(128) == 64


def test_top_knn_collector_ties_prefer_lower_doc(jv):
    c = jv.TopKnnCollector(3)
    for doc, s in [(7, 0.5), (3, 0.5), (9, 0.9), (1, 0.5), (4, 0.1)]:
        c.collect(doc, s)
    td = c.top_docs()
    assert [(sd.doc, sd.score) for sd in td] == [(9, 0.9), (1, 0.5), (3, 0.5)]
    assert c.min_competitive_similarity() == 0.5


def test_jvector_knn_collector_defaults(jv):
    c = jv.JVectorKnnCollector(jv.TopKnnCollector(10))
    assert c.k() == 10 and c.over_query_factor == 5 and c.threshold == 0.0 and c.rerank_floor == 0.0
    c.inc_visited_count(12)
    assert c.visited_count() == 12


# GraphNodeIdToDocMapTests.java
def test_graph_node_id_to_doc_map(jv):
    m = jv.GraphNodeIdToDocMap([4, 2, -1, 0], max_docs=6)
    assert m.get_lucene_doc_id(0) == 4 and m.get_lucene_doc_id(2) == -1
    assert m.get_jvector_node_id(4) == 0 and m.get_jvector_node_id(2) == 1 and m.get_jvector_node_id(0) == 3
    assert m.get_jvector_node_id(1) == -1 and m.get_jvector_node_id(5) == -1
    assert m.max_docs == 6
    with pytest.raises(ValueError):
        jv.GraphNodeIdToDocMap([0, 9], max_docs=5)


def test_make_accept_bits_is_fixedbitset_layout(jv, oracle):
    mask = np.zeros(130, bool)
    mask[[0, 63, 64, 129]] = True
    w = jv.make_accept_bits(mask)
    assert w.dtype == np.uint64 and w.shape == (3,)
    assert int(w[0]) == (1 << 0) | (1 << 63) and int(w[1]) == 1 and int(w[2]) == 1 << 1
    np.testing.assert_array_equal(w, oracle.make_accept_bits(mask))
    assert jv.make_accept_bits(np.stack([mask, ~mask])).shape == (2, 3)


def test_byte_vectors_unsupported(jv):
    w = jv.JVectorWriter()
    w.add_field("f", jv.VectorSimilarityFunction.EUCLIDEAN)
    with pytest.raises(NotImplementedError):
        w.add_value("f", 0, np.zeros(4, np.int8))


def test_similarity_ordinals_match_meta_file(jv):
    V = jv.VectorSimilarityFunction
    assert [V.EUCLIDEAN.jvector_ord, V.DOT_PRODUCT.jvector_ord, V.COSINE.jvector_ord, V.MAXIMUM_INNER_PRODUCT.jvector_ord] == [0, 1, 2, 3]
