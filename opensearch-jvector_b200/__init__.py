"""opensearch-jvector_b200 — B200-native (sm_100a) implementation of the opensearch-jvector query hot
path behind the reference's codec API.  Numeric work happens only in lib/libjvgpu.so (hand-written
CUDA, C-ABI in include/jvgpu.h); importing the package does not need a GPU, calling it does."""
from . import native
from .codec import (DEFAULT_OVER_QUERY_FACTOR, FieldData, GraphNodeIdToDocMap, JVectorIndexQuantization,
                    JVectorKnnCollector, JVectorKnnFloatVectorQuery, JVectorReader, JVectorWriter, KNNCounter, ScoreDoc,
                    Segment, TopKnnCollector, VectorSimilarityFunction, default_num_subspaces)
from . import segment_files
from .index import (GpuIndex, SearchResult, graph_build, graph_build_pq, graph_extend, graph_remove_deleted, make_accept_bits, merge_topk,
                    pq_decode, pq_encode, pq_train)

__all__ = [
    "native", "segment_files", "GpuIndex", "SearchResult", "graph_build", "graph_build_pq", "graph_extend", "graph_remove_deleted", "make_accept_bits", "merge_topk", "pq_decode", "pq_encode", "pq_train",
    "JVectorReader", "JVectorWriter", "JVectorKnnCollector", "JVectorKnnFloatVectorQuery", "TopKnnCollector",
    "GraphNodeIdToDocMap", "JVectorIndexQuantization", "KNNCounter", "ScoreDoc", "Segment", "FieldData",
    "VectorSimilarityFunction", "default_num_subspaces", "DEFAULT_OVER_QUERY_FACTOR",
]
