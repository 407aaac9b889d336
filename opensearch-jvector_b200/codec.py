"""Host-side mirror of the reference's codec adapter for the hot path (Python because no JVM exists in
this image; the Java binding a maintainer would add is in INTEGRATION.md).

Same names, argument meaning and error behaviour as
  org.opensearch.knn.index.codec.jvector.{JVectorReader, JVectorWriter, JVectorKnnCollector,
  JVectorKnnFloatVectorQuery, GraphNodeIdToDocMap, JVectorIndexQuantization, JVectorFormat}
so the parity tests read like the reference's own (KNNJVectorTests.java).  Everything numeric goes
through libjvgpu's C-ABI; nothing here computes scores on the CPU.
"""
from __future__ import annotations

import enum
import heapq
import threading
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import native as N
from .index import GpuIndex, graph_build, graph_build_pq, graph_extend, graph_remove_deleted, make_accept_bits, pq_encode, pq_train

# ---- constants (KNNConstants.java:83-114, JVectorFormat.java:22-35) -------------------------------
DEFAULT_MAX_CONN = 32
DEFAULT_BEAM_WIDTH = 100
DEFAULT_NEIGHBOR_OVERFLOW = 1.2
DEFAULT_ALPHA = 1.2
DEFAULT_MINIMUM_BATCH_SIZE_FOR_QUANTIZATION = 1024
DEFAULT_OVER_QUERY_FACTOR = 5
DEFAULT_QUERY_SIMILARITY_THRESHOLD = 0.0
DEFAULT_QUERY_RERANK_FLOOR = 0.0
PQ_TRAINING_SAMPLE_LIMIT = 128_000  # jVector ProductQuantization.MAX_PQ_TRAINING_SET_SIZE
PQ_LLOYD_ITERATIONS = 6


class VectorSimilarityFunction(enum.Enum):
    """Lucene's enum (same ordinals as JV_SIM_*); `.meta_ord` is the meta-file simOrd (JVectorReader.java:389-409)."""

    EUCLIDEAN = N.SIM_EUCLIDEAN
    DOT_PRODUCT = N.SIM_DOT
    COSINE = N.SIM_COSINE
    MAXIMUM_INNER_PRODUCT = N.SIM_MIP

    @property
    def jvector_ord(self) -> int:
        """JV_SIM_* of the library = Lucene's enum ordinal."""
        return self.value

    @property
    def meta_ord(self) -> int:
        """simOrd of the meta record: VectorSimilarityMapper.distFuncToOrd (JVectorReader.java:407-413) = indexOf in
        [EUCLIDEAN, DOT_PRODUCT, COSINE, DOT_PRODUCT], so MAXIMUM_INNER_PRODUCT is written as 1."""
        return N.SIM_DOT if self is VectorSimilarityFunction.MAXIMUM_INNER_PRODUCT else self.value


def default_num_subspaces(original_dimension: int) -> int:
    """JVectorIndexQuantization.PQ.defaultNumSubspaces, JVectorIndexQuantization.java:428-446."""
    d = original_dimension
    if d <= 32:
        return d
    if d <= 64:
        return 32
    if d <= 200:
        return int(d * 0.5)
    if d <= 400:
        return 100
    if d <= 768:
        return int(d * 0.25)
    if d <= 1536:
        return 192
    return int(d * 0.125)


class GraphNodeIdToDocMap:
    """GraphNodeIdToDocMap.java: ordinal <-> Lucene docId arrays."""

    NO_VECTOR_OR_DELETED_DOC = -1

    def __init__(self, ordinals_to_doc_ids: Sequence[int], max_docs: Optional[int] = None):
        self.graph_node_ids_to_doc_ids = np.ascontiguousarray(ordinals_to_doc_ids, dtype=np.int32)
        live = self.graph_node_ids_to_doc_ids[self.graph_node_ids_to_doc_ids >= 0]
        self.max_docs = int(max_docs) if max_docs is not None else (int(live.max()) + 1 if live.size else 0)
        if live.size and int(live.max()) >= self.max_docs:
            raise ValueError("docId exceeds maxDocs")
        self.doc_ids_to_graph_node_ids = np.full(self.max_docs, -1, dtype=np.int32)
        ords = np.nonzero(self.graph_node_ids_to_doc_ids >= 0)[0]
        self.doc_ids_to_graph_node_ids[self.graph_node_ids_to_doc_ids[ords]] = ords

    def get_lucene_doc_id(self, graph_node_id: int) -> int:
        return int(self.graph_node_ids_to_doc_ids[graph_node_id])

    def get_jvector_node_id(self, doc_id: int) -> int:
        return int(self.doc_ids_to_graph_node_ids[doc_id])


# ---- collectors (Lucene TopKnnCollector + JVectorKnnCollector.java) -------------------------------
@dataclass(order=True)
class ScoreDoc:
    score: float
    doc: int


class TopKnnCollector:
    """Lucene TopKnnCollector: bounded min-heap of (score, doc); ties prefer the lower docId."""

    def __init__(self, k: int, visit_limit: int = 2**31 - 1):
        self._k = k
        self._heap: List[tuple] = []  # (score, -doc): worst on top
        self._visited = 0
        self._visit_limit = visit_limit

    def k(self) -> int:
        return self._k

    def collect(self, doc_id: int, similarity: float) -> bool:
        item = (float(similarity), -int(doc_id))
        if len(self._heap) < self._k:
            heapq.heappush(self._heap, item)
            return True
        if item > self._heap[0]:
            heapq.heapreplace(self._heap, item)
            return True
        return False

    def inc_visited_count(self, count: int) -> None:
        self._visited += count

    def visited_count(self) -> int:
        return self._visited

    def visit_limit(self) -> int:
        return self._visit_limit

    def early_terminated(self) -> bool:
        return self._visited >= self._visit_limit

    def min_competitive_similarity(self) -> float:
        return self._heap[0][0] if len(self._heap) >= self._k else float("-inf")

    def top_docs(self) -> List[ScoreDoc]:
        return [ScoreDoc(s, -nd) for s, nd in sorted(self._heap, reverse=True)]


@dataclass
class JVectorKnnCollector:
    """JVectorKnnCollector.java:16-21 — carries threshold / rerankFloor / overQueryFactor to the reader."""

    delegate: TopKnnCollector
    threshold: float = DEFAULT_QUERY_SIMILARITY_THRESHOLD
    rerank_floor: float = DEFAULT_QUERY_RERANK_FLOOR
    over_query_factor: int = DEFAULT_OVER_QUERY_FACTOR

    def k(self):
        return self.delegate.k()

    def collect(self, doc_id, similarity):
        return self.delegate.collect(doc_id, similarity)

    def inc_visited_count(self, count):
        self.delegate.inc_visited_count(count)

    def visited_count(self):
        return self.delegate.visited_count()

    def top_docs(self):
        return self.delegate.top_docs()


class KNNCounter:
    """plugin/stats/KNNCounter.java:30-37 — the counters JVectorReader.java:189-193 bumps."""

    _lock = threading.Lock()
    KNN_QUERY_VISITED_NODES = 0
    KNN_QUERY_RERANKED_COUNT = 0
    KNN_QUERY_EXPANDED_NODES = 0
    KNN_QUERY_EXPANDED_BASE_LAYER_NODES = 0
    KNN_QUERY_GRAPH_SEARCH_TIME = 0.0   # milliseconds (:192; here the device time of the batch, added once per batch)
    KNN_QUANTIZATION_TRAINING_TIME = 0.0  # milliseconds (JVectorIndexQuantization.java:132)

    @classmethod
    def add_search_time(cls, millis: float):
        with cls._lock:
            cls.KNN_QUERY_GRAPH_SEARCH_TIME += float(millis)

    @classmethod
    def add_training_time(cls, millis: float):
        with cls._lock:
            cls.KNN_QUANTIZATION_TRAINING_TIME += float(millis)

    @classmethod
    def add(cls, visited, reranked, expanded, expanded_base):
        with cls._lock:
            cls.KNN_QUERY_VISITED_NODES += int(visited)
            cls.KNN_QUERY_RERANKED_COUNT += int(reranked)
            cls.KNN_QUERY_EXPANDED_NODES += int(expanded)
            cls.KNN_QUERY_EXPANDED_BASE_LAYER_NODES += int(expanded_base)


# ---- segment container ------------------------------------------------------------------------------
@dataclass
class FieldData:
    """Decoded arrays of one field of one flushed/merged segment (what OnDiskGraphIndex + PQVectors + the
    meta record hold; persisted layout: SURVEY Appendix B)."""

    similarity: VectorSimilarityFunction
    vectors: np.ndarray                     # [n, dim] fp32, ordinal order (InlineVectors)
    adjacency: np.ndarray                   # [n, R] int32, -1 padded
    entry_node: int
    doc_map: GraphNodeIdToDocMap
    pq_m: int = 0
    pq_k: int = 0
    pq_codebooks: Optional[np.ndarray] = None
    pq_global_centroid: Optional[np.ndarray] = None
    pq_codes: Optional[np.ndarray] = None


@dataclass
class Segment:
    max_doc: int
    fields: Dict[str, FieldData] = field(default_factory=dict)

    @classmethod
    def from_files(cls, directory, field_infos: Dict[int, object], max_doc: int, segment_name: str = "_0",
                   segment_suffix: str = "JVector_0", load_flags: int = 0) -> "Segment":
        """The decoded arrays of a persisted segment (meta file + one data file per field) — what a merge needs
        from its input readers (JVectorWriter.mergeOneField reads graph, vectors and PQ of every segment through their readers,
        JVectorWriter.java:1040-1160).  `field_infos`: field number -> name or (name, VectorSimilarityFunction), as for
        JVectorReader.open.  The neighbours-score-cache file of the leading segment (:1174) is not read: the merge recomputes the
        cached neighbour scores on the device (jv_graph_extend).  No GPU is needed here."""
        from pathlib import Path as _Path

        from .segment_files import META_EXTENSION, SegmentFiles, field_data_file_name, segment_file_name
        directory = _Path(directory)
        seg = cls(max_doc=int(max_doc))
        with SegmentFiles(directory / segment_file_name(segment_name, segment_suffix, META_EXTENSION), load_flags) as sf:
            for i, m in enumerate(sf.metas):
                info = field_infos[m.field_number]
                name, sim = (info if isinstance(info, tuple) else (info, None))
                if sim is not None:
                    sf.set_lucene_similarity(i, sim.jvector_ord)
                d = sf.load_field(i, directory / field_data_file_name(segment_name, segment_suffix, name), load_flags)
                fd = FieldData(VectorSimilarityFunction(sf.metas[i].similarity), d["vectors"], d["adjacency"], d["entry_node"],
                               GraphNodeIdToDocMap(sf.doc_map(i), m.max_doc))
                if d["pq_m"] > 0:
                    fd.pq_m, fd.pq_k, fd.pq_codebooks = d["pq_m"], d["pq_k"], d["pq_codebooks"]
                    fd.pq_global_centroid, fd.pq_codes = d["pq_global_centroid"], d["pq_codes"]
                seg.fields[name] = fd
        return seg


class JVectorIndexQuantization:
    """PQ strategy of JVectorIndexQuantization.java:114-140 on the GPU (train = K8, encode = K6)."""

    @staticmethod
    def compute_pq_vectors(vectors: np.ndarray, similarity: VectorSimilarityFunction, num_subspaces: int, device: int = 0,
                           seed: int = 0):
        n = vectors.shape[0]
        clusters = min(256, n)                                        # :122
        center = similarity == VectorSimilarityFunction.EUCLIDEAN     # :127
        sample = vectors
        if n > PQ_TRAINING_SAMPLE_LIMIT:
            rng = np.random.default_rng(seed)
            sample = vectors[np.sort(rng.choice(n, PQ_TRAINING_SAMPLE_LIMIT, replace=False))]
        import time as _time
        t0 = _time.perf_counter()
        codebooks, gcent = pq_train(sample, num_subspaces, clusters, center, PQ_LLOYD_ITERATIONS, seed, device)
        KNNCounter.add_training_time((_time.perf_counter() - t0) * 1e3)                   # :132
        codes = pq_encode(vectors, num_subspaces, clusters, codebooks, gcent, device)     # :133
        return num_subspaces, clusters, codebooks, gcent, codes


class JVectorWriter:
    """Flush path of JVectorWriter.java:216-283: buffer vectors, quantise when n >= minBatch, build the
    Vamana graph, hand back the decoded segment arrays.  Graph build and PQ training run on the GPU through
    the C-ABI "next" entry points (jv_graph_build / jv_pq_train)."""

    def __init__(self, max_conn: int = DEFAULT_MAX_CONN, beam_width: int = DEFAULT_BEAM_WIDTH,
                 neighbor_overflow: float = DEFAULT_NEIGHBOR_OVERFLOW, alpha: float = DEFAULT_ALPHA,
                 min_batch_size_for_quantization: int = DEFAULT_MINIMUM_BATCH_SIZE_FOR_QUANTIZATION,
                 num_pq_subspaces=default_num_subspaces, device: int = 0):
        self.max_conn, self.beam_width = max_conn, beam_width
        self.neighbor_overflow, self.alpha = neighbor_overflow, alpha
        self.min_batch = min_batch_size_for_quantization
        self.num_pq_subspaces = num_pq_subspaces
        self.device = device
        self._fields: Dict[str, dict] = {}

    def add_field(self, name: str, similarity: VectorSimilarityFunction):
        self._fields[name] = {"sim": similarity, "docs": [], "vecs": []}

    def add_value(self, name: str, doc_id: int, vector: Sequence[float]):
        f = self._fields[name]
        v = np.asarray(vector)
        if v.dtype.kind not in "fiu" or v.dtype == np.uint8 or v.dtype == np.int8:
            raise NotImplementedError("Byte vectors are not supported by jVector (JVectorWriter.java:176-184)")
        f["docs"].append(int(doc_id))
        f["vecs"].append(np.asarray(v, dtype=np.float32))

    def flush(self, max_doc: int) -> Segment:
        seg = Segment(max_doc=max_doc)
        for name, f in self._fields.items():
            vecs = np.stack(f["vecs"]).astype(np.float32) if f["vecs"] else np.zeros((0, 1), np.float32)
            sim: VectorSimilarityFunction = f["sim"]
            doc_map = GraphNodeIdToDocMap(f["docs"], max_doc)
            n = vecs.shape[0]
            fd = FieldData(sim, vecs, np.zeros((n, self.max_conn), np.int32), 0, doc_map)
            if n >= self.min_batch:                                   # JVectorWriter.java:267-279: quantise first ...
                m = self.num_pq_subspaces(vecs.shape[1])
                fd.pq_m, fd.pq_k, fd.pq_codebooks, fd.pq_global_centroid, fd.pq_codes = \
                    JVectorIndexQuantization.compute_pq_vectors(vecs, sim, m, self.device)
                # ... then getGraph(quantizationResult.buildScoreProvider(), ..) (:238-244): PQ build scores
                fd.adjacency, fd.entry_node = graph_build_pq(fd.pq_codes, vecs.shape[1], fd.pq_k, fd.pq_codebooks, fd.pq_global_centroid,
                                                             sim.jvector_ord, self.max_conn, self.beam_width, self.neighbor_overflow,
                                                             self.alpha, self.device)
            elif n > 0:                                               # randomAccessScoreProvider (:274-278): exact build scores
                fd.adjacency, fd.entry_node = graph_build(vecs, sim.jvector_ord, self.max_conn, self.beam_width,
                                                          self.neighbor_overflow, self.alpha, self.device)
            seg.fields[name] = fd
        return seg

    def merge(self, segments: Sequence[Segment], live_docs: Optional[Sequence[Optional[np.ndarray]]] = None) -> Segment:
        """mergeOneField over whole segments (JVectorWriter.java:1040-1160): the merged segment holds the live documents of
        `segments` in order (docIds re-based like Lucene's MergeState doc maps: segment i starts after the live docs of the
        segments before it).  Per field the segment with the most live vectors is the LEADING one (:784-848): its graph is kept, the
        other segments' live vectors are inserted into it (tryLeadingSegmentMerge, :1166-1341 -> jv_graph_extend), its deleted nodes are consolidated away
        (markNodeDeleted + cleanup -> jv_graph_remove_deleted) and the ordinals are compacted as the on-disk writer does.  PQ: the
        leading segment's codebooks are reused and all merged vectors re-encoded (mergePQ, :1072-1124), or trained when it has none.
        `live_docs[i]`: bool mask over segment i's docIds (None = all live)."""
        live_docs = list(live_docs) if live_docs is not None else [None] * len(segments)
        merged = Segment(max_doc=0)
        bases, base = [], 0
        remap = []                                   # per segment: old docId -> new docId (-1 = deleted)
        for seg, live in zip(segments, live_docs):
            mask = np.ones(seg.max_doc, bool) if live is None else np.asarray(live, bool)
            new_ids = np.full(seg.max_doc, -1, np.int64)
            new_ids[mask] = base + np.arange(int(mask.sum()))
            remap.append(new_ids)
            bases.append(base)
            base += int(mask.sum())
        merged.max_doc = base
        names = []
        for seg in segments:                         # the union of the segments' vector fields, first appearance first
            for name in seg.fields:
                if name not in names:
                    names.append(name)
        for name in names:
            # the LEADING reader is the one with the most live vectors, ties -> the later one, and it is swapped (not rotated) to
            # position 0 (JVectorWriter.java:784-848); docIds are still re-based in the original segment order (MergeState doc maps)
            keeps = {}
            for si, seg in enumerate(segments):
                fd = seg.fields.get(name)
                if fd is None:
                    continue
                ords = fd.doc_map.graph_node_ids_to_doc_ids
                new_docs = np.where(ords >= 0, remap[si][np.maximum(ords, 0)], -1) if len(ords) else np.zeros(0, np.int64)
                keeps[si] = (fd, new_docs, new_docs >= 0)
            if not keeps:
                continue
            lead_idx, lead_live = -1, -1
            for si, (_, _, keep) in keeps.items():
                if int(keep.sum()) >= lead_live:
                    lead_idx, lead_live = si, int(keep.sum())
            order = sorted(keeps)
            pos = order.index(lead_idx)
            order[0], order[pos] = order[pos], order[0]
            lead = keeps[order[0]][0]
            vec_parts, doc_parts = [], []
            lead_keep = keeps[order[0]][2]
            for si in order:
                fd, new_docs, keep = keeps[si]
                vec_parts.append(fd.vectors[keep])
                doc_parts.append(new_docs[keep])
            vecs = np.concatenate(vec_parts).astype(np.float32)
            docs = np.concatenate(doc_parts).astype(np.int32)
            n = vecs.shape[0]
            out = FieldData(lead.similarity, vecs, np.zeros((n, self.max_conn), np.int32), 0, GraphNodeIdToDocMap(docs, merged.max_doc))
            n0 = lead.vectors.shape[0]
            sim_ord = lead.similarity.jvector_ord
            rebuild = False                                           # no usable leading graph: build from scratch (below)
            if n > 0:
                if n0 > 0 and lead_keep.any() and lead.adjacency.shape[1] == self.max_conn:
                    # heap ordinal space of the reference (:1240-1262): every leading node, deleted ones included, then the others
                    heap_vecs = np.concatenate([lead.vectors] + vec_parts[1:]).astype(np.float32)
                    adj, entry = lead.adjacency, lead.entry_node
                    if heap_vecs.shape[0] > n0:                       # builder.addGraphNode for the other segments' vectors
                        adj = graph_extend(heap_vecs, adj, entry, sim_ord, self.beam_width, self.neighbor_overflow, self.alpha, self.device)
                    if not lead_keep.all():                           # builder.markNodeDeleted + cleanup (:1318-1327)
                        dead = np.zeros(heap_vecs.shape[0], bool)
                        dead[:n0] = ~lead_keep
                        adj, entry = graph_remove_deleted(heap_vecs, adj, entry, dead, sim_ord, self.alpha, self.device)
                        live = ~dead                                  # the on-disk writer compacts ordinals, order preserved
                        to_final = np.full(heap_vecs.shape[0], -1, np.int64)
                        to_final[live] = np.arange(int(live.sum()))
                        adj = np.where(adj[live] >= 0, to_final[np.maximum(adj[live], 0)], -1).astype(np.int32)
                        entry = int(to_final[entry])
                    out.adjacency, out.entry_node = np.ascontiguousarray(adj, np.int32), entry
                else:
                    rebuild = True
            if n > 0 and lead.pq_codes is not None:
                # mergePQ, :1072-1124: the leading reader's codebooks are kept as they are ("We are not refining PQ codes on
                # merge presently") and every merged vector is re-encoded with them: PQVectors.encodeAndBuild = K6
                out.pq_m, out.pq_k, out.pq_codebooks, out.pq_global_centroid = lead.pq_m, lead.pq_k, lead.pq_codebooks, lead.pq_global_centroid
                out.pq_codes = pq_encode(vecs, lead.pq_m, lead.pq_k, lead.pq_codebooks, lead.pq_global_centroid, self.device)
            elif n >= self.min_batch:                                  # no codebooks yet: computePqVectors over the merged vectors
                m = self.num_pq_subspaces(vecs.shape[1])
                out.pq_m, out.pq_k, out.pq_codebooks, out.pq_global_centroid, out.pq_codes = \
                    JVectorIndexQuantization.compute_pq_vectors(vecs, lead.similarity, m, self.device)
            if rebuild:
                if out.pq_codes is not None:                          # "PQ codebooks found, building graph from scratch with PQ vectors", :1143-1151
                    out.adjacency, out.entry_node = graph_build_pq(out.pq_codes, vecs.shape[1], out.pq_k, out.pq_codebooks, out.pq_global_centroid,
                                                                   sim_ord, self.max_conn, self.beam_width, self.neighbor_overflow,
                                                                   self.alpha, self.device)
                else:                                                 # randomAccessScoreProvider, :1139-1141
                    out.adjacency, out.entry_node = graph_build(vecs, sim_ord, self.max_conn, self.beam_width,
                                                                self.neighbor_overflow, self.alpha, self.device)
            merged.fields[name] = out
        return merged

    @staticmethod
    def write(segment: Segment, directory, segment_name: str = "_0", segment_suffix: str = "JVector_0",
              field_numbers: Optional[Dict[str, int]] = None, **kw):
        """Persist a flushed segment in the reference's file layout (JVectorWriter.java:134-165,383-433,469-510,
        573-577).  Returns {"meta": path, field name: data path}."""
        from .segment_files import write_segment
        return write_segment(segment, directory, segment_name, segment_suffix, field_numbers=field_numbers, **kw)


class JVectorReader:
    """JVectorReader.java — per-segment reader; `search` is the drop-in for JVectorReader.java:130-210."""

    def __init__(self, segment: Segment, device: int = 0, flags: int = N.FLAG_LUT_U8):
        """`flags`: JV_INDEX_FLAG_* of every field's device index; the default is the production configuration (8-bit ADC tables;
        segments the 8-bit path does not cover — un-quantised, K < 256, non-uniform sub-vectors — ignore the flag)."""
        self._segment = segment
        self._entries: Dict[str, GpuIndex] = {}
        self._closed = False
        for name, fd in segment.fields.items():   # FieldEntry ctor, :284-337
            self._entries[name] = GpuIndex(
                fd.similarity.jvector_ord, fd.vectors, fd.adjacency, fd.entry_node,
                ord_to_doc=fd.doc_map.graph_node_ids_to_doc_ids, max_doc=segment.max_doc, pq_m=fd.pq_m, pq_k=fd.pq_k,
                pq_codebooks=fd.pq_codebooks, pq_global_centroid=fd.pq_global_centroid, pq_codes=fd.pq_codes,
                device=device, flags=flags)

    @classmethod
    def open(cls, directory, field_infos: Dict[int, object], segment_name: str = "_0", segment_suffix: str = "JVector_0",
             device: int = 0, flags: int = N.FLAG_LUT_U8, load_flags: int = 0) -> "JVectorReader":
        """JVectorReader(SegmentReadState), JVectorReader.java:52-81: read the meta file, then one FieldEntry per record
        (:255-337) — here a single native call per field (jv_segment_index_create) that parses the field data file and
        copies graph, vectors, PQ codebooks + codes and doc map to the device.  `field_infos` maps Lucene field numbers to
        names, or to (name, VectorSimilarityFunction) pairs (FieldInfos lives outside this codec).  The pair form is REQUIRED for
        MAXIMUM_INNER_PRODUCT fields: their meta record says DOT_PRODUCT (distFuncToOrd) and only FieldInfo knows better."""
        from pathlib import Path as _Path

        from .segment_files import META_EXTENSION, SegmentFiles, field_data_file_name, segment_file_name
        directory = _Path(directory)
        self = object.__new__(cls)
        self._segment, self._entries, self._closed = None, {}, False
        self._files: Dict[str, tuple] = {}
        self._seg_files = SegmentFiles(directory / segment_file_name(segment_name, segment_suffix, META_EXTENSION), load_flags)
        try:
            for i, m in enumerate(self._seg_files.metas):
                name = field_infos[m.field_number]
                if isinstance(name, tuple):
                    name, lucene_sim = name
                    self._seg_files.set_lucene_similarity(i, lucene_sim.jvector_ord)
                path = directory / field_data_file_name(segment_name, segment_suffix, name)
                self._entries[name] = self._seg_files.index_create(i, path, device, flags, load_flags)
                self._files[name] = (i, path)
        except Exception:
            self.close()
            raise
        return self

    def field_index(self, field: str) -> GpuIndex:
        return self._entries[field]

    @staticmethod
    def _wrap(knn_collector) -> JVectorKnnCollector:
        if isinstance(knn_collector, JVectorKnnCollector):
            return knn_collector
        # JVectorReader.java:133-144: re-wrap a plain collector with the defaults
        return JVectorKnnCollector(knn_collector, DEFAULT_QUERY_SIMILARITY_THRESHOLD, DEFAULT_QUERY_RERANK_FLOOR,
                                   DEFAULT_OVER_QUERY_FACTOR)

    def search(self, field: str, target, knn_collector, accept_docs=None) -> None:
        """KnnVectorsReader.search(String, float[], KnnCollector, AcceptDocs).  `accept_docs`: None or a
        bool mask / FixedBitSet words over Lucene docIds."""
        t = np.asarray(target)
        if t.dtype in (np.int8, np.uint8):
            raise NotImplementedError("Byte vector search is not supported yet with jVector")  # :241-245
        self.search_batch(field, t.reshape(1, -1), [knn_collector], accept_docs)

    def search_batch(self, field: str, targets, knn_collectors: Sequence, accept_docs=None) -> None:
        """New entry point: one call for many queries of one field (same k / parameters per batch)."""
        if self._closed:
            raise ValueError("reader is closed")
        ix = self._entries[field]
        cols = [self._wrap(c) for c in knn_collectors]
        c0 = cols[0]
        k = c0.k()
        bits = None
        if accept_docs is not None:
            a = np.asarray(accept_docs)
            bits = a if a.dtype == np.uint64 else make_accept_bits(a)
        if ix.n == 0:
            return
        res = ix.search(np.asarray(targets, dtype=np.float32), k, k * c0.over_query_factor, c0.threshold, c0.rerank_floor,
                        bits)
        KNNCounter.add_search_time(res.timing.get("total_ms", 0.0))                       # :192
        for i, col in enumerate(cols):
            for j in range(int(res.counts[i])):
                col.collect(int(res.docs[i, j]), float(res.scores[i, j]))            # :175-177
            visited, expanded, expanded_base, reranked = (int(x) for x in res.stats[i])
            KNNCounter.add(visited, reranked, expanded, expanded_base)                  # :189-193
            if visited + expanded > 0:
                col.inc_visited_count(visited + expanded)                               # :204-207

    def exact_search(self, field: str, target, k: int, accept_docs=None) -> List[ScoreDoc]:
        """Lucene exactSearch over FloatVectorValues.scorer(target) (JVectorVectorScorer.java:36-53)."""
        ix = self._entries[field]
        bits = None
        if accept_docs is not None:
            a = np.asarray(accept_docs)
            bits = a if a.dtype == np.uint64 else make_accept_bits(a)
        if ix.n == 0:
            return []
        docs, scores, counts = ix.exact_topk(np.asarray(target, dtype=np.float32).reshape(1, -1), k, bits)
        return [ScoreDoc(float(scores[0, j]), int(docs[0, j])) for j in range(int(counts[0]))]

    def get_float_vector_values(self, field: str) -> np.ndarray:
        if self._segment is None:        # opened from files: decode the inline vectors on demand
            i, path = self._files[field]
            return self._seg_files.load_field(i, path)["vectors"]
        return self._segment.fields[field].vectors

    def check_integrity(self) -> None:
        """JVectorReader.checkIntegrity, JVectorReader.java:87-99 (file-backed readers)."""
        from .segment_files import check_integrity
        for _, path in getattr(self, "_files", {}).values():
            check_integrity(path)

    def close(self):
        for ix in self._entries.values():
            ix.close()
        self._entries.clear()
        if getattr(self, "_seg_files", None) is not None:
            self._seg_files.close()
            self._seg_files = None
        self._closed = True


@dataclass
class JVectorKnnFloatVectorQuery:
    """JVectorKnnFloatVectorQuery.java:50-70 — per-leaf approximateSearch with the jVector parameters."""

    field: str
    target: Sequence[float]
    k: int
    over_query_factor: int = DEFAULT_OVER_QUERY_FACTOR
    threshold: float = DEFAULT_QUERY_SIMILARITY_THRESHOLD
    rerank_floor: float = DEFAULT_QUERY_RERANK_FLOOR

    def search(self, readers: Sequence[JVectorReader], accept_docs: Optional[Sequence] = None,
               doc_bases: Optional[Sequence[int]] = None) -> List[ScoreDoc]:
        """Per-leaf search + TopDocs.merge(k) across leaves (ties -> lower global docId)."""
        merged: List[ScoreDoc] = []
        for li, r in enumerate(readers):
            col = JVectorKnnCollector(TopKnnCollector(self.k), self.threshold, self.rerank_floor, self.over_query_factor)
            r.search(self.field, self.target, col, None if accept_docs is None else accept_docs[li])
            base = 0 if doc_bases is None else doc_bases[li]
            merged += [ScoreDoc(sd.score, sd.doc + base) for sd in col.top_docs()]
        merged.sort(key=lambda sd: (-sd.score, sd.doc))
        return merged[: self.k]
