"""Device-resident index handle: thin, typed wrapper over the C-ABI (include/jvgpu.h).

One `GpuIndex` = one field of one segment, i.e. what `JVectorReader.FieldEntry` holds after
JVectorReader.java:284-337 (graph + inline vectors + PQVectors + GraphNodeIdToDocMap), copied to HBM.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import native as N


def _ptr(a) -> Optional[int]:
    return None if a is None else a.ctypes.data


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def make_accept_bits(accept_mask) -> np.ndarray:
    """bool[maxDoc] (or [nq, maxDoc]) -> Lucene FixedBitSet words: bit d = word d>>6, bit d&63."""
    m = np.asarray(accept_mask, dtype=bool)
    if m.ndim == 2:
        return np.stack([make_accept_bits(r) for r in m])
    nwords = (m.shape[0] + 63) // 64
    padded = np.zeros(nwords * 64, dtype=bool)
    padded[: m.shape[0]] = m
    return np.packbits(padded.reshape(nwords, 64), axis=1, bitorder="little").view(np.uint64).reshape(nwords)


@dataclass
class SearchResult:
    docs: np.ndarray     # [nq, k] int32, -1 padded, sorted by (score desc, doc asc)
    scores: np.ndarray   # [nq, k] float32 (jVector-scaled)
    counts: np.ndarray   # [nq]
    stats: np.ndarray    # [nq, 4] visited, expanded, expanded_base, reranked (JVectorReader.java:183-187)
    timing: dict


def pq_encode(vectors, m: int, k: int, codebooks, global_centroid=None, device: int = 0, return_ms: bool = False):
    """PQVectors.encodeAndBuild on the GPU (K6)."""
    v = _f32(vectors)
    n, dim = v.shape
    cb = _f32(codebooks)
    g = _f32(global_centroid)
    out = np.empty((n, m), dtype=np.uint8)
    ms = C.c_float(0)
    N.check(N.load().jv_pq_encode(device, _ptr(v), n, dim, m, k, _ptr(cb), _ptr(g), _ptr(out), C.addressof(ms)))
    return (out, float(ms.value)) if return_ms else out


def pq_train(vectors, m: int, k: int, center: bool, iters: int = 6, seed: int = 0, device: int = 0):
    """ProductQuantization.compute on the GPU.  Returns (codebooks, global_centroid|None)."""
    v = _f32(vectors)
    n, dim = v.shape
    cb = np.empty(k * dim, dtype=np.float32)
    g = np.zeros(dim, dtype=np.float32) if center else None
    N.check(N.load().jv_pq_train(device, _ptr(v), n, dim, m, k, int(center), iters, seed, _ptr(cb), _ptr(g)))
    return cb, g


def graph_build(vectors, similarity: int, max_degree: int = 32, beam_width: int = 100, neighbor_overflow: float = 1.2,
                alpha: float = 1.2, device: int = 0):
    """GraphIndexBuilder on the GPU.  Returns (adjacency[n, R] int32, entry_node)."""
    v = _f32(vectors)
    n, dim = v.shape
    adj = np.empty((n, max_degree), dtype=np.int32)
    entry = C.c_int32(0)
    N.check(N.load().jv_graph_build(device, _ptr(v), n, dim, similarity, max_degree, beam_width, neighbor_overflow, alpha,
                                    _ptr(adj), C.addressof(entry)))
    return adj, int(entry.value)


def pq_decode(codes, dim: int, k: int, codebooks, global_centroid=None, device: int = 0):
    """ProductQuantization.decode on the GPU: reconstructions [n, dim] float32 of the code rows."""
    c = np.ascontiguousarray(codes, dtype=np.uint8)
    n, m = c.shape
    cb = _f32(codebooks)
    g = _f32(global_centroid)
    out = np.empty((n, dim), dtype=np.float32)
    N.check(N.load().jv_pq_decode(device, _ptr(c), n, dim, m, k, _ptr(cb), _ptr(g), _ptr(out)))
    return out


def graph_build_pq(codes, dim: int, k: int, codebooks, global_centroid, similarity: int, max_degree: int = 32, beam_width: int = 100,
                   neighbor_overflow: float = 1.2, alpha: float = 1.2, device: int = 0):
    """GraphIndexBuilder with BuildScoreProvider.pqBuildScoreProvider on the GPU (JVectorWriter.java:238-244, 1143-1151): build
    scores between PQ reconstructions.  Returns (adjacency[n, R] int32, entry_node)."""
    c = np.ascontiguousarray(codes, dtype=np.uint8)
    n, m = c.shape
    cb = _f32(codebooks)
    g = _f32(global_centroid)
    adj = np.empty((n, max_degree), dtype=np.int32)
    entry = C.c_int32(0)
    N.check(N.load().jv_graph_build_pq(device, _ptr(c), n, dim, similarity, m, k, _ptr(cb), _ptr(g), max_degree, beam_width,
                                       neighbor_overflow, alpha, _ptr(adj), C.addressof(entry)))
    return adj, int(entry.value)


def graph_extend(vectors, seed_adjacency, seed_entry: int, similarity: int, beam_width: int = 100, neighbor_overflow: float = 1.2,
                 alpha: float = 1.2, device: int = 0):
    """Leading-segment merge, insert-only (JVectorWriter.java:1166-1341): ordinals [0, len(seed_adjacency)) keep their graph
    and entry node, the remaining vectors are inserted.  Returns adjacency[n, R] int32."""
    v = _f32(vectors)
    n, dim = v.shape
    seed = np.ascontiguousarray(seed_adjacency, dtype=np.int32)
    n0, r = seed.shape
    adj = np.empty((n, r), dtype=np.int32)
    N.check(N.load().jv_graph_extend(device, _ptr(v), n, n0, _ptr(seed), seed_entry, dim, similarity, r, beam_width, neighbor_overflow,
                                     alpha, _ptr(adj)))
    return adj


def graph_remove_deleted(vectors, adjacency, entry_node: int, deleted, similarity: int, alpha: float = 1.2, device: int = 0):
    """markNodeDeleted + cleanup of a merge (JVectorWriter.java:1318-1327): FreshDiskANN-style delete consolidation on the GPU.
    Same ordinal space in and out.  Returns (adjacency[n, R], entry_node)."""
    v = _f32(vectors)
    n, dim = v.shape
    a = np.ascontiguousarray(adjacency, dtype=np.int32)
    d = np.ascontiguousarray(np.asarray(deleted, dtype=bool).astype(np.uint8))
    out = np.empty_like(a)
    e = C.c_int32(0)
    N.check(N.load().jv_graph_remove_deleted(device, _ptr(v), n, dim, similarity, a.shape[1], alpha, _ptr(a), _ptr(d), entry_node,
                                             _ptr(out), C.addressof(e)))
    return out, int(e.value)


def merge_topk(docs, scores, k: int, device: int = 0):
    """[g, nq, k] per-shard lists -> merged [nq, k] (K7)."""
    d = np.ascontiguousarray(docs, dtype=np.int32)
    s = _f32(scores)
    g, nq, kk = d.shape
    if kk != k:
        raise ValueError("last dimension must equal k")
    od = np.empty((nq, k), dtype=np.int32)
    os_ = np.empty((nq, k), dtype=np.float32)
    oc = np.empty(nq, dtype=np.int32)
    N.check(N.load().jv_merge_topk(device, g, nq, k, _ptr(d), _ptr(s), _ptr(od), _ptr(os_), _ptr(oc)))
    return od, os_, oc


class GpuIndex:
    def __init__(self, similarity: int, vectors, adjacency, entry_node: int, ord_to_doc=None, max_doc: Optional[int] = None,
                 pq_m: int = 0, pq_k: int = 0, pq_codebooks=None, pq_global_centroid=None, pq_codes=None, device: int = 0,
                 flags: int = 0, nvq_m: int = 0, nvq_bytes=None, nvq_params=None, nvq_global_mean=None):
        """`vectors` may be None for an nvq+pq segment (NVQ-inline vectors + auxiliary PQ codes): the reranker then scores
        the dequantised NVQ vectors (JVectorReader.java:352-358) and brute force is unsupported, as in the reference."""
        lib = N.load()
        nb = None if nvq_bytes is None else np.ascontiguousarray(nvq_bytes, dtype=np.uint8)
        v = _f32(vectors)
        if v is None:
            if nb is None:
                raise ValueError("vectors or nvq_bytes are required")
            self.n, self.dim = nb.shape
        else:
            if v.ndim != 2:
                raise ValueError("vectors must be [n, dim]")
            self.n, self.dim = v.shape
        adj = np.ascontiguousarray(adjacency, dtype=np.int32).reshape(self.n, -1) if self.n else np.zeros((0, 1), np.int32)
        o2d = None if ord_to_doc is None else np.ascontiguousarray(ord_to_doc, dtype=np.int32)
        self.max_doc = int(max_doc) if max_doc is not None else (
            self.n if o2d is None else int(o2d.max(initial=-1)) + 1)
        cb = _f32(pq_codebooks)
        g = _f32(pq_global_centroid)
        codes = None if pq_codes is None else np.ascontiguousarray(pq_codes, dtype=np.uint8)
        self.similarity = similarity
        self.device = device
        self.has_pq = codes is not None
        self.pq_m, self.pq_k = (pq_m, pq_k) if self.has_pq else (0, 0)
        self.max_degree = adj.shape[1]
        d = N.IndexDesc()
        d.struct_size = C.sizeof(N.IndexDesc)
        d.similarity, d.dim, d.max_degree, d.n = similarity, self.dim, self.max_degree, self.n
        d.entry_node, d.max_doc = entry_node, self.max_doc
        d.adjacency, d.vectors, d.ord_to_doc = _ptr(adj), _ptr(v), _ptr(o2d)
        d.pq_m, d.pq_k = self.pq_m, self.pq_k
        d.pq_codebooks, d.pq_global_centroid, d.pq_codes = _ptr(cb), _ptr(g), _ptr(codes)
        d.device, d.flags = device, flags
        nprm, ngm = _f32(nvq_params), _f32(nvq_global_mean)
        d.nvq_m = nvq_m if nb is not None else 0
        d.nvq_bytes, d.nvq_params, d.nvq_global_mean = _ptr(nb), _ptr(nprm), _ptr(ngm)
        h = C.c_void_p()
        N.check(lib.jv_index_create(C.addressof(d), C.addressof(h)))
        self._h = h

    @classmethod
    def from_handle(cls, handle, similarity: int, n: int, dim: int, max_doc: int, device: int = 0, has_pq: bool = False,
                    pq_m: int = 0, pq_k: int = 0, max_degree: int = 0) -> "GpuIndex":
        """Wrap a jv_index* created natively (jv_segment_index_create); the wrapper owns the handle."""
        self = object.__new__(cls)
        self.n, self.dim, self.max_doc = int(n), int(dim), int(max_doc)
        self.similarity, self.device = int(similarity), int(device)
        self.has_pq, self.pq_m, self.pq_k, self.max_degree = bool(has_pq), int(pq_m), int(pq_k), int(max_degree)
        self._h = handle
        return self

    # -- lifetime (FieldEntry.close, JVectorReader.java:367-378)
    def close(self):
        if getattr(self, "_h", None):
            N.load().jv_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self):
        if not self._h:
            raise ValueError("index is closed")
        return self._h

    def device_bytes(self) -> int:
        b = C.c_int64(0)
        N.check(N.load().jv_index_device_bytes(self.handle, C.addressof(b)))
        return int(b.value)

    def visited_overflows(self) -> int:
        """Queries (since creation) whose shared-memory visited set filled up."""
        b = C.c_int64(0)
        N.check(N.load().jv_index_debug_counter(self.handle, 0, C.addressof(b)))
        return int(b.value)

    def refresh_knobs(self):
        """Re-read the JVGPU_* diagnostic environment knobs (they are cached by the library)."""
        b = C.c_int64(0)
        N.check(N.load().jv_index_debug_counter(self.handle, 200, C.addressof(b)))

    def exact_tc_counters(self):
        """(brute-force batches answered by the tensor-core path, batches handed back to the fp32 kernel after an overflow)."""
        b = C.c_int64(0)
        out = []
        for which in (4, 5):
            N.check(N.load().jv_index_debug_counter(self.handle, which, C.addressof(b)))
            out.append(int(b.value))
        return tuple(out)

    PHASES = ("setup", "table_build", "select", "neighbour_rows", "scoring", "merge", "emit", "steps",
              "sub_code_words", "sub_lookups", "sub_offers", "sub_11", "sub_12", "sub_13", "sub_14", "sub_15")

    def phase_cycles(self, reset: bool = False) -> dict:
        """Per-phase SM cycles of the fast traversal kernel (thread 0 of each CTA, summed) + step count."""
        out = {}
        b = C.c_int64(0)
        for i, name in enumerate(self.PHASES):
            N.check(N.load().jv_index_debug_counter(self.handle, 8 + i, C.addressof(b)))
            out[name] = int(b.value)
        if reset:
            N.check(N.load().jv_index_debug_counter(self.handle, 100, C.addressof(b)))
        return out

    def _params(self, k, rerank_k, threshold, rerank_floor, accept_ptr, stride, expand_width=0):
        p = N.SearchParams()
        p.struct_size = C.sizeof(N.SearchParams)
        p.k, p.rerank_k, p.threshold, p.rerank_floor = k, rerank_k, threshold, rerank_floor
        p.expand_width = expand_width
        p.accept_bits, p.accept_stride_words = accept_ptr, stride
        return p

    # -- K1+K2(+K4)+K3
    def search(self, queries, k: int, rerank_k: int, threshold: float = 0.0, rerank_floor: float = 0.0,
               accept_bits=None, expand_width: int = 0) -> SearchResult:
        q = _f32(np.atleast_2d(queries))
        if q.shape[1] != self.dim:
            raise ValueError(f"query dimension {q.shape[1]} != index dimension {self.dim}")
        nq = q.shape[0]
        docs = np.full((nq, k), -1, dtype=np.int32)
        scores = np.zeros((nq, k), dtype=np.float32)
        counts = np.zeros(nq, dtype=np.int32)
        stats = np.zeros((nq, 4), dtype=np.int32)
        bits, stride = None, 0
        if accept_bits is not None:
            bits = np.ascontiguousarray(accept_bits, dtype=np.uint64)
            stride = 0 if bits.ndim == 1 else bits.shape[1]
        p = self._params(k, rerank_k, threshold, rerank_floor, _ptr(bits), stride, expand_width)
        t = N.BatchTiming()
        N.check(N.load().jv_search_batch(self.handle, _ptr(q), nq, C.addressof(p), _ptr(docs), _ptr(scores), _ptr(counts),
                                         _ptr(stats), C.addressof(t)))
        timing = {f: getattr(t, f) for f, _ in N.BatchTiming._fields_ if f != "reserved"}
        return SearchResult(docs, scores, counts, stats, timing)

    def search_dev(self, d_queries_ptr: int, nq: int, k: int, rerank_k: int, d_out_doc: int, d_out_score: int,
                   d_out_count: int, d_stats: int = None, threshold: float = 0.0, rerank_floor: float = 0.0,
                   d_accept_bits: int = None, accept_stride_words: int = 0, expand_width: int = 0) -> dict:
        """Device-pointer variant (inputs resident in HBM): returns the device-side timing."""
        p = self._params(k, rerank_k, threshold, rerank_floor, d_accept_bits, accept_stride_words, expand_width)
        t = N.BatchTiming()
        N.check(N.load().jv_search_batch_dev(self.handle, d_queries_ptr, nq, C.addressof(p), d_out_doc, d_out_score,
                                             d_out_count, d_stats, C.addressof(t)))
        return {f: getattr(t, f) for f, _ in N.BatchTiming._fields_ if f != "reserved"}

    # -- K5
    def exact_topk(self, queries, k: int, accept_bits=None):
        q = _f32(np.atleast_2d(queries))
        if q.shape[1] != self.dim:
            raise ValueError(f"query dimension {q.shape[1]} != index dimension {self.dim}")
        nq = q.shape[0]
        docs = np.full((nq, k), -1, dtype=np.int32)
        scores = np.zeros((nq, k), dtype=np.float32)
        counts = np.zeros(nq, dtype=np.int32)
        bits, stride = None, 0
        if accept_bits is not None:
            bits = np.ascontiguousarray(accept_bits, dtype=np.uint64)
            stride = 0 if bits.ndim == 1 else bits.shape[1]
        N.check(N.load().jv_exact_topk(self.handle, _ptr(q), nq, k, _ptr(bits), stride, _ptr(docs), _ptr(scores), _ptr(counts)))
        return docs, scores, counts

    def exact_topk_dev(self, d_queries_ptr: int, nq: int, k: int, d_out_doc: int, d_out_score: int, d_out_count: int,
                       d_accept_bits: int = None, accept_stride_words: int = 0):
        N.check(N.load().jv_exact_topk_dev(self.handle, d_queries_ptr, nq, k, d_accept_bits, accept_stride_words, d_out_doc,
                                           d_out_score, d_out_count))

    # -- K1 / a4 test hooks
    def pq_lut(self, queries) -> np.ndarray:
        q = _f32(np.atleast_2d(queries))
        out = np.empty((q.shape[0], self.pq_m, self.pq_k), dtype=np.float32)
        N.check(N.load().jv_pq_lut(self.handle, _ptr(q), q.shape[0], _ptr(out)))
        return out

    def pq_lut_q8(self, queries):
        """8-bit ADC tables of an index created with FLAG_LUT_U8: (q8 [nq, M, 256] u8, params [nq, 2] = (delta, base))."""
        q = _f32(np.atleast_2d(queries))
        q8 = np.empty((q.shape[0], self.pq_m, 256), dtype=np.uint8)
        params = np.empty((q.shape[0], 2), dtype=np.float32)
        N.check(N.load().jv_pq_lut_q8(self.handle, _ptr(q), q.shape[0], _ptr(q8), _ptr(params)))
        return q8, params

    def adc_scores(self, queries, nodes) -> np.ndarray:
        q = _f32(np.atleast_2d(queries))
        nd = np.ascontiguousarray(np.atleast_2d(nodes), dtype=np.int32)
        out = np.empty(nd.shape, dtype=np.float32)
        N.check(N.load().jv_pq_adc_scores(self.handle, _ptr(q), q.shape[0], _ptr(nd), nd.shape[1], _ptr(out)))
        return out
