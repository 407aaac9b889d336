"""ctypes front-end of the CPU oracle (oracle/jv_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

Parity status: "unpinned" for PQ codes / LUT / traversal order (no golden vectors exist in the
reference and jVector 4.0.0-rc.9 itself cannot run here); pinned for the analytic known-answer tests
the reference holds (tests/test_oracle_kat.py).  See the header of jv_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libjvoracle.so"
_SIMD_PATH = _HERE / "_build" / "libjvcpusimd.so"  # tuned CPU arm of the benchmark (jv_cpu_simd.c), never the checker

SIM_EUCLIDEAN, SIM_DOT, SIM_COSINE, SIM_MIP = 0, 1, 2, 3


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed Makefile (gcc only)."""
    stale = False
    for lib_path, src in ((_LIB_PATH, _HERE / "jv_oracle.c"), (_SIMD_PATH, _HERE / "jv_cpu_simd.c")):
        stale |= not lib_path.exists() or lib_path.stat().st_mtime < src.stat().st_mtime
    if force or stale:
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)
    return _LIB_PATH


class IndexDesc(C.Structure):
    """Mirror of jv_index_desc in include/jvgpu.h (the oracle consumes the same description)."""

    _fields_ = [
        ("struct_size", C.c_int32),
        ("similarity", C.c_int32),
        ("dim", C.c_int32),
        ("max_degree", C.c_int32),
        ("n", C.c_int64),
        ("entry_node", C.c_int32),
        ("max_doc", C.c_int32),
        ("adjacency", C.c_void_p),
        ("vectors", C.c_void_p),
        ("ord_to_doc", C.c_void_p),
        ("pq_m", C.c_int32),
        ("pq_k", C.c_int32),
        ("pq_codebooks", C.c_void_p),
        ("pq_global_centroid", C.c_void_p),
        ("pq_codes", C.c_void_p),
        ("device", C.c_int32),
        ("flags", C.c_uint32),
        ("nvq_m", C.c_int32),
        ("nvq_reserved", C.c_int32),
        ("nvq_bytes", C.c_void_p),
        ("nvq_params", C.c_void_p),
        ("nvq_global_mean", C.c_void_p),
    ]


class QueryStats(C.Structure):
    _fields_ = [("visited", C.c_int32), ("expanded", C.c_int32), ("expanded_base", C.c_int32), ("reranked", C.c_int32)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.jvo_index_create.restype = C.c_void_p
        _lib.jvo_index_create.argtypes = [C.POINTER(IndexDesc)]
        _lib.jvo_index_destroy.argtypes = [C.c_void_p]
        _lib.jvo_exact_score.restype = C.c_float
        _lib.jvo_default_num_subspaces.restype = C.c_int32
        _lib.jvo_num_threads.restype = C.c_int32
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def java_random_floats(seed: int, count: int) -> np.ndarray:
    """`new java.util.Random(seed)` then `count` x nextFloat() (TestUtils.java:108-124)."""
    out = np.empty(count, dtype=np.float32)
    lib().jvo_java_random_floats(C.c_int64(seed), C.c_int64(count), _p(out))
    return out


def java_random_vectors(n: int, dim: int, seed: int) -> np.ndarray:
    return java_random_floats(seed, n * dim).reshape(n, dim)


def exact_score(sim: int, q, x) -> float:
    q = _f32(q)
    x = _f32(x)
    return float(lib().jvo_exact_score(C.c_int32(sim), _p(q), _p(x), C.c_int32(q.shape[0])))


def default_num_subspaces(dim: int) -> int:
    return int(lib().jvo_default_num_subspaces(C.c_int32(dim)))


def pq_subspaces(dim: int, m: int):
    sizes = np.empty(m, dtype=np.int32)
    offs = np.empty(m, dtype=np.int32)
    lib().jvo_pq_subspaces(C.c_int32(dim), C.c_int32(m), _p(sizes), _p(offs))
    return sizes, offs


def pq_train(vectors, m: int, k: int, center: bool, iters: int = 6, seed: int = 0):
    v = _f32(vectors)
    n, dim = v.shape
    cb = np.empty(k * dim, dtype=np.float32)
    g = np.zeros(dim, dtype=np.float32) if center else None
    lib().jvo_pq_train(_p(v), C.c_int64(n), C.c_int32(dim), C.c_int32(m), C.c_int32(k), C.c_int32(int(center)),
                       C.c_int32(iters), C.c_uint64(seed), _p(cb), _p(g))
    return cb, g


def pq_encode(vectors, m: int, k: int, codebooks, gcent=None, threads: int = 0) -> np.ndarray:
    v = _f32(vectors)
    n, dim = v.shape
    cb = _f32(codebooks)
    g = _f32(gcent)
    out = np.empty((n, m), dtype=np.uint8)
    lib().jvo_pq_encode(_p(v), C.c_int64(n), C.c_int32(dim), C.c_int32(m), C.c_int32(k), _p(cb), _p(g), _p(out),
                        C.c_int32(threads))
    return out


def pq_decode(codes, dim: int, k: int, codebooks, gcent=None) -> np.ndarray:
    """ProductQuantization.decode (jVector 4.0.0-rc.9): the reconstruction of a code row is the concatenation of its centroids
    (+ the global centroid, one fp32 add per element).  Sub-vector split of SURVEY A.3: size = dim // M, the first dim % M
    subspaces one wider; codebooks laid out [m][c][size_m]."""
    c = np.ascontiguousarray(codes, dtype=np.uint8)
    n, m = c.shape
    cb = _f32(codebooks).ravel()
    out = np.empty((n, dim), dtype=np.float32)
    base, rem, off = dim // m, dim % m, 0
    for j in range(m):
        size = base + (1 if j < rem else 0)
        book = cb[k * off:k * off + k * size].reshape(k, size)
        out[:, off:off + size] = book[c[:, j]]
        off += size
    if gcent is not None:
        out = (out + _f32(gcent).reshape(1, dim)).astype(np.float32)
    return out


def graph_build_pq(codes, dim: int, k: int, codebooks, gcent, sim: int, max_degree: int = 32, beam_width: int = 100,
                   overflow: float = 1.2, alpha: float = 1.2, max_batch: int = 8192, frac: float = 0.02):
    """Fixture builder with PQ build scores (BuildScoreProvider.pqBuildScoreProvider, JVectorWriter.java:238-244, 1143-1151): the
    provider decodes node i and scores it against the codes of the other nodes, for the neighbour search, for the diversity
    pruning and for the approximate centroid alike — every build-time score is a score between reconstructions, so the builder
    runs on decode(codes).  (The summation order of the jar's ADC is not known here — un-vendored — hence the canonical order of
    the exact builder; parity unpinned like the rest of the builder)"""
    return graph_build(pq_decode(codes, dim, k, codebooks, gcent), sim, max_degree, beam_width, overflow, alpha, max_batch, frac)


def pq_lut(sim: int, dim: int, m: int, k: int, codebooks, gcent, queries) -> np.ndarray:
    q = _f32(queries)
    cb = _f32(codebooks)
    g = _f32(gcent)
    out = np.empty((q.shape[0], m, k), dtype=np.float32)
    lib().jvo_pq_lut(C.c_int32(sim), C.c_int32(dim), C.c_int32(m), C.c_int32(k), _p(cb), _p(g), _p(q),
                     C.c_int32(q.shape[0]), _p(out))
    return out


def nvq_encode(vectors, nvq_m: int = 2, growth_rate: float = 3.0, midpoint: float = 0.0):
    """FIXTURE NVQ encoder (see jv_oracle.c): returns (bytes [n, dim] u8, params [n, nvq_m, 4], global_mean [dim])."""
    v = _f32(vectors)
    n, dim = v.shape
    b = np.empty((n, dim), dtype=np.uint8)
    prm = np.empty((n, nvq_m, 4), dtype=np.float32)
    g = np.empty(dim, dtype=np.float32)
    lib().jvo_nvq_encode(_p(v), C.c_int64(n), C.c_int32(dim), C.c_int32(nvq_m), C.c_float(growth_rate), C.c_float(midpoint),
                         _p(b), _p(prm), _p(g))
    return b, prm, g


def nvq_dequantize(nvq_bytes, nvq_params, global_mean) -> np.ndarray:
    """nvqDequantize, JVectorIndexQuantization.java:316-341."""
    b = np.ascontiguousarray(nvq_bytes, dtype=np.uint8)
    prm = _f32(nvq_params)
    g = _f32(global_mean)
    n, dim = b.shape
    out = np.empty((n, dim), dtype=np.float32)
    lib().jvo_nvq_dequantize(C.c_int64(n), C.c_int32(dim), C.c_int32(prm.shape[1]), _p(b), _p(prm), _p(g), _p(out))
    return out


def graph_build(vectors, sim: int, max_degree: int = 32, beam_width: int = 100, overflow: float = 1.2,
                alpha: float = 1.2, max_batch: int = 8192, frac: float = 0.02):
    """Fixture: batched-insert Vamana with exact build scores.  Returns (adjacency[n,R], entry)."""
    v = _f32(vectors)
    n, dim = v.shape
    adj = np.empty((n, max_degree), dtype=np.int32)
    entry = C.c_int32(0)
    lib().jvo_graph_build(_p(v), C.c_int64(n), C.c_int32(dim), C.c_int32(sim), C.c_int32(max_degree),
                          C.c_int32(beam_width), C.c_float(overflow), C.c_float(alpha), C.c_int32(max_batch),
                          C.c_float(frac), _p(adj), C.byref(entry))
    return adj, int(entry.value)


def graph_extend(vectors, seed_adjacency, seed_entry: int, sim: int, beam_width: int = 100, overflow: float = 1.2,
                 alpha: float = 1.2, max_batch: int = 8192, frac: float = 0.02):
    """Fixture: leading-segment merge, insert-only (JVectorWriter.java:1166-1341): the first len(seed_adjacency) ordinals
    keep their graph, the remaining vectors are added with the same batched-insert schedule.  Returns adjacency[n, R]."""
    v = _f32(vectors)
    n, dim = v.shape
    seed = np.ascontiguousarray(seed_adjacency, dtype=np.int32)
    n0, r = seed.shape
    adj = np.empty((n, r), dtype=np.int32)
    st = lib().jvo_graph_extend(_p(v), C.c_int64(n), C.c_int64(n0), _p(seed), C.c_int32(seed_entry), C.c_int32(dim), C.c_int32(sim),
                                C.c_int32(r), C.c_int32(beam_width), C.c_float(overflow), C.c_float(alpha), C.c_int32(max_batch),
                                C.c_float(frac), _p(adj))
    if st != 0:
        raise ValueError("bad seed graph")
    return adj


def graph_remove_deleted(vectors, adjacency, entry: int, deleted, sim: int, alpha: float = 1.2):
    """Fixture: delete consolidation (markNodeDeleted + cleanup, JVectorWriter.java:1318-1327; FreshDiskANN 4.2).  Same ordinal
    space in and out: rows of deleted nodes come back empty.  Returns (adjacency[n, R], entry) — entry = -1 when nothing is live."""
    v = _f32(vectors)
    n, dim = v.shape
    a = np.ascontiguousarray(adjacency, dtype=np.int32)
    d = np.ascontiguousarray(np.asarray(deleted, dtype=bool).astype(np.uint8))
    out = np.empty_like(a)
    e = C.c_int32(0)
    lib().jvo_graph_remove_deleted(_p(v), C.c_int64(n), C.c_int32(dim), C.c_int32(sim), C.c_int32(a.shape[1]), C.c_float(alpha),
                                   _p(a), _p(d), C.c_int32(entry), _p(out), C.byref(e))
    return out, int(e.value)


class OracleIndex:
    """One field of one segment, as decoded arrays (what FieldEntry holds, JVectorReader.java:284-337)."""

    def __init__(self, similarity: int, vectors, adjacency, entry_node: int, ord_to_doc=None, max_doc=None,
                 pq_m: int = 0, pq_k: int = 0, pq_codebooks=None, pq_global_centroid=None, pq_codes=None,
                 adc_order: int = 0, nvq_m: int = 0, nvq_bytes=None, nvq_params=None, nvq_global_mean=None):
        self.vectors = _f32(vectors)
        self.n, self.dim = self.vectors.shape
        self.adjacency = np.ascontiguousarray(adjacency, dtype=np.int32)
        self.ord_to_doc = None if ord_to_doc is None else np.ascontiguousarray(ord_to_doc, dtype=np.int32)
        self.max_doc = int(max_doc) if max_doc is not None else (
            self.n if self.ord_to_doc is None else int(self.ord_to_doc.max(initial=-1)) + 1)
        self.codebooks = _f32(pq_codebooks)
        self.gcent = _f32(pq_global_centroid)
        self.codes = None if pq_codes is None else np.ascontiguousarray(pq_codes, dtype=np.uint8)
        self.similarity = similarity
        d = IndexDesc()
        d.struct_size = C.sizeof(IndexDesc)
        d.similarity = similarity
        d.dim = self.dim
        d.max_degree = self.adjacency.shape[1] if self.adjacency.ndim == 2 else 0
        d.n = self.n
        d.entry_node = entry_node
        d.max_doc = self.max_doc
        d.adjacency = self.adjacency.ctypes.data
        d.vectors = self.vectors.ctypes.data
        d.ord_to_doc = None if self.ord_to_doc is None else self.ord_to_doc.ctypes.data
        d.pq_m = pq_m if self.codes is not None else 0
        d.pq_k = pq_k
        d.pq_codebooks = None if self.codebooks is None else self.codebooks.ctypes.data
        d.pq_global_centroid = None if self.gcent is None else self.gcent.ctypes.data
        d.pq_codes = None if self.codes is None else self.codes.ctypes.data
        self.nvq_bytes = None if nvq_bytes is None else np.ascontiguousarray(nvq_bytes, dtype=np.uint8)
        self.nvq_params = _f32(nvq_params)
        self.nvq_gmean = _f32(nvq_global_mean)
        d.nvq_m = nvq_m if self.nvq_bytes is not None else 0
        d.nvq_bytes = None if self.nvq_bytes is None else self.nvq_bytes.ctypes.data
        d.nvq_params = None if self.nvq_params is None else self.nvq_params.ctypes.data
        d.nvq_global_mean = None if self.nvq_gmean is None else self.nvq_gmean.ctypes.data
        self.desc = d
        self._h = lib().jvo_index_create(C.byref(d))
        if adc_order:
            lib().jvo_index_set_adc_order(C.c_void_p(self._h), C.c_int32(adc_order))

    def close(self):
        if self._h:
            lib().jvo_index_destroy(C.c_void_p(self._h))
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _bits(self, accept_bits, nq):
        if accept_bits is None:
            return None, 0
        b = np.ascontiguousarray(accept_bits, dtype=np.uint64)
        stride = 0 if b.ndim == 1 else b.shape[1]
        return b, stride

    def search(self, queries, k: int, rerank_k: int, threshold: float = 0.0, rerank_floor: float = 0.0,
               accept_bits=None, threads: int = 0):
        q = _f32(np.atleast_2d(queries))
        nq = q.shape[0]
        docs = np.empty((nq, k), dtype=np.int32)
        scores = np.empty((nq, k), dtype=np.float32)
        counts = np.empty(nq, dtype=np.int32)
        stats = (QueryStats * nq)()
        b, stride = self._bits(accept_bits, nq)
        rc = lib().jvo_search_batch(C.c_void_p(self._h), _p(q), C.c_int32(nq), C.c_int32(k), C.c_int32(rerank_k),
                                    C.c_float(threshold), C.c_float(rerank_floor), _p(b), C.c_int64(stride), _p(docs),
                                    _p(scores), _p(counts), stats, C.c_int32(threads))
        if rc != 0:
            raise ValueError("rerankK must be >= topK")
        st = np.frombuffer(stats, dtype=np.int32).reshape(nq, 4).copy()
        return docs, scores, counts, st

    def search_approx(self, queries, rerank_k: int, threshold: float = 0.0, accept_bits=None):
        q = _f32(np.atleast_2d(queries))
        nq = q.shape[0]
        nodes = np.empty((nq, rerank_k), dtype=np.int32)
        scores = np.empty((nq, rerank_k), dtype=np.float32)
        counts = np.empty(nq, dtype=np.int32)
        stats = (QueryStats * nq)()
        b, stride = self._bits(accept_bits, nq)
        lib().jvo_search_approx(C.c_void_p(self._h), _p(q), C.c_int32(nq), C.c_int32(rerank_k), C.c_float(threshold),
                                _p(b), C.c_int64(stride), _p(nodes), _p(scores), _p(counts), stats)
        st = np.frombuffer(stats, dtype=np.int32).reshape(nq, 4).copy()
        return nodes, scores, counts, st

    def adc_scores(self, queries, nodes):
        q = _f32(np.atleast_2d(queries))
        nd = np.ascontiguousarray(np.atleast_2d(nodes), dtype=np.int32)
        out = np.empty(nd.shape, dtype=np.float32)
        lib().jvo_pq_adc_scores(C.c_void_p(self._h), _p(q), C.c_int32(q.shape[0]), _p(nd), C.c_int32(nd.shape[1]), _p(out))
        return out

    def lut_q8(self, queries):
        """8-bit ADC tables (adc_order = -8): (q8 [nq, M, K] u8, params [nq, 2] = (delta, base))."""
        q = _f32(np.atleast_2d(queries))
        m, k = self.desc.pq_m, self.desc.pq_k
        q8 = np.empty((q.shape[0], m, k), dtype=np.uint8)
        params = np.empty((q.shape[0], 2), dtype=np.float32)
        rc = lib().jvo_pq_lut_q8(C.c_void_p(self._h), _p(q), C.c_int32(q.shape[0]), _p(q8), _p(params))
        if rc != 0:
            raise ValueError("lut_q8 needs a PQ index created with adc_order=-8")
        return q8, params

    def exact_topk(self, queries, k: int, accept_bits=None, threads: int = 0):
        q = _f32(np.atleast_2d(queries))
        nq = q.shape[0]
        docs = np.empty((nq, k), dtype=np.int32)
        scores = np.empty((nq, k), dtype=np.float32)
        counts = np.empty(nq, dtype=np.int32)
        b, stride = self._bits(accept_bits, nq)
        lib().jvo_exact_topk(C.c_void_p(self._h), _p(q), C.c_int32(nq), C.c_int32(k), _p(b), C.c_int64(stride),
                             _p(docs), _p(scores), _p(counts), C.c_int32(threads))
        return docs, scores, counts


def merge_topk(docs, scores, k: int):
    """docs/scores [g, nq, k] -> merged top-k (tie -> lower doc)."""
    d = np.ascontiguousarray(docs, dtype=np.int32)
    s = _f32(scores)
    g, nq, kk = d.shape
    assert kk == k
    od = np.empty((nq, k), dtype=np.int32)
    os_ = np.empty((nq, k), dtype=np.float32)
    oc = np.empty(nq, dtype=np.int32)
    lib().jvo_merge_topk(C.c_int32(g), C.c_int32(nq), C.c_int32(k), _p(d), _p(s), _p(od), _p(os_), _p(oc))
    return od, os_, oc


_simd = None


def simd_lib() -> C.CDLL:
    global _simd
    if _simd is None:
        build()
        _simd = C.CDLL(str(_SIMD_PATH))
        _simd.jvs_create.restype = C.c_void_p
        _simd.jvs_create.argtypes = [C.POINTER(IndexDesc)]
        _simd.jvs_destroy.argtypes = [C.c_void_p]
        _simd.jvs_isa.restype = C.c_int32
        _simd.jvs_num_procs.restype = C.c_int32
    return _simd


class SimdIndex:
    """The tuned CPU arm (jv_cpu_simd.c): same algorithm as OracleIndex.search for unfiltered queries with the default collector
    parameters, AVX-512 / AVX2 gathers and FMAs, free summation order.  Bench baseline; gated against the checker by recall."""

    def __init__(self, oracle_index: "OracleIndex"):
        self._keep = oracle_index  # owns the arrays the description points to
        self._h = simd_lib().jvs_create(C.byref(oracle_index.desc))
        if not self._h:
            raise ValueError("the tuned CPU arm needs K = 256 and uniform sub-vectors")
        self.isa = "avx512" if simd_lib().jvs_isa() == 512 else "avx2"

    def close(self):
        if self._h:
            simd_lib().jvs_destroy(C.c_void_p(self._h))
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def search(self, queries, k: int, rerank_k: int, threads: int = 0):
        q = _f32(np.atleast_2d(queries))
        nq = q.shape[0]
        docs = np.empty((nq, k), dtype=np.int32)
        scores = np.empty((nq, k), dtype=np.float32)
        counts = np.empty(nq, dtype=np.int32)
        stats = (QueryStats * nq)()
        rc = simd_lib().jvs_search_batch(C.c_void_p(self._h), _p(q), C.c_int32(nq), C.c_int32(k), C.c_int32(rerank_k), _p(docs), _p(scores),
                                         _p(counts), stats, C.c_int32(threads))
        if rc != 0:
            raise ValueError("bad arguments")
        return docs, scores, counts, np.frombuffer(stats, dtype=np.int32).reshape(nq, 4).copy()


def host_cores() -> int:
    """Cores this process may run on (torchrun exports OMP_NUM_THREADS=1: never trust omp_get_max_threads for a baseline)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def num_threads() -> int:
    return int(lib().jvo_num_threads())


def make_accept_bits(accept_mask) -> np.ndarray:
    """bool[maxDoc] -> Lucene FixedBitSet words (bit d = word d>>6, bit d&63)."""
    m = np.asarray(accept_mask, dtype=bool)
    nwords = (m.shape[0] + 63) // 64
    padded = np.zeros(nwords * 64, dtype=bool)
    padded[: m.shape[0]] = m
    return np.packbits(padded.reshape(nwords, 64), axis=1, bitorder="little").view(np.uint64).reshape(nwords)
