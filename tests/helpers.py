"""Shared fixtures for the parity tests: seeded data, oracle-built segments, recall."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from oracle import oracle as O


@dataclass
class Fixture:
    sim: int
    base: np.ndarray
    queries: np.ndarray
    adjacency: np.ndarray
    entry: int
    ord_to_doc: Optional[np.ndarray] = None
    max_doc: Optional[int] = None
    pq_m: int = 0
    pq_k: int = 0
    codebooks: Optional[np.ndarray] = None
    gcent: Optional[np.ndarray] = None
    codes: Optional[np.ndarray] = None

    def oracle_index(self, adc_order: int = 0) -> O.OracleIndex:
        return O.OracleIndex(self.sim, self.base, self.adjacency, self.entry, self.ord_to_doc, self.max_doc, self.pq_m,
                             self.pq_k, self.codebooks, self.gcent, self.codes, adc_order=adc_order)

    def gpu_index(self, jv, flags: int = 0):
        return jv.GpuIndex(self.sim, self.base, self.adjacency, self.entry, ord_to_doc=self.ord_to_doc, max_doc=self.max_doc,
                           pq_m=self.pq_m, pq_k=self.pq_k, pq_codebooks=self.codebooks, pq_global_centroid=self.gcent,
                           pq_codes=self.codes, flags=flags)


def clustered(n: int, dim: int, nq: int, seed: int, clusters: int = 32, spread: float = 0.35, normalize: bool = False):
    """Gaussian-mixture data ("Cohere/Deep/OpenAI-shaped" at test scale)."""
    rng = np.random.default_rng(seed)
    cent = rng.standard_normal((clusters, dim)).astype(np.float32)
    base = cent[rng.integers(0, clusters, n)] + spread * rng.standard_normal((n, dim)).astype(np.float32)
    qs = cent[rng.integers(0, clusters, nq)] + spread * rng.standard_normal((nq, dim)).astype(np.float32)
    base, qs = base.astype(np.float32), qs.astype(np.float32)
    if normalize:
        base /= np.linalg.norm(base, axis=1, keepdims=True)
        qs /= np.linalg.norm(qs, axis=1, keepdims=True)
    return base, qs


def embedded(n: int, dim: int, nq: int, seed: int, latent: int = 24, clusters: int = 32):
    """Embedding-shaped data (bench.py gen_data at test scale): a Gaussian mixture in a `latent`-d space mapped into
    `dim` dimensions by a fixed random map plus small isotropic noise, L2-normalised."""
    rng = np.random.default_rng(seed)
    W = (rng.standard_normal((latent, dim)) / np.sqrt(latent)).astype(np.float32)
    cent = rng.standard_normal((clusters, latent)).astype(np.float32)

    def draw(m):
        z = cent[rng.integers(0, clusters, m)] + 0.6 * rng.standard_normal((m, latent)).astype(np.float32)
        x = z @ W + 0.02 * rng.standard_normal((m, dim)).astype(np.float32)
        return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)

    return draw(n), draw(nq)


def make_fixture(sim: int, base: np.ndarray, queries: np.ndarray, max_degree: int = 16, beam_width: int = 100, pq_m: int = 0,
                 pq_k: int = 256, ord_to_doc=None, max_doc=None, seed: int = 7) -> Fixture:
    """Segment built entirely by the oracle's fixture builders (CPU): graph + optional PQ."""
    adj, entry = O.graph_build(base, sim, max_degree, beam_width)
    fx = Fixture(sim, np.ascontiguousarray(base, np.float32), np.ascontiguousarray(queries, np.float32), adj, entry,
                 ord_to_doc, max_doc)
    if pq_m:
        k = min(pq_k, base.shape[0])
        cb, g = O.pq_train(base, pq_m, k, center=(sim == O.SIM_EUCLIDEAN), iters=6, seed=seed)
        fx.pq_m, fx.pq_k, fx.codebooks, fx.gcent = pq_m, k, cb, g
        fx.codes = O.pq_encode(base, pq_m, k, cb, g)
    return fx


def recall(found: np.ndarray, truth: np.ndarray) -> float:
    """|result ∩ GT| / k averaged over queries (JVectorWriterMergeTests.java:199-212)."""
    k = truth.shape[1]
    hits = 0
    for f, t in zip(found, truth):
        hits += len(set(int(x) for x in f if x >= 0) & set(int(x) for x in t if x >= 0))
    return hits / (k * len(truth))


def lucene_score(sim: int, q, x) -> float:
    """Lucene VectorSimilarityFunction.compare in float64 (CommonTestUtils.java:84-93)."""
    q = np.asarray(q, np.float64)
    x = np.asarray(x, np.float64)
    if sim == O.SIM_EUCLIDEAN:
        return 1.0 / (1.0 + float(((q - x) ** 2).sum()))
    if sim == O.SIM_DOT:
        return (1.0 + float(q @ x)) / 2.0
    if sim == O.SIM_COSINE:
        return (1.0 + float(q @ x) / float(np.sqrt((q @ q) * (x @ x)))) / 2.0
    d = float(q @ x)  # MAXIMUM_INNER_PRODUCT
    return 1.0 / (1.0 - d) if d < 0 else d + 1.0
