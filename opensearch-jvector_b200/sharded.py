"""Segments / index shards partitioned across the GPUs of one box.

One process per GPU (torch.distributed).  Every rank owns a disjoint set of documents with its own graph, codes and
doc map — exactly like independent Lucene segments / OpenSearch shards, which the reference searches independently
and merges by score (Lucene TopDocs.merge; multi-shard IT JVectorEngineIT.java:175-216).  A query batch is
replicated to all ranks, each rank searches its resident shard, per-rank top-k lists (with docIds made global by the
shard's base offset) are all-gathered and merged on the device (K7, jv_merge_topk_dev), ties -> lower global docId.

torch.distributed is plumbing only: the only collective on the data path is one all-gather of [nq, k] (doc, score)
lists (~800 KB per GPU at nq = 10k, k = 10).  The class is backend-agnostic so the protocol (offsets, gather layout,
merge order) is covered by world_size-2 gloo tests on CPU with injected search / merge callables.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional, Tuple

import numpy as np


@dataclass
class ShardSpec:
    rank: int
    world: int
    n_total: int

    @property
    def begin(self) -> int:
        return (self.n_total * self.rank) // self.world

    @property
    def end(self) -> int:
        return (self.n_total * (self.rank + 1)) // self.world

    @property
    def size(self) -> int:
        return self.end - self.begin


def partition(n_total: int, world: int):
    """Contiguous doc-id ranges per rank (global docId = shard base + local docId)."""
    return [ShardSpec(r, world, n_total) for r in range(world)]


class ShardedSearcher:
    """search(queries) on every rank -> identical merged (docs, scores) on every rank."""

    def __init__(self, dist, spec: ShardSpec, local_search: Callable, merge: Callable, device=None):
        """`local_search(queries[nq,dim], k) -> (docs[nq,k] int32 local ids, -1 padded; scores[nq,k] f32)` as torch
        tensors on `device`; `merge(docs[g,nq,k], scores[g,nq,k], k) -> (docs[nq,k], scores[nq,k])`."""
        self.dist, self.spec, self.local_search, self.merge, self.device = dist, spec, local_search, merge, device

    def search(self, queries, k: int):
        import torch
        dist = self.dist
        q = queries
        if dist is not None and dist.is_initialized() and self.spec.world > 1:
            dist.broadcast(q, src=0)  # the batch is replicated: every shard answers every query
        if getattr(q, "is_cuda", False):
            # the library searches on its own non-blocking stream: the broadcast (NCCL work ordered against torch's current
            # stream) and any producer kernel of `queries` must have finished before it reads them
            torch.cuda.current_stream(q.device).synchronize()
        docs, scores = self.local_search(q, k)
        docs = torch.where(docs >= 0, docs + self.spec.begin, docs)  # global docIds
        if self.spec.world == 1 or dist is None:
            return docs, scores
        nq = docs.shape[0]
        gd = torch.empty((self.spec.world, nq, k), dtype=docs.dtype, device=docs.device)
        gs = torch.empty((self.spec.world, nq, k), dtype=scores.dtype, device=scores.device)
        # concatenation along dim 0 == the [g][nq][k] layout K7 expects (gloo insists on the flat view)
        dist.all_gather_into_tensor(gd.view(self.spec.world * nq, k), docs.contiguous())
        dist.all_gather_into_tensor(gs.view(self.spec.world * nq, k), scores.contiguous())
        return self.merge(gd, gs, k)


def gpu_local_search(index, rerank_k_factor: int = 5, expand_width: int = 0):
    """local_search callable over a GpuIndex using the device-pointer entry point (results stay in HBM)."""
    import torch

    def fn(queries, k):
        nq = queries.shape[0]
        dev = queries.device
        docs = torch.empty((nq, k), dtype=torch.int32, device=dev)
        scores = torch.empty((nq, k), dtype=torch.float32, device=dev)
        counts = torch.empty((nq,), dtype=torch.int32, device=dev)
        index.search_dev(queries.data_ptr(), nq, k, k * rerank_k_factor, docs.data_ptr(), scores.data_ptr(), counts.data_ptr(),
                         expand_width=expand_width)
        return docs, scores

    return fn


def gpu_merge(device_index: int):
    """merge callable on the device (K7)."""
    import ctypes as C

    import torch

    from . import native as N

    def fn(gd, gs, k):
        g, nq, _ = gd.shape
        od = torch.empty((nq, k), dtype=torch.int32, device=gd.device)
        os_ = torch.empty((nq, k), dtype=torch.float32, device=gd.device)
        oc = torch.empty((nq,), dtype=torch.int32, device=gd.device)
        N.check(N.load().jv_merge_topk_dev(device_index, g, nq, k, gd.data_ptr(), gs.data_ptr(), od.data_ptr(), os_.data_ptr(),
                                           oc.data_ptr(), None))
        return od, os_

    return fn
