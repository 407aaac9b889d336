"""N > 1 path on CPU: world_size-2 gloo processes exercise the sharding protocol of sharded.py (doc-id offsets,
query broadcast, all-gather layout [g][nq][k], merge with ties -> lower global docId) with the oracle standing in
for the per-shard search and the merge kernels.  The sharded result must equal a single unsharded exact search."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank: int, world: int, port: int, out_dir: str):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    import jvpkg
    from oracle import oracle as O

    jv = jvpkg.load()
    from opensearch_jvector_b200.sharded import ShardedSearcher, partition

    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)  # same data on every rank; each rank keeps its slice
    n, dim, nq, k = 3001, 24, 17, 10
    base = rng.standard_normal((n, dim)).astype(np.float32)
    base[1500] = base[10]  # an exact tie across shards: the lower global docId must win
    queries = rng.standard_normal((nq, dim)).astype(np.float32)
    spec = partition(n, world)[rank]
    shard = base[spec.begin:spec.end]
    adj = np.full((shard.shape[0], 2), -1, np.int32)
    ora = O.OracleIndex(O.SIM_EUCLIDEAN, shard, adj, 0)

    def local_search(q, kk):
        d, s, _ = ora.exact_topk(q.numpy(), kk)
        return torch.from_numpy(d), torch.from_numpy(s)

    def merge(gd, gs, kk):
        d, s, _ = O.merge_topk(gd.numpy(), gs.numpy(), kk)
        return torch.from_numpy(d), torch.from_numpy(s)

    searcher = ShardedSearcher(dist, spec, local_search, merge)
    q = torch.from_numpy(queries.copy()) if rank == 0 else torch.zeros(nq, dim)  # only rank 0 holds the batch
    docs, scores = searcher.search(q, k)
    np.save(os.path.join(out_dir, f"docs{rank}.npy"), docs.numpy())
    np.save(os.path.join(out_dir, f"scores{rank}.npy"), scores.numpy())
    dist.destroy_process_group()


def test_two_rank_sharded_search_matches_unsharded(tmp_path):
    import torch.multiprocessing as mp

    from oracle import oracle as O

    world, port = 2, 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(7)
    n, dim, nq, k = 3001, 24, 17, 10
    base = rng.standard_normal((n, dim)).astype(np.float32)
    base[1500] = base[10]
    queries = rng.standard_normal((nq, dim)).astype(np.float32)
    full = O.OracleIndex(O.SIM_EUCLIDEAN, base, np.full((n, 2), -1, np.int32), 0)
    wd, ws, _ = full.exact_topk(queries, k)
    for r in range(world):
        d = np.load(tmp_path / f"docs{r}.npy")
        s = np.load(tmp_path / f"scores{r}.npy")
        np.testing.assert_array_equal(d, wd)
        np.testing.assert_array_equal(s, ws)
    # the planted duplicate: wherever doc 10 and doc 1500 both appear, 10 comes first
    qd = base[10][None, :]
    d2, _, _ = full.exact_topk(qd, 3)
    assert list(d2[0][:2]) == [10, 1500]


def test_partition_covers_all_docs():
    import jvpkg
    jvpkg.load()
    from opensearch_jvector_b200.sharded import partition
    for n in (1, 7, 1000, 1_000_003):
        for w in (1, 2, 4, 8):
            parts = partition(n, w)
            assert parts[0].begin == 0 and parts[-1].end == n
            assert all(a.end == b.begin for a, b in zip(parts, parts[1:]))
            assert sum(p.size for p in parts) == n
