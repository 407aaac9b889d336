"""Generates tests/golden/golden_small.npz: seeded inputs (java.util.Random, the reference's own fixture generator) and the
oracle's outputs for every stage of the hot path, at a size that fits in the repository.  The file pins the oracle against
drift (tests/test_golden.py, CPU) and gives the device path a committed target that does not depend on the oracle being
built on the GPU box (tests/test_gpu_golden.py).

    python tests/golden/make_golden.py        # rewrites golden_small.npz; commit the result

Inputs are NOT stored when they can be regenerated bit-exactly from a seed (java_random_vectors); everything else is.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402

SEED_BASE, SEED_QUERY = 42, 43  # KNNJVectorTests / JVectorWriterMergeTests seeds
N, DIM, NQ, M, R = 1500, 32, 12, 8, 16


def build():
    base = O.java_random_vectors(N, DIM, SEED_BASE)
    q = O.java_random_vectors(NQ, DIM, SEED_QUERY)
    out = {"n": N, "dim": DIM, "nq": NQ, "m": M, "r": R, "seed_base": SEED_BASE, "seed_query": SEED_QUERY}
    for name, sim in (("l2", O.SIM_EUCLIDEAN), ("dot", O.SIM_DOT), ("cos", O.SIM_COSINE)):
        adj, entry = O.graph_build(base, sim, R, 100)
        cb, g = O.pq_train(base, M, 256, center=(sim == O.SIM_EUCLIDEAN), iters=6, seed=7)
        codes = O.pq_encode(base, M, 256, cb, g)
        ix = O.OracleIndex(sim, base, adj, entry, pq_m=M, pq_k=256, pq_codebooks=cb, pq_global_centroid=g, pq_codes=codes, adc_order=32)
        docs, scores, counts, stats = ix.search(q, 10, 50)
        ed, es, ec = ix.exact_topk(q, 10)
        lut = O.pq_lut(sim, DIM, M, 256, cb, g, q[:2])
        ix8 = O.OracleIndex(sim, base, adj, entry, pq_m=M, pq_k=256, pq_codebooks=cb, pq_global_centroid=g, pq_codes=codes, adc_order=-8)
        q8, q8p = ix8.lut_q8(q[:2])
        d8, s8, c8, st8 = ix8.search(q, 10, 50)
        out.update({f"{name}_adj": adj, f"{name}_entry": entry, f"{name}_cb": cb, f"{name}_codes": codes,
                    f"{name}_docs": docs, f"{name}_scores": scores, f"{name}_counts": counts, f"{name}_stats": stats,
                    f"{name}_exact_docs": ed, f"{name}_exact_scores": es, f"{name}_lut": lut, f"{name}_q8": q8, f"{name}_q8p": q8p,
                    f"{name}_docs8": d8, f"{name}_scores8": s8})
        if g is not None:
            out[f"{name}_gcent"] = g
    # merge path: leading graph over the first 1000 ordinals extended by the rest, then 10 % of the leading nodes (entry
    # included) consolidated away — inputs regenerable from the seeds, outputs stored
    for name, sim in (("l2", O.SIM_EUCLIDEAN), ("cos", O.SIM_COSINE)):
        seed_adj, seed_entry = O.graph_build(base[:1000], sim, R, 100)
        ext = O.graph_extend(base, seed_adj, seed_entry, sim)
        dead = np.zeros(N, bool)
        dead[np.random.default_rng(99).choice(1000, 100, replace=False)] = True
        dead[seed_entry] = True
        cons, cons_entry = O.graph_remove_deleted(base, ext, seed_entry, dead, sim)
        out.update({f"{name}_merge_seed_entry": seed_entry, f"{name}_merge_ext": ext, f"{name}_merge_dead": dead,
                    f"{name}_merge_cons": cons, f"{name}_merge_cons_entry": cons_entry})
    b, prm, gm = O.nvq_encode(base[:64], 2)
    out.update({"nvq_bytes": b, "nvq_params": prm, "nvq_gmean": gm, "nvq_deq": O.nvq_dequantize(b, prm, gm)})
    return out


if __name__ == "__main__":
    data = build()
    path = Path(__file__).with_name("golden_small.npz")
    np.savez_compressed(path, **data)
    print(path, f"{path.stat().st_size / 1024:.0f} KiB", len(data), "arrays")
