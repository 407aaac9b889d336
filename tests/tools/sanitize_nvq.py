"""Small NVQ-inline (nvq+pq) run for compute-sanitizer (lives under tests/: it uses the oracle's fixture NVQ encoder).
usage: compute-sanitizer --tool memcheck python tests/tools/sanitize_nvq.py"""
import sys
import numpy as np
sys.path.insert(0, ".")
import jvpkg
from oracle import oracle as O
jv = jvpkg.load()
rng = np.random.default_rng(0)
n, dim, m = 2000, 128, 64
base = rng.standard_normal((n, dim)).astype(np.float32)
q = base[:16] + 0.01
cb, g = jv.pq_train(base, m, 256, True, 2, 1)
codes = jv.pq_encode(base, m, 256, cb, g)
adj, entry = jv.graph_build(base, 0, 16, 100, 1.2, 1.2)
b, prm, gm = O.nvq_encode(base, 2)
with jv.GpuIndex(0, None, adj, entry, pq_m=m, pq_k=256, pq_codebooks=cb, pq_global_centroid=g, pq_codes=codes,
                 nvq_m=2, nvq_bytes=b, nvq_params=prm, nvq_global_mean=gm) as gi:
    for e in (-1, 4):
        r = gi.search(q, 10, 100, expand_width=e)
        print("E", e, r.docs[0][:4], r.scores[0][:2])
print("ok")
