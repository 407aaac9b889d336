"""Large staged batch through jv_search_batch (chunked H2D pipeline) for compute-sanitizer."""
import sys
import numpy as np
sys.path.insert(0, ".")
import jvpkg
jv = jvpkg.load()
rng = np.random.default_rng(0)
n, dim, m = 3000, 64, 16
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
base = rng.standard_normal((n, dim)).astype(np.float32)
q = rng.standard_normal((nq, dim)).astype(np.float32)
cb, g = jv.pq_train(base[:2000], m, 256, False, 2, 1)
codes = jv.pq_encode(base, m, 256, cb, g)
adj, entry = jv.graph_build(base, 1, 32, 100, 1.2, 1.2)
with jv.GpuIndex(1, base, adj, entry, pq_m=m, pq_k=256, pq_codebooks=cb, pq_codes=codes, flags=jv.native.FLAG_LUT_U8) as gi:
    for _ in range(3):
        r = gi.search(q, 10, 50)
    print("ok", r.docs[0][:3], r.stats[:, 0].mean())
