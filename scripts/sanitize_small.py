"""Small end-to-end run for compute-sanitizer: device PQ train + encode + graph build + strict/fast search + rerank."""
import sys
import numpy as np
sys.path.insert(0, ".")
import jvpkg
jv = jvpkg.load()
rng = np.random.default_rng(0)
n, dim, m = int(sys.argv[1]) if len(sys.argv) > 1 else 3000, 64, 16
cent = rng.standard_normal((16, dim)).astype(np.float32)
base = (cent[rng.integers(0, 16, n)] + 0.4 * rng.standard_normal((n, dim))).astype(np.float32)
q = base[:64] + 0.01
cb, _ = jv.pq_train(base[:2000], m, 256, False, 2, 1)
codes = jv.pq_encode(base, m, 256, cb)
adj, entry = jv.graph_build(base, 1, 32, 100, 1.2, 1.2)
with jv.GpuIndex(1, base, adj, entry, pq_m=m, pq_k=256, pq_codebooks=cb, pq_codes=codes) as gi:
    for e in (-1, 1, 4):
        r = gi.search(q, 10, 50, expand_width=e)
        print("E", e, "visited", r.stats[:, 0].mean(), "docs0", r.docs[0][:3])
    mask = rng.random(n) < 0.2
    r = gi.search(q, 10, 50, accept_bits=jv.make_accept_bits(mask))
    d, s, c = gi.exact_topk(q, 10)
with jv.GpuIndex(1, base, adj, entry, pq_m=m, pq_k=256, pq_codebooks=cb, pq_codes=codes, flags=jv.native.FLAG_LUT_F16) as gi:
    r = gi.search(q, 10, 50)
print("ok")
