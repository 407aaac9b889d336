"""Small run of the 8-bit table path for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import jvpkg
jv = jvpkg.load()
rng = np.random.default_rng(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
SHAPES = ((64, 16, 1), (96, 48, 0), (128, 16, 2), (384, 192, 1), (768, 192, 2))  # M = 192: the manager / expander / scorer kernel
if len(sys.argv) > 2 and sys.argv[2] == "beam":  # racecheck is slow: only the M = 192 shape
    SHAPES = ((384, 192, 1),)
for dim, m, sim in SHAPES:
    cent = rng.standard_normal((16, dim)).astype(np.float32)
    base = (cent[rng.integers(0, 16, n)] + 0.4 * rng.standard_normal((n, dim))).astype(np.float32)
    q = base[:33] + 0.01
    cb, g = jv.pq_train(base[:2000], m, 256, sim == 0, 2, 1)
    codes = jv.pq_encode(base, m, 256, cb, g)
    adj, entry = jv.graph_build(base, sim, 32, 100, 1.2, 1.2)
    with jv.GpuIndex(sim, base, adj, entry, pq_m=m, pq_k=256, pq_codebooks=cb, pq_global_centroid=g, pq_codes=codes,
                     flags=jv.native.FLAG_LUT_U8) as gi:
        for e in (1, 3, 4):
            r = gi.search(q, 10, 50, expand_width=e)
            print("dim", dim, "E", e, "visited", r.stats[:, 0].mean(), "docs0", r.docs[0][:3])
        mask = rng.random(n) < 0.2
        r = gi.search(q, 10, 50, accept_bits=jv.make_accept_bits(mask))  # filtered flavour of the 8-bit path
        print("dim", dim, "filtered visited", r.stats[:, 0].mean(), "accepted only", bool(mask[r.docs[r.docs >= 0]].all()))
print("ok")
