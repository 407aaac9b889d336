"""GPU box, under compute-sanitizer: the merge-path kernels (seeded build, delete consolidation) at a small size.
usage: compute-sanitizer --tool memcheck|racecheck|synccheck python scripts/sanitize_merge.py [n] [dim]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import jvpkg
jv = jvpkg.load()
rng = np.random.default_rng(0)
n, dim = (int(sys.argv[1]) if len(sys.argv) > 1 else 1500), (int(sys.argv[2]) if len(sys.argv) > 2 else 24)
cent = rng.standard_normal((16, dim)).astype(np.float32)
base = (cent[rng.integers(0, 16, n)] + 0.4 * rng.standard_normal((n, dim))).astype(np.float32)
n0 = n * 2 // 3
adj0, entry = jv.graph_build(base[:n0], 2, 16, 100, 1.2, 1.2)
adj = jv.graph_extend(base, adj0, entry, 2, 100)
dead = rng.random(n) < 0.3
dead[entry] = True
out, e2 = jv.graph_remove_deleted(base, adj, entry, dead, 2)
assert (out[dead] == -1).all() and not dead[e2]
print("ok", entry, e2, (out[~dead] >= 0).sum(1).mean())
