import sys
import numpy as np
sys.path.insert(0, ".")
import jvpkg
jv = jvpkg.load()
rng = np.random.default_rng(0)
n, dim = int(sys.argv[1]), int(sys.argv[2])
cent = rng.standard_normal((16, dim)).astype(np.float32)
base = (cent[rng.integers(0, 16, n)] + 0.4 * rng.standard_normal((n, dim))).astype(np.float32)
adj, entry = jv.graph_build(base, 1, 32, 100, 1.2, 1.2)
print("ok", entry, (adj >= 0).sum(1).mean())
