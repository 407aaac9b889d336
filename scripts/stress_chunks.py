"""Stress the chunked H2D pipeline of jv_search_batch with pinned host queries (diagnostics)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch, jvpkg
jv = jvpkg.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
flags = int(sys.argv[3]) if len(sys.argv) > 3 else jv.native.FLAG_LUT_U8
dim, m, nq = 768, 192, 10000
rng = np.random.default_rng(0)
base = rng.standard_normal((n, dim)).astype(np.float32)
base /= np.linalg.norm(base, axis=1, keepdims=True)
q = rng.standard_normal((nq, dim)).astype(np.float32)
q /= np.linalg.norm(q, axis=1, keepdims=True)
cb, g = jv.pq_train(base[:4000], m, 256, False, 2, 1)
codes = jv.pq_encode(base, m, 256, cb, g)
adj, entry = jv.graph_build(base, 1, 32, 100, 1.2, 1.2)
hq = torch.from_numpy(q).pin_memory()
with jv.GpuIndex(1, base, adj, entry, pq_m=m, pq_k=256, pq_codebooks=cb, pq_codes=codes, flags=flags) as gi:
    ref = gi.search(q[:2000], 10, 50)
    t0 = time.time()
    for i in range(iters):
        r = gi.search(hq.numpy(), 10, 50)
        if not np.array_equal(r.docs[:2000], ref.docs):
            print("MISMATCH at iteration", i, int((r.docs[:2000] != ref.docs).any(axis=1).sum()))
            break
    print("done", i + 1, "iterations", f"{(time.time() - t0) / (i + 1) * 1e3:.2f} ms/iter")
