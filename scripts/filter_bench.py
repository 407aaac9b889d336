"""Filtered queries at bench scale (cfg-4 flavour: Bernoulli accept bitset over docIds): QPS and recall of the current path."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch, jvpkg, bench
jv = jvpkg.load()
N = jv.native
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2-1Mx768-dot-pq192"
sel = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
width = int(sys.argv[4]) if len(sys.argv) > 4 else 0
w = dict(bench.WORKLOADS[wl])
host, d_queries = bench.build_fixture(torch, jv, w, 0, 1234, w["n"], lambda m: None)
k, rk = w["k"], w["k"] * w["over"]
gi = jv.GpuIndex(w["sim"], host["base"], host["adj"], host["entry"], pq_m=w["pq_m"], pq_k=256, pq_codebooks=host["cb"], pq_codes=host["codes"],
                 flags=N.FLAG_LUT_U8)
rng = np.random.default_rng(3236)
mask = rng.random(w["n"]) < sel
bits = jv.make_accept_bits(mask)
q = host["queries"][:nq]
gt, _, _ = gi.exact_topk(q, k, accept_bits=bits)
for _ in range(2):
    r = gi.search(q, k, rk, accept_bits=bits, expand_width=width)
t0 = time.perf_counter()
r = gi.search(q, k, rk, accept_bits=bits, expand_width=width)
dt = time.perf_counter() - t0
rec = float(np.mean([len(set(a[a >= 0].tolist()) & set(b.tolist())) / k for a, b in zip(r.docs, gt)]))
print(f"selectivity {sel}: {nq / dt:,.0f} queries/s (search_ms {r.timing['search_ms']:.2f}, rerank_ms {r.timing['rerank_ms']:.2f}) recall@{k} {rec:.4f} "
      f"visited/query {r.stats[:, 0].mean():.0f} expanded/query {r.stats[:, 1].mean():.0f} overflows {gi.visited_overflows()}")
