"""Diagnostic (GPU box): where does recall@10 go at 1M x 768?  graph quality vs PQ approximation."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import torch, jvpkg, bench
jv = jvpkg.load()
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2-1Mx768-dot-pq192"
w = dict(bench.WORKLOADS[wl])
for kv in sys.argv[2:]:
    k_, v_ = kv.split("=")
    w[k_] = int(v_)
print("workload", w, flush=True)
host, dq = bench.build_fixture(torch, jv, w, 0, 1234, w["n"], lambda m: print("[diag]", m, flush=True))
nq = 1000
q = host["queries"][:nq]
rec = lambda f, t: float(np.mean([len(set(a.tolist()) & set(b.tolist())) / t.shape[1] for a, b in zip(f, t)]))
gi = jv.GpuIndex(w["sim"], host["base"], host["adj"], host["entry"], pq_m=w["pq_m"], pq_k=256, pq_codebooks=host["cb"], pq_codes=host["codes"])
gt, _, _ = gi.exact_topk(q, 10)
for L in (50, 100):
    for E in (1, 4):
        r = gi.search(q, 10, L, expand_width=E)
        print(f"PQ   L={L} E={E} recall@10={rec(r.docs, gt):.4f} visited={r.stats[:,0].mean():.0f} expanded={r.stats[:,1].mean():.0f}", flush=True)
# ADC candidate recall (no graph): top-L by ADC over ALL nodes
nn = 100
n = w["n"]
adc = np.empty((nn, n), np.float32)
step = 250_000
for s in range(0, n, step):
    nodes = np.tile(np.arange(s, min(n, s + step), dtype=np.int32), (nn, 1))
    adc[:, s:s + nodes.shape[1]] = gi.adc_scores(q[:nn], nodes)
for L in (50, 100, 200):
    top = np.argpartition(-adc, L, axis=1)[:, :L]
    print(f"ADC-only top-{L} contains GT@10: {rec(top, gt[:nn]):.4f}", flush=True)
gi.close()
gx = jv.GpuIndex(w["sim"], host["base"], host["adj"], host["entry"])
for L in (50, 100, 200):
    r = gx.search(q, 10, L, expand_width=1)
    print(f"EXACT-score traversal L={L} recall@10={rec(r.docs, gt):.4f} visited={r.stats[:,0].mean():.0f} expanded={r.stats[:,1].mean():.0f}", flush=True)
deg = (host["adj"] >= 0).sum(1)
print("degree mean/min", deg.mean(), deg.min(), "entry", host["entry"])
