"""Latency / throughput of jv_search_batch vs batch size (pinned host buffers, cfg2 index): the serving view of the path."""
import sys, time, ctypes as C
import numpy as np
sys.path.insert(0, ".")
import torch, jvpkg, bench
jv = jvpkg.load()
N = jv.native
lib = N.load()
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2-1Mx768-dot-pq192"
w = dict(bench.WORKLOADS[wl])
host, d_queries = bench.build_fixture(torch, jv, w, 0, 1234, w["n"], lambda m: None)
k, rk, dim = w["k"], w["k"] * w["over"], w["dim"]
gi = jv.GpuIndex(w["sim"], host["base"], host["adj"], host["entry"], pq_m=w["pq_m"], pq_k=256, pq_codebooks=host["cb"],
                 pq_global_centroid=host.get("gcent"), pq_codes=host["codes"], flags=N.FLAG_LUT_U8)
hq = torch.from_numpy(host["queries"]).pin_memory()
nqmax = hq.shape[0]
h_doc = torch.empty(nqmax, k, dtype=torch.int32).pin_memory()
h_score = torch.empty(nqmax, k, dtype=torch.float32).pin_memory()
h_cnt = torch.empty(nqmax, dtype=torch.int32).pin_memory()
p = gi._params(k, rk, 0.0, 0.0, None, 0, 0)
print(f"{'batch':>7s} {'median ms':>10s} {'p99 ms':>9s} {'queries/s':>12s}")
for nq in (1, 8, 64, 512, 2048, 10000):
    ts = []
    for it in range(60):
        off = (it * nq) % max(1, nqmax - nq)
        t0 = time.perf_counter()
        N.check(lib.jv_search_batch(gi.handle, hq.data_ptr() + off * dim * 4, nq, C.addressof(p), h_doc.data_ptr(), h_score.data_ptr(),
                                    h_cnt.data_ptr(), None, None))
        ts.append((time.perf_counter() - t0) * 1e3)
    ts = np.sort(ts[10:])
    print(f"{nq:7d} {np.median(ts):10.3f} {ts[int(len(ts) * 0.99)]:9.3f} {nq / (np.median(ts) * 1e-3):12,.0f}")
