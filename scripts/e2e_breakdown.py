"""Where the end-to-end time of jv_search_batch goes (pinned host buffers, cfg2): device events vs wall clock."""
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
k, rk, nq, dim = w["k"], w["k"] * w["over"], w["nq"], w["dim"]
gi = jv.GpuIndex(w["sim"], host["base"], host["adj"], host["entry"], pq_m=w["pq_m"], pq_k=256, pq_codebooks=host["cb"], pq_codes=host["codes"],
                 flags=N.FLAG_LUT_U8)
hq = torch.from_numpy(host["queries"]).pin_memory()
h_doc = torch.empty(nq, k, dtype=torch.int32).pin_memory()
h_score = torch.empty(nq, k, dtype=torch.float32).pin_memory()
h_cnt = torch.empty(nq, dtype=torch.int32).pin_memory()
h_stats = torch.empty(nq, 4, dtype=torch.int32).pin_memory()
p = gi._params(k, rk, 0.0, 0.0, None, 0, 0)
t = N.BatchTiming()
for label, st in (("with stats", h_stats.data_ptr()), ("no stats", None)):
    for _ in range(3):
        N.check(lib.jv_search_batch(gi.handle, hq.data_ptr(), nq, C.addressof(p), h_doc.data_ptr(), h_score.data_ptr(), h_cnt.data_ptr(), st, C.addressof(t)))
    walls, devs = [], []
    for _ in range(20):
        t0 = time.perf_counter()
        N.check(lib.jv_search_batch(gi.handle, hq.data_ptr(), nq, C.addressof(p), h_doc.data_ptr(), h_score.data_ptr(), h_cnt.data_ptr(), st, C.addressof(t)))
        walls.append((time.perf_counter() - t0) * 1e3)
        devs.append(t.total_ms)
    print(f"{label}: wall {np.median(walls):.3f} ms  device ev0->ev4 {np.median(devs):.3f} ms  (first chunk: h2d {t.h2d_ms:.3f} search {t.search_ms:.3f} rerank {t.rerank_ms:.3f})")
