"""Single-query / small-batch latency of jv_search_batch: wall clock vs the device-side phases the library reports."""
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
h_doc = torch.empty(64, k, dtype=torch.int32).pin_memory()
h_score = torch.empty(64, k, dtype=torch.float32).pin_memory()
h_cnt = torch.empty(64, dtype=torch.int32).pin_memory()
E = int(sys.argv[2]) if len(sys.argv) > 2 else 0  # expand width (0 = library default)
p = gi._params(k, rk, 0.0, 0.0, None, 0, E)
t = N.BatchTiming()
for nq in (1, 8, 64):
    rows = []
    for it in range(200):
        t0 = time.perf_counter()
        N.check(lib.jv_search_batch(gi.handle, hq.data_ptr() + ((it * nq) % 9000) * dim * 4, nq, C.addressof(p), h_doc.data_ptr(), h_score.data_ptr(),
                                    h_cnt.data_ptr(), None, C.addressof(t)))
        rows.append(((time.perf_counter() - t0) * 1e3, t.total_ms, t.h2d_ms, t.lut_ms, t.search_ms, t.rerank_ms, t.d2h_ms))
    r = np.median(np.array(rows[20:]), axis=0)
    print(f"E {E} nq {nq:3d}: wall {r[0]*1e3:6.1f} us | device total {r[1]*1e3:6.1f} us = h2d {r[2]*1e3:5.1f} + search {r[4]*1e3:5.1f} (table build {r[3]*1e3:5.1f}) + rerank {r[5]*1e3:5.1f} + d2h {r[6]*1e3:5.1f}")
