"""Replays bench.py's call sequence (device-pointer steps, then pinned-host e2e steps) in a loop (diagnostics)."""
import sys, time, ctypes as C
import numpy as np
sys.path.insert(0, ".")
import torch, jvpkg, bench
jv = jvpkg.load()
N = jv.native
lib = N.load()
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2-small-100kx768"
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 30
w = dict(bench.WORKLOADS[wl])
host, d_queries = bench.build_fixture(torch, jv, w, 0, 1234, w["n"], lambda m: None)
k, rk, nq, dim = w["k"], w["k"] * w["over"], w["nq"], w["dim"]
gi = jv.GpuIndex(w["sim"], host["base"], host["adj"], host["entry"], pq_m=w["pq_m"], pq_k=256, pq_codebooks=host["cb"], pq_codes=host["codes"],
                 flags=N.FLAG_LUT_U8)
dev = torch.device("cuda", 0)
out_doc = torch.empty(nq, k, dtype=torch.int32, device=dev)
out_score = torch.empty(nq, k, dtype=torch.float32, device=dev)
out_count = torch.empty(nq, dtype=torch.int32, device=dev)
stats = torch.empty(nq, 4, dtype=torch.int32, device=dev)
hq = torch.from_numpy(host["queries"]).pin_memory()
h_doc = torch.empty(nq, k, dtype=torch.int32).pin_memory()
h_score = torch.empty(nq, k, dtype=torch.float32).pin_memory()
h_cnt = torch.empty(nq, dtype=torch.int32).pin_memory()
h_stats = torch.empty(nq, 4, dtype=torch.int32).pin_memory()
p = gi._params(k, rk, 0.0, 0.0, None, 0, 0)
for r in range(rounds):
    for _ in range(5):
        gi.search_dev(d_queries.data_ptr(), nq, k, rk, out_doc.data_ptr(), out_score.data_ptr(), out_count.data_ptr(), stats.data_ptr())
    ref = out_doc.cpu().numpy()
    for _ in range(5):
        N.check(lib.jv_search_batch(gi.handle, hq.data_ptr(), nq, C.addressof(p), h_doc.data_ptr(), h_score.data_ptr(), h_cnt.data_ptr(),
                                    h_stats.data_ptr(), None))
    if not np.array_equal(h_doc.numpy(), ref):
        bad = np.nonzero((h_doc.numpy() != ref).any(axis=1))[0]
        truth = gi.search(host["queries"], k, rk).docs  # pageable path
        print("MISMATCH in round", r, "queries", len(bad), "first", bad[:5], "last", bad[-5:],
              "dev==pageable", np.array_equal(ref, truth), "pinned==pageable", np.array_equal(h_doc.numpy(), truth))
        print("host row", h_doc.numpy()[bad[0]], "dev row", ref[bad[0]], "cnt", h_cnt.numpy()[bad[0]], "stats", h_stats.numpy()[bad[0]])
        break
print("done", rounds, "rounds")
