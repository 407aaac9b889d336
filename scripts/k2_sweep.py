"""GPU box: one fixture, many traversal-kernel configurations (kernel flavour, warps per CTA, expansions in flight, CTAs per SM).

    python scripts/k2_sweep.py [workload] [spec ...]        spec = name:ENV=V,ENV=V:E      e.g.  beam8:JVGPU_Q8_WARPS=8:4

Prints one JSON line per configuration: K2 kernel ms (CUDA events around the traversal launch), recall@k against the exact
top-k, visited / expanded per query and the fraction of the measured HBM roofline on the GPU's own visit count."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402

import bench  # noqa: E402
import jvpkg  # noqa: E402

jv = jvpkg.load()
args = sys.argv[1:]
wl = args[0] if args and ":" not in args[0] else "cfg2-1Mx768-dot-pq192"
specs = [a for a in args if ":" in a] or [
    "sync-e4:JVGPU_Q8_SYNC=1:4", "beam-d1-e1:JVGPU_Q8_DEPTH=1:1", "beam-d1-e4:JVGPU_Q8_DEPTH=1:4", "beam-d2-e2:JVGPU_Q8_DEPTH=2:2",
    "beam-d2-e3:JVGPU_Q8_DEPTH=2:3", "beam-d2-e4:JVGPU_Q8_DEPTH=2:4"]
KNOBS = ("JVGPU_Q8_SYNC", "JVGPU_Q8_DEPTH", "JVGPU_Q8_WARPS", "JVGPU_Q8_OCC", "JVGPU_PROFILE", "JVGPU_Q8_FUSED")

w = dict(bench.WORKLOADS[wl])
host, dq = bench.build_fixture(torch, jv, w, 0, 1234, w["n"], lambda m: print("[sweep]", m, file=sys.stderr, flush=True))
k, rk, nq, m, R, dim = w["k"], w["k"] * w["over"], w["nq"], w["pq_m"], w["R"], w["dim"]
gi = jv.GpuIndex(w["sim"], host["base"], host["adj"], host["entry"], pq_m=m, pq_k=256, pq_codebooks=host["cb"], pq_global_centroid=host.get("gcent"),
                 pq_codes=host["codes"], flags=jv.native.FLAG_LUT_U8)
dev = torch.device("cuda", 0)
od = torch.empty(nq, k, dtype=torch.int32, device=dev)
os_ = torch.empty(nq, k, dtype=torch.float32, device=dev)
oc = torch.empty(nq, dtype=torch.int32, device=dev)
st = torch.empty(nq, 4, dtype=torch.int32, device=dev)
gd = torch.empty(nq, k, dtype=torch.int32, device=dev)
gs = torch.empty(nq, k, dtype=torch.float32, device=dev)
gc = torch.empty(nq, dtype=torch.int32, device=dev)
gi.exact_topk_dev(dq.data_ptr(), nq, k, gd.data_ptr(), gs.data_ptr(), gc.data_ptr())
truth = gd.cpu().numpy()
peak, _ = bench.measured_peak()

for spec in specs:
    name, envs, E = spec.split(":")
    for kn in KNOBS:
        os.environ.pop(kn, None)
    for kv in filter(None, envs.split(",")):
        a, b = kv.split("=")
        os.environ[a] = b
    gi.refresh_knobs()
    try:
        for _ in range(3):
            gi.search_dev(dq.data_ptr(), nq, k, rk, od.data_ptr(), os_.data_ptr(), oc.data_ptr(), st.data_ptr(), expand_width=int(E))
        ts = [gi.search_dev(dq.data_ptr(), nq, k, rk, od.data_ptr(), os_.data_ptr(), oc.data_ptr(), st.data_ptr(), expand_width=int(E))
              for _ in range(7)]
    except Exception as e:  # a broken configuration must not stop the sweep
        print(json.dumps({"name": name, "error": str(e)}), flush=True)
        break
    k2 = float(np.median([t["search_ms"] - t.get("lut_ms", 0.0) for t in ts]))
    s = st.cpu().numpy()
    adc, _ = bench.algorithmic_bytes(s, m, R, dim)
    found = od.cpu().numpy()
    first = found.copy()
    gi.search_dev(dq.data_ptr(), nq, k, rk, od.data_ptr(), os_.data_ptr(), oc.data_ptr(), st.data_ptr(), expand_width=int(E))
    again = od.cpu().numpy()
    prof = None
    if "JVGPU_PROFILE" in envs:  # per-phase cycles per expansion (pipelined kernel) / per step (synchronous kernel)
        gi.phase_cycles(reset=True)
        gi.search_dev(dq.data_ptr(), nq, k, rk, od.data_ptr(), os_.data_ptr(), oc.data_ptr(), st.data_ptr(), expand_width=int(E))
        ph = gi.phase_cycles(reset=True)
        vals = list(ph.values())
        div = max(vals[7], 1)
        prof = {"per": "step", "count_per_query": round(vals[7] / nq, 1),
                "cycles": [round(v / div) for v in vals[:7]] + [round(v / div) for v in vals[8:]]}
    print(json.dumps({"name": name, "prof": prof, "env": envs, "E": int(E), "k2_ms": round(k2, 4), "lut_ms": round(float(np.median([t.get("lut_ms", 0) for t in ts])), 4),
                      "rerank_ms": round(float(np.median([t["rerank_ms"] for t in ts])), 4),
                      "recall": round(bench.recall_at_k(found, truth), 5), "visited": round(float(s[:, 0].mean()), 1),
                      "expanded": round(float(s[:, 1].mean()), 1), "frac": round(adc / (k2 * 1e-3) / 1e9 / peak, 4),
                      "mqps_k2": round(nq / k2 / 1e3, 3), "same_rows_run_to_run": round(float((first == again).all(axis=1).mean()), 5)}), flush=True)
gi.close()
