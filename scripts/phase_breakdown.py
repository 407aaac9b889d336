"""GPU box: per-phase cycle breakdown of the fast traversal kernel on a bench workload."""
import sys, json
import numpy as np
sys.path.insert(0, ".")
import torch, jvpkg, bench
jv = jvpkg.load()
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2-1Mx768-dot-pq192"
widths = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["4"])]
table = sys.argv[3] if len(sys.argv) > 3 else "u8"  # u8: phases = setup+table wait, dedupe, select, neighbour rows, scoring, rank merge, emit
w = dict(bench.WORKLOADS[wl])
host, dq = bench.build_fixture(torch, jv, w, 0, 1234, w["n"], lambda m: None)
gi = jv.GpuIndex(w["sim"], host["base"], host["adj"], host["entry"], pq_m=w["pq_m"], pq_k=256, pq_codebooks=host["cb"], pq_codes=host["codes"],
                 flags={"u8": jv.native.FLAG_LUT_U8, "fp16": jv.native.FLAG_LUT_F16}[table])
q = host["queries"]
for E in widths:
    gi.search(q, 10, 50, expand_width=E)
    gi.phase_cycles(reset=True)
    r = gi.search(q, 10, 50, expand_width=E)
    ph = gi.phase_cycles(reset=True)
    nq = len(q)
    steps = ph.pop("steps")
    sub = {k_: ph.pop(k_) for k_ in list(ph) if k_.startswith("sub_")}
    tot = sum(ph.values())
    print(f"E={E} search_ms={r.timing['search_ms']:.3f} steps/query={steps/nq:.1f} cycles/query={tot/nq:.0f} visited={r.stats[:,0].mean():.0f}")
    for k_, v in ph.items():
        per_step = v / max(steps, 1) if k_ in ("select", "neighbour_rows", "scoring", "merge", "table_build") else 0
        print(f"   {k_:15s} {100*v/tot:5.1f}%  {v/nq:9.0f} cyc/query" + (f"  {per_step:7.0f} cyc/step" if per_step else ""))
    for k_, v in sub.items():
        print(f"   (scoring) {k_:15s} {v/max(steps,1):7.0f} cyc/step")
