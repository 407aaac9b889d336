"""GPU box: brute-force exact top-k (K5) at the headline size, tensor-core path against the fp32 kernel.
    python scripts/exact_bench.py [workload]
Prints one JSON line: ms per 10k-query batch for both paths (CUDA-synchronous wall clock around jv_exact_topk_dev), equality of the
results, TFLOP/s of the contraction and the candidates the tensor-core pass hands to the exact re-scoring."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402

import bench  # noqa: E402
import jvpkg  # noqa: E402

jv = jvpkg.load()
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2-1Mx768-dot-pq192"
w = dict(bench.WORKLOADS[wl])
dev = torch.device("cuda", 0)
base, dq = bench.gen_data(torch, w, dev, 1234, w["n"], w["nq"])
n, dim, nq, k = w["n"], w["dim"], w["nq"], w["k"]
adj = np.full((n, 4), -1, np.int32)
adj[:, 0] = (np.arange(n) + 1) % n
gi = jv.GpuIndex(w["sim"], base.cpu().numpy(), adj, 0)
del base
od = [torch.empty(nq, k, dtype=torch.int32, device=dev) for _ in range(2)]
os_ = [torch.empty(nq, k, dtype=torch.float32, device=dev) for _ in range(2)]
oc = torch.empty(nq, dtype=torch.int32, device=dev)
res = {}
for name, env, slot in (("tensor_core", "1", 0), ("fp32", "0", 1)):
    os.environ["JVGPU_EXACT_TC"] = env
    gi.refresh_knobs()
    gi.exact_topk_dev(dq.data_ptr(), nq, k, od[slot].data_ptr(), os_[slot].data_ptr(), oc.data_ptr())  # warm-up (bf16 copy on the first call)
    ts = []
    for _ in range(3 if env == "1" else 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gi.exact_topk_dev(dq.data_ptr(), nq, k, od[slot].data_ptr(), os_[slot].data_ptr(), oc.data_ptr())
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    res[name + "_ms"] = min(ts)
res["tc_counters"] = gi.exact_tc_counters()
res["ids_identical"] = bool((od[0] == od[1]).all().item())
res["score_bits_identical"] = bool((os_[0].view(torch.int32) == os_[1].view(torch.int32)).all().item())
res["contraction_tflops_tc"] = 2.0 * nq * n * dim * (1 + 1.0 / (160 // k)) / (res["tensor_core_ms"] * 1e-3) / 1e12  # pass B + the 1/stride sample of pass A
res["workload"] = wl
print(json.dumps(res), flush=True)
gi.close()
