"""GPU box: time of a leading-segment merge (jv_graph_extend_dev) against a rebuild from scratch (jv_graph_build_dev) on the cfg2 data.
usage: python scripts/merge_bench.py [n_total] [n_leading]"""
import ctypes as C, sys, time
sys.path.insert(0, ".")
import torch, jvpkg, bench
jv = jvpkg.load()
N = jv.native
lib = N.load()
w = dict(bench.WORKLOADS["cfg2-1Mx768-dot-pq192"])
n = int(sys.argv[1]) if len(sys.argv) > 1 else w["n"]
n0 = int(sys.argv[2]) if len(sys.argv) > 2 else n * 4 // 5
dev = torch.device("cuda", 0)
base, queries = bench.gen_data(torch, w, dev, 1234, n, 1000)
R = w["R"]
adj0 = torch.empty(n0, R, dtype=torch.int32, device=dev)
entry = C.c_int32(0)
t0 = time.time()
N.check(lib.jv_graph_build_dev(0, base.data_ptr(), n0, w["dim"], w["sim"], R, 100, 1.2, 1.2, adj0.data_ptr(), C.addressof(entry)))
torch.cuda.synchronize(); t_lead = time.time() - t0
adj = torch.empty(n, R, dtype=torch.int32, device=dev)
t0 = time.time()
N.check(lib.jv_graph_extend_dev(0, base.data_ptr(), n, n0, adj0.data_ptr(), entry.value, w["dim"], w["sim"], R, 100, 1.2, 1.2, adj.data_ptr()))
torch.cuda.synchronize(); t_ext = time.time() - t0
full = torch.empty(n, R, dtype=torch.int32, device=dev)
e2 = C.c_int32(0)
t0 = time.time()
N.check(lib.jv_graph_build_dev(0, base.data_ptr(), n, w["dim"], w["sim"], R, 100, 1.2, 1.2, full.data_ptr(), C.addressof(e2)))
torch.cuda.synchronize(); t_full = time.time() - t0
print(f"n={n} leading={n0}: leading build {t_lead:.2f}s, extend by {n - n0} nodes {t_ext:.2f}s, rebuild from scratch {t_full:.2f}s")
for name, a, e in (("extended", adj, entry.value), ("rebuilt", full, e2.value)):
    gi = jv.GpuIndex(w["sim"], base.cpu().numpy(), a.cpu().numpy(), e)
    q = queries.cpu().numpy()
    res = gi.search(q, 10, 50)
    gd, _, _ = gi.exact_topk(q, 10)
    print(f"  {name}: recall@10 {bench.recall_at_k(res.docs, gd):.4f} (exact traversal, no PQ), visited/query {res.stats[:, 0].mean():.0f}, mean degree {(a >= 0).sum(1).float().mean().item():.1f}")
    gi.close()
# delete consolidation at scale: 10 % of the nodes deleted
dead = torch.rand(n, device=dev, generator=torch.Generator(device=dev).manual_seed(5)) < 0.10
dead_u8 = dead.to(torch.uint8).contiguous()
cons = torch.empty(n, R, dtype=torch.int32, device=dev)
e3 = C.c_int32(0)
t0 = time.time()
N.check(lib.jv_graph_remove_deleted_dev(0, base.data_ptr(), n, w["dim"], w["sim"], R, 1.2, full.data_ptr(), dead_u8.data_ptr(), e2.value,
                                        cons.data_ptr(), C.addressof(e3)))
torch.cuda.synchronize(); t_cons = time.time() - t0
touched = int(((cons != full).any(1) & ~dead).sum().item())
print(f"delete consolidation: {int(dead.sum().item())} of {n} nodes deleted, {touched} live rows repaired in {t_cons:.2f}s")
