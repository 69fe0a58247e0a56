"""Diagnostic: CUDA vs fp64 arbiter (replayed discrete decisions) vs fp32 oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch, util
cases = [
    dict(sampler="imle", train=False, channels=300, num_graphs=5, mean_nodes=8, mean_edges=40, k=2, seed=102, steps=1),
    dict(sampler="gumbel", train=True, channels=300, num_graphs=5, mean_nodes=8, mean_edges=40, k=2, seed=104, steps=1),
    dict(sampler="imle", train=True, channels=300, num_graphs=64, mean_nodes=20, mean_edges=150, k=2, seed=964, steps=1),
]
for cfg in cases:
    o32, o64 = util.run_oracle_fp64_arbiter(cfg)
    o32, o64 = o32[0], o64[0]
    cu = util.run_cuda_case(cfg)[0]
    print("==", cfg["sampler"], "train" if cfg["train"] else "eval", "B", cfg["num_graphs"], "mask eq cu/o32", bool(torch.equal(cu["mask"], o32["mask"])))
    for k in ("h", "gx", "g_edge_attr", "g_instr", "g_glf"):
        print(f"   {k:12s} cu-o64 {util.rel_err(cu[k], o64[k]):.2e}  o32-o64 {util.rel_err(o32[k], o64[k]):.2e}  cu-o32 {util.rel_err(cu[k], o32[k]):.2e}")
    rows = []
    for name, w in o64["param_grads"].items():
        if w is None: continue
        rows.append((util.rel_err(o32["param_grads"][name], w), util.rel_err(cu["param_grads"][name], w), name))
    rows.sort(reverse=True)
    for r in rows[:8]:
        print(f"   pgrad {r[2]:34s} cu-o64 {r[1]:.2e}  o32-o64 {r[0]:.2e}")
    print("   worst cu-o64 over all param grads:", max(r[1] for r in rows))
