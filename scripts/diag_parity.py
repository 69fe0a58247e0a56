"""Diagnostic: CUDA vs oracle(fp32) vs oracle(fp64) error table for MGAT cases (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import util

def run64(cfg):
    torch.set_default_dtype(torch.float64)
    try:
        import isg_oracle as O
        from isg_b200 import synth
        C, B, seed, sampler, train = cfg["channels"], cfg["num_graphs"], cfg["seed"], cfg["sampler"], cfg["train"]
        b = synth.make_batch(B, channels=C, mean_nodes=cfg["mean_nodes"], mean_edges=cfg["mean_edges"], seed=seed)
        model = O.OracleMGAT(channels=C, sampler_type=sampler, sample_k=cfg["k"]).double()
        model.load_state_dict({k: v.double() for k, v in synth.make_state_dict(C, 4, 4, seed).items()})
        model.train(train)
        N = b["x"].shape[0]
        noise = util.case_noise(sampler, B, b["nmax"], seed).double()
        drop = util.case_dropout(N, train, seed)
        drop = drop.double() if drop is not None else None
        x = b["x"].double().requires_grad_(True); ea = b["edge_attr"].double().requires_grad_(True)
        iv = b["instr_vectors"].double().requires_grad_(True); gl = b["global_language_feats"].double().requires_grad_(True)
        h, mask, _, _ = model(x, b["edge_index"], iv, gl, ea, b["batch"], noise=noise, theta_dropout_mask=drop)
        w = torch.sin(torch.arange(h.numel(), dtype=torch.float32)).double().view_as(h)
        loss = (h * w).sum() / h.shape[0] + (h * h).mean()
        loss.backward()
        pg = {k: (p.grad.clone() if p.grad is not None else None) for k, p in model.named_ref_parameters()}
        return dict(h=h.detach(), mask=mask.detach(), gx=x.grad, g_edge_attr=ea.grad, g_instr=iv.grad, g_glf=gl.grad, param_grads=pg)
    finally:
        torch.set_default_dtype(torch.float32)

cases = [
    dict(sampler="imle", train=False, channels=300, num_graphs=5, mean_nodes=8, mean_edges=40, k=2, seed=102, steps=1),
    dict(sampler="imle", train=True, channels=300, num_graphs=5, mean_nodes=8, mean_edges=40, k=2, seed=101, steps=1),
    dict(sampler="gumbel", train=True, channels=300, num_graphs=5, mean_nodes=8, mean_edges=40, k=2, seed=104, steps=1),
    dict(sampler="imle", train=True, channels=300, num_graphs=64, mean_nodes=20, mean_edges=150, k=2, seed=964, steps=1),
]
for cfg in cases:
    o32 = util.run_oracle_case(cfg)[0]
    o64 = run64(cfg)
    cu = util.run_cuda_case(cfg)[0]
    print("==", cfg["sampler"], "train" if cfg["train"] else "eval", "B", cfg["num_graphs"], "mask eq(cu,o32)", bool(torch.equal(cu["mask"], o32["mask"])), "mask eq(o32,o64)", float((o32["mask"].double()-o64["mask"]).abs().max()))
    for k in ("h", "gx", "g_edge_attr", "g_instr", "g_glf"):
        print(f"   {k:12s} cu-o64 {util.rel_err(cu[k], o64[k]):.2e}  o32-o64 {util.rel_err(o32[k], o64[k]):.2e}  cu-o32 {util.rel_err(cu[k], o32[k]):.2e}")
    rows = []
    for name, w in o64["param_grads"].items():
        if w is None: continue
        rows.append((util.rel_err(cu["param_grads"][name], w), util.rel_err(o32["param_grads"][name], w), name))
    rows.sort(reverse=True)
    for r in rows[:6]:
        print(f"   pgrad {r[2]:34s} cu-o64 {r[0]:.2e}  o32-o64 {r[1]:.2e}")
