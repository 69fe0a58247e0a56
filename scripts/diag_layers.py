"""Diagnostic: per-layer forward values and gradients, CUDA vs oracle fp32 (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import util
import isg_oracle as O
from isg_b200 import synth
from isg_b200.isubgvqa import MGAT

def run(cfg):
    C, B, seed, sampler, train = cfg["channels"], cfg["num_graphs"], cfg["seed"], cfg["sampler"], cfg["train"]
    b = synth.make_batch(B, channels=C, mean_nodes=cfg["mean_nodes"], mean_edges=cfg["mean_edges"], seed=seed)
    sd = synth.make_state_dict(C, 4, 4, seed)
    N = b["x"].shape[0]
    noise = util.case_noise(sampler, B, b["nmax"], seed); drop = util.case_dropout(N, train, seed)
    # oracle
    om = O.OracleMGAT(channels=C, sampler_type=sampler, sample_k=cfg["k"]); om.load_state_dict(sd); om.train(train)
    x = b["x"].clone().requires_grad_(True); ea = b["edge_attr"].clone().requires_grad_(True)
    iv = b["instr_vectors"].clone().requires_grad_(True); gl = b["global_language_feats"].clone().requires_grad_(True)
    h, mask, aux = om(x, b["edge_index"], iv, gl, ea, b["batch"], noise=noise, theta_dropout_mask=drop, return_aux=True)
    ot = {}
    for i in range(4):
        ot[f"conv_out.{i}"] = aux["conv_out"][i]; ot[f"proj.{i}"] = aux["proj"][i]; ot[f"h.{i}"] = aux["h"][i]
    for t in ot.values(): t.retain_grad()
    util.loss_fn(h).backward()
    # cuda
    cm = MGAT(channels=C, num_ins=4, heads=4, use_instr=True, masking_thresholds=[1.0,1.0,1.0,0.1], use_topk=True, interpretable_mode=False, sampler_type=sampler, sample_k=cfg["k"])
    cm.load_state_dict(sd); cm.cuda(); cm.train(train); cm.debug_tensors = {}
    cm.convs[3].mask.injected_noise = noise.cuda(); cm.convs[3].mask.injected_dropout_mask = drop.cuda() if drop is not None else None
    xc = b["x"].cuda().requires_grad_(True); eac = b["edge_attr"].cuda().requires_grad_(True)
    ivc = b["instr_vectors"].cuda().requires_grad_(True); glc = b["global_language_feats"].cuda().requires_grad_(True)
    hc, maskc, _, _ = cm(xc, b["edge_index"].cuda(), ivc, glc, eac, b["batch"].cuda())
    util.loss_fn(hc).backward()
    print("==", sampler, "train" if train else "eval", "B", B, "mask eq", bool(torch.equal(maskc.cpu(), mask)), "sel", float(mask.sum()))
    for k in sorted(ot, key=lambda s: (int(s.split('.')[1]), s)):
        print(f"   {k:12s} val {util.rel_err(cm.debug_tensors[k], ot[k]):.2e}   grad {util.rel_err(cm.debug_tensors[k].grad, ot[k].grad):.2e}")
    for k, a, bb in (("gx", xc.grad, x.grad), ("g_ea", eac.grad, ea.grad), ("g_iv", ivc.grad, iv.grad), ("g_glf", glc.grad, gl.grad)):
        print(f"   {k:12s} {util.rel_err(a, bb):.2e}")
    cp = dict(cm.named_parameters()); rows = []
    for name, p in om.named_ref_parameters():
        if p.grad is None: continue
        rows.append((util.rel_err(cp[name].grad, p.grad), name))
    rows.sort(reverse=True)
    for r in rows[:10]: print(f"   pgrad {r[1]:34s} {r[0]:.2e}")
    # where is g_ea wrong?  per-edge error, and is the edge masked in layer 3 / what is its dst in-degree
    err = (eac.grad.cpu() - ea.grad).abs().max(dim=1).values
    top = err.topk(5).indices
    ei = b["edge_index"]; deg = torch.bincount(ei[1], minlength=N)
    for e in top.tolist():
        s_, d_ = int(ei[0][e]), int(ei[1][e])
        print(f"   edge {e} ({s_}->{d_}) err {float(err[e]):.2e} ref {float(ea.grad[e].abs().max()):.2e} mask_src {float(mask[s_])} mask_dst {float(mask[d_])} indeg {int(deg[d_])} graph {int(b['batch'][d_])} n_graph {int((b['batch']==b['batch'][d_]).sum())}")

run(dict(sampler="imle", train=False, channels=300, num_graphs=5, mean_nodes=8, mean_edges=40, k=2, seed=102, steps=1))
run(dict(sampler="imle", train=True, channels=300, num_graphs=64, mean_nodes=20, mean_edges=150, k=2, seed=964, steps=1))
