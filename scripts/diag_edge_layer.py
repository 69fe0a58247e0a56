"""Diagnostic: capture the actual per-layer inputs of the CUDA edge kernel inside an MGAT run and
compare its outputs/gradients against the fp64 oracle edge function on those same tensors."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch, util
import isg_oracle as O
from isg_b200 import synth, ops
from isg_b200.isubgvqa import MGAT

cfg = dict(sampler="imle", train=False, channels=300, num_graphs=5, mean_nodes=8, mean_edges=40, k=2, seed=102, steps=1)
C, B, seed = 300, cfg["num_graphs"], cfg["seed"]
b = synth.make_batch(B, channels=C, mean_nodes=cfg["mean_nodes"], mean_edges=cfg["mean_edges"], seed=seed)
cm = MGAT(channels=C, num_ins=4, heads=4, use_instr=True, masking_thresholds=[1.0,1.0,1.0,0.1], use_topk=True, interpretable_mode=False, sampler_type="imle", sample_k=2)
cm.load_state_dict(synth.make_state_dict(C, 4, 4, seed)); cm.cuda(); cm.eval()
cm.convs[3].mask.injected_noise = util.case_noise("imle", B, b["nmax"], seed).cuda()
captured = []
orig_apply = ops.GatEdge.apply
def spy(x_l, x_r, e_proj, att, bias, edge_mask, gi, heads, slope):
    leaves = [t.detach().clone().requires_grad_(True) if t is not None else None for t in (x_l, x_r, e_proj, att, bias, edge_mask)]
    out, alpha = orig_apply(*leaves, gi, heads, slope)
    rec = dict(leaves=leaves, out=out, alpha=alpha)
    captured.append(rec)
    # re-attach to the outer graph through a pass-through function
    class Bridge(torch.autograd.Function):
        @staticmethod
        def forward(ctx, *ins):
            return out.detach().clone()
        @staticmethod
        def backward(ctx, g):
            rec["g_out"] = g.detach().clone()
            out.backward(g)
            return tuple(l.grad if l is not None else None for l in leaves)
    o2 = Bridge.apply(*[t for t in (x_l, x_r, e_proj, att, bias, edge_mask)])
    return o2, alpha
ops.GatEdge.apply = spy
import isg_b200.isubgvqa.mgat_v2_conv as mv
x = b["x"].cuda().requires_grad_(True)
h, mask, _, _ = cm(x, b["edge_index"].cuda(), b["instr_vectors"].cuda(), b["global_language_feats"].cuda(), b["edge_attr"].cuda().requires_grad_(True), b["batch"].cuda())
util.loss_fn(h).backward()
ei = b["edge_index"]; N = b["x"].shape[0]
for li, rec in enumerate(captured):
    xl, xr, ep, att, bias, em = [t.detach().cpu().double().requires_grad_(True) if t is not None else None for t in rec["leaves"]]
    out, alpha = O.gat_edge(xl.view(N,4,C), xr.view(N,4,C), ep.view(-1,4,C), att, ei, em)
    out = out.reshape(N, 4*C) + bias
    out.backward(rec["g_out"].cpu().double())
    print(f"== layer {li}: |x_l|max {float(xl.abs().max()):.2f} alpha max-per-seg mean {float(torch.zeros(N,4,dtype=torch.float64).index_reduce_(0, ei[1], alpha.detach(), 'amax', include_self=True).mean()):.4f}")
    print(f"   out {util.rel_err(rec['out'], out):.2e} alpha {util.rel_err(rec['alpha'], alpha):.2e}")
    for name, lc, lo in zip(("g_xl","g_xr","g_ep","g_att","g_bias","g_mask"), rec["leaves"], (xl,xr,ep,att,bias,em)):
        if lc is None: continue
        print(f"   {name:7s} rel {util.rel_err(lc.grad, lo.grad):.2e}   |ref|max {float(lo.grad.abs().max()):.3e}")
    # fp32 oracle on the same tensors
    l32 = [t.detach().cpu().float().requires_grad_(True) if t is not None else None for t in rec["leaves"]]
    o32, a32 = O.gat_edge(l32[0].view(N,4,C), l32[1].view(N,4,C), l32[2].view(-1,4,C), l32[3], ei, l32[5])
    (o32.reshape(N,4*C) + l32[4]).backward(rec["g_out"].cpu())
    print(f"   [fp32 oracle vs fp64] g_xr {util.rel_err(l32[1].grad, xr.grad):.2e} g_ep {util.rel_err(l32[2].grad, ep.grad):.2e} g_xl {util.rel_err(l32[0].grad, xl.grad):.2e}")
