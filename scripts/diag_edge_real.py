import os, sys, time
t0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch, util
import isg_oracle as O
from isg_b200 import synth, ops
from isg_b200.isubgvqa import MGAT
print("import", time.time() - t0)
C, B, seed = 300, 5, 102
b = synth.make_batch(B, channels=C, mean_nodes=8, mean_edges=40, seed=seed)
sd = synth.make_state_dict(C, 4, 4, seed)
noise = util.case_noise("imle", B, b["nmax"], seed)
cm = MGAT(channels=C, num_ins=4, heads=4, use_instr=True, masking_thresholds=[1.0,1.0,1.0,0.1], use_topk=True, interpretable_mode=False, sampler_type="imle", sample_k=2)
cm.load_state_dict(sd); cm.cuda(); cm.eval()
cm.convs[3].mask.injected_noise = noise.cuda()
ops._DEBUG_EDGE_BWD = []
xc = b["x"].cuda().requires_grad_(True); eac = b["edge_attr"].cuda().requires_grad_(True)
hc, maskc, _, _ = cm(xc, b["edge_index"].cuda(), b["instr_vectors"].cuda(), b["global_language_feats"].cuda(), eac, b["batch"].cuda())
util.loss_fn(hc).backward()
torch.cuda.synchronize()
ei = b["edge_index"]; N = b["x"].shape[0]
for idx, rec in enumerate(ops._DEBUG_EDGE_BWD):  # order: layer 3, 2, 1, 0
    layer = 3 - idx
    r = {k: (v.detach().cpu().double() if v is not None else None) for k, v in rec.items()}
    xl, xr, ep, att, bias = [r[k].clone().requires_grad_(True) for k in ("x_l", "x_r", "e_proj", "att", "bias")]
    em = r["em"].clone().requires_grad_(True) if r["em"] is not None else None
    out, alpha = O.gat_edge(xl.view(N,4,C), xr.view(N,4,C), ep.view(-1,4,C), att, ei, em)
    out = out.reshape(N, 4*C) + bias
    out.backward(r["g_out"])
    print(f"== layer {layer}: saved alpha vs recomputed {util.rel_err(r['alpha'], alpha):.2e}  saved out vs recomputed {util.rel_err(r['out'], out):.2e}")
    print(f"   g_xl {util.rel_err(r['g_xl'], xl.grad):.2e} g_xr {util.rel_err(r['g_xr'], xr.grad):.2e} g_ep {util.rel_err(r['g_ep'], ep.grad):.2e} g_att {util.rel_err(r['g_att'], att.grad):.2e}")
    e_node = (r["g_xr"] - xr.grad).abs().amax(dim=1)
    bad = (e_node > 1e-5 * float(xr.grad.abs().max())).nonzero().flatten().tolist()
    print("   bad nodes g_xr:", bad)
    ea = (r["alpha"] - alpha.detach()).abs().amax(dim=1)
    bad_e = (ea > 1e-5).nonzero().flatten().tolist()
    print("   bad alpha edges:", bad_e[:20], [ (int(ei[0][e]), int(ei[1][e])) for e in bad_e[:20]])
print("total", time.time() - t0)
