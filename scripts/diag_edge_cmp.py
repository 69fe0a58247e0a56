import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch, util
import isg_oracle as O
from isg_b200 import synth, ops
from isg_b200.isubgvqa import MGAT
C, B, seed = 300, 5, 102
b = synth.make_batch(B, channels=C, mean_nodes=8, mean_edges=40, seed=seed)
sd = synth.make_state_dict(C, 4, 4, seed)
noise = util.case_noise("imle", B, b["nmax"], seed)
om = O.OracleMGAT(channels=C, sampler_type="imle", sample_k=2); om.load_state_dict(sd); om.eval(); om.debug_tensors = {}
x = b["x"].clone().requires_grad_(True); ea = b["edge_attr"].clone().requires_grad_(True)
h, mask, aux = om(x, b["edge_index"], b["instr_vectors"], b["global_language_feats"], ea, b["batch"], noise=noise, return_aux=True)
for t in aux["conv_out"]: t.retain_grad()
util.loss_fn(h).backward()
cm = MGAT(channels=C, num_ins=4, heads=4, use_instr=True, masking_thresholds=[1.0,1.0,1.0,0.1], use_topk=True, interpretable_mode=False, sampler_type="imle", sample_k=2)
cm.load_state_dict(sd); cm.cuda(); cm.eval()
cm.convs[3].mask.injected_noise = noise.cuda()
ops._DEBUG_EDGE_BWD = []
xc = b["x"].cuda().requires_grad_(True); eac = b["edge_attr"].cuda().requires_grad_(True)
hc, maskc, _, _ = cm(xc, b["edge_index"].cuda(), b["instr_vectors"].cuda(), b["global_language_feats"].cuda(), eac, b["batch"].cuda())
util.loss_fn(hc).backward()
rec = ops._DEBUG_EDGE_BWD[2]  # layer 1
N = b["x"].shape[0]; ei = b["edge_index"]
go_c = rec["g_out"].cpu(); go_o = aux["conv_out"][1].grad
print("g_out per-node max err:", [f"{float(e):.1e}" for e in (go_c - go_o).abs().amax(dim=1)])
print("g_out per-node max ref:", [f"{float(e):.1e}" for e in go_o.abs().amax(dim=1)])
for name in ("x_l", "x_r", "e_proj"):
    vc = rec[name].cpu(); vo = om.debug_tensors[f"{name}.1"].detach().reshape(vc.shape)
    print(name, "input val err", util.rel_err(vc, vo))
# oracle edge function on ORACLE inputs but CUDA g_out, and vice versa
def edge_bwd(xl, xr, ep, att, bias, gout, dtype=torch.float64):
    xl, xr, ep, att, bias = [t.detach().clone().to(dtype).requires_grad_(True) for t in (xl, xr, ep, att, bias)]
    out, alpha = O.gat_edge(xl.view(N,4,C), xr.view(N,4,C), ep.view(-1,4,C), att, ei, None)
    (out.reshape(N,4*C) + bias).backward(gout.to(dtype))
    return xl.grad, xr.grad, ep.grad
o_in = [om.debug_tensors[f"{n}.1"].reshape(-1, 4*C) for n in ("x_l", "x_r", "e_proj")] + [om.p("convs.1.att"), om.p("convs.1.bias")]
c_in = [rec[n].cpu() for n in ("x_l", "x_r", "e_proj", "att", "bias")]
ref_xr = om.debug_tensors["x_r.1"].grad.reshape(N, 4*C)
for label, ins, g in (("oracle-in/oracle-g", o_in, go_o), ("oracle-in/cuda-g", o_in, go_c), ("cuda-in/oracle-g", c_in, go_o), ("cuda-in/cuda-g", c_in, go_c)):
    gxl, gxr, gep = edge_bwd(*ins, g)
    print(f"{label:22s} g_xr vs oracle-chain x_r.grad: {util.rel_err(gxr, ref_xr):.2e}   node29 err {float((gxr[29].float()-ref_xr[29]).abs().max()):.2e}")
print("cuda kernel g_xr vs oracle-chain:", util.rel_err(rec["g_xr"], ref_xr))
