import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch, util
import isg_oracle as O
from isg_b200 import synth
from isg_b200.isubgvqa import MGAT
cfg = dict(sampler="imle", train=False, channels=300, num_graphs=5, mean_nodes=8, mean_edges=40, k=2, seed=102, steps=1)
C, B, seed = 300, 5, 102
b = synth.make_batch(B, channels=C, mean_nodes=8, mean_edges=40, seed=seed)
sd = synth.make_state_dict(C, 4, 4, seed)
noise = util.case_noise("imle", B, b["nmax"], seed)
om = O.OracleMGAT(channels=C, sampler_type="imle", sample_k=2); om.load_state_dict(sd); om.eval(); om.debug_tensors = {}
x = b["x"].clone().requires_grad_(True); ea = b["edge_attr"].clone().requires_grad_(True)
h, mask, _, _ = om(x, b["edge_index"], b["instr_vectors"], b["global_language_feats"], ea, b["batch"], noise=noise)
util.loss_fn(h).backward()
cm = MGAT(channels=C, num_ins=4, heads=4, use_instr=True, masking_thresholds=[1.0,1.0,1.0,0.1], use_topk=True, interpretable_mode=False, sampler_type="imle", sample_k=2)
cm.load_state_dict(sd); cm.cuda(); cm.eval()
for c in cm.convs: c.debug_tensors = {}
cm.convs[3].mask.injected_noise = noise.cuda()
xc = b["x"].cuda().requires_grad_(True); eac = b["edge_attr"].cuda().requires_grad_(True)
hc, maskc, _, _ = cm(xc, b["edge_index"].cuda(), b["instr_vectors"].cuda(), b["global_language_feats"].cuda(), eac, b["batch"].cuda())
util.loss_fn(hc).backward()
for i in range(4):
    for name in ("xg", "x_l", "x_r", "e_proj"):
        tc = cm.convs[i].debug_tensors[name]; to = om.debug_tensors[f"{name}.{i}"].reshape(tc.shape)
        gc = tc.grad; go = om.debug_tensors[f"{name}.{i}"].grad.reshape(tc.shape)
        print(f"layer {i} {name:7s} val {util.rel_err(tc, to):.2e} grad {util.rel_err(gc, go):.2e} contiguous {tc.is_contiguous()} gradcontig {gc.is_contiguous()} gstride {gc.stride()}")
    cp = dict(cm.named_parameters())
    for pn in (f"convs.{i}.lin_l.weight", f"convs.{i}.lin_r.weight", f"convs.{i}.lin_edge.weight", f"convs.{i}.lin_l.bias", f"convs.{i}.lin_r.bias"):
        print(f"      {pn:28s} {util.rel_err(cp[pn].grad, om.p(pn).grad):.2e}")
    # recompute param grads from the CUDA tensors with torch on the GPU (float64)
    d = cm.convs[i].debug_tensors
    gwl = d["x_l"].grad.double().t() @ d["xg"].double(); gwr = d["x_r"].grad.double().t() @ d["xg"].double()
    print(f"      torch-recomputed from cuda tensors: lin_l.w {util.rel_err(cp[f'convs.{i}.lin_l.weight'].grad, gwl):.2e} lin_r.w {util.rel_err(cp[f'convs.{i}.lin_r.weight'].grad, gwr):.2e}  vs oracle: {util.rel_err(gwl, om.p(f'convs.{i}.lin_l.weight').grad):.2e} {util.rel_err(gwr, om.p(f'convs.{i}.lin_r.weight').grad):.2e}")
print("---- extra: where is layer-1 x_r.grad wrong?")
d = cm.convs[1].debug_tensors
gc = d["x_r"].grad.cpu(); go = om.debug_tensors["x_r.1"].grad.reshape(gc.shape)
err = (gc - go).abs().amax(dim=1); ref = go.abs().amax(dim=1)
print("per-node err:", [f"{float(e):.1e}" for e in err])
print("per-node ref:", [f"{float(e):.1e}" for e in ref])
print("batch:", b["batch"].tolist())
errh = (gc - go).abs().view(-1, 4, 300).amax(dim=2)
print("per-node-head err (first 12 nodes):", errh[:12])
from isg_b200.graph import get_graph_index
gi = get_graph_index(b["edge_index"].cuda(), b["batch"].cuda(), B)
want = O.csr_build(b["edge_index"], b["x"].shape[0])
for key in ("dst_ptr", "dst_eid", "dst_nbr", "src_ptr", "src_eid", "src_nbr"):
    print(key, "intact:", bool(torch.equal(getattr(gi, key).cpu(), want[key])))
