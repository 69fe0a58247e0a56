"""Measurement of the SURVEY.md section 8(f) rows at the c3 size (256 graphs, N ~ 4.9 K nodes, E ~ 39.8 K edges):
  f1  GlobalAttention pooling (models/att_pooling.py:57-77)           fwd+bwd
  f2  SceneGraphEncoder MetaLayer + float64 GraphNorm (:91-146)       fwd+bwd, next to the CPU oracle port
  f3  batch construction: collate with / without the per-image CSR cache, device build vs upload
  f4  FusedClipAdam (training/train_epoch.py:111-118) vs GradScaler.unscale_ + clip_grad_norm_ + torch Adam on the GPU
CUDA events, median of `reps`, L2 flushed between launches.  One JSON line per row."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import isg_b200  # noqa: E402,F401
from isg_b200 import collate, synth  # noqa: E402
from isg_b200.graph import GraphIndex, clear_cache  # noqa: E402

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
B, C = 256, 300
b = synth.make_batch(B, channels=C, mean_nodes=20, mean_edges=150, seed=3407)
N, E = b["x"].shape[0], b["edge_index"].shape[1]
peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def med(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def out(row, **kw):
    print(json.dumps(dict(row=row, graphs=B, N=N, E=E, **kw)), flush=True)


ei, batch = b["edge_index"].to(dev), b["batch"].to(dev)
x, ea = b["x"].to(dev), b["edge_attr"].to(dev)

# ---------------------------------------------------------------- f1
from isg_b200.isubgvqa import GlobalAttention  # noqa: E402

pool = GlobalAttention(C, C).to(dev)
u = b["global_language_feats"].to(dev)
mask = (torch.rand(N, 1, device=dev) > 0.5).float()


def f1():
    xx = x.detach().requires_grad_(True)
    o = pool(xx, u, batch, node_mask=mask)
    (o[0] if isinstance(o, tuple) else o).square().mean().backward()


t = med(f1)
out("f1 GlobalAttention fwd+bwd (2 projections + fused masked softmax-pool kernel)", ms=round(t, 4))

# ---------------------------------------------------------------- f2
from isg_b200.isubgvqa import GraphNorm64, SceneGraphEncodingLayer, encode_scene_graph  # noqa: E402

sd, gnp = synth.make_sgenc_state_dict(C, C, C, 3407)
layer, gn = SceneGraphEncodingLayer(C, C, C), GraphNorm64(C)
layer.load_state_dict(sd)
gn.load_state_dict(gnp)
layer.to(dev), gn.to(dev)


def f2():
    xx, ee = x.detach().requires_grad_(True), ea.detach().requires_grad_(True)
    xn, en = encode_scene_graph(layer, gn, xx, ei, ee, batch, B)
    (xn.square().mean() + en.square().mean()).backward()


t = med(f2)
flops = 3 * 2 * (E * (900 * 300 + 300 * 300) + 2 * N * (600 * 300 + 300 * 300))  # fwd + dgrad + wgrad of the six Linear layers
import isg_oracle as O  # noqa: E402

p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
g = {k: v.clone().requires_grad_(True) for k, v in gnp.items()}
torch.set_num_threads(os.cpu_count() or 1)
t0 = time.perf_counter()
xo, eo = b["x"].clone().requires_grad_(True), b["edge_attr"].clone().requires_grad_(True)
xn, en = O.scene_graph_encode(xo, b["edge_index"], eo, b["batch"], p, g["weight"], g["bias"], g["mean_scale"], B)
(xn.square().mean() + en.square().mean()).backward()
cpu_ms = 1e3 * (time.perf_counter() - t0)
out("f2 SceneGraphEncoder layer fwd+bwd (MetaLayer edge / node MLPs + float64 GraphNorm on the device)", ms=round(t, 4),
    graphs_per_s=round(B / (t * 1e-3)), tflops_fp32_equiv=round(flops / (t * 1e-3) / 1e12, 1),
    cpu_oracle_ms=round(cpu_ms, 1), cpu_cores=os.cpu_count())

# ---------------------------------------------------------------- f3
gs = []
for gidx in range(B):
    nodes = (b["batch"] == gidx).nonzero().flatten()
    n0 = int(nodes[0])
    keep = b["batch"][b["edge_index"][0]] == gidx
    gs.append(dict(x=b["x"][nodes], edge_index=b["edge_index"][:, keep] - n0, edge_attr=b["edge_attr"][keep], image_id=gidx))
cache = collate.SceneGraphCsrCache()
t0 = time.perf_counter()
collate.collate_scene_graphs(gs, cache)
cold_ms = 1e3 * (time.perf_counter() - t0)
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    hb = collate.collate_scene_graphs(gs, cache, pin=True)
    ts.append(1e3 * (time.perf_counter() - t0))
warm_ms = sorted(ts)[2]
t_build = med(lambda: GraphIndex(ei, batch, B))
t_upload = med(lambda: GraphIndex.from_host(ei, batch, hb["host_index"]))
out("f3 batch construction: collate_scene_graphs (host) and the index on the device", collate_cold_ms=round(cold_ms, 2),
    collate_cached_ms=round(warm_ms, 2), device_csr_build_ms=round(t_build, 4), host_csr_upload_ms=round(t_upload, 4))

# ---------------------------------------------------------------- f4
from isg_b200.isubgvqa import MGAT  # noqa: E402
from isg_b200.optim import FusedClipAdam  # noqa: E402

model = MGAT(channels=C, num_ins=4, heads=4, use_instr=True, masking_thresholds=[1.0, 1.0, 1.0, 0.1], use_topk=True,
             interpretable_mode=False, sampler_type="aimle", sample_k=2).to(dev)
params = [p_ for p_ in model.parameters()]
for p_ in params:
    p_.grad = torch.randn_like(p_) * 1e-3
nparam = sum(p_.numel() for p_ in params)
fused = FusedClipAdam(params, lr=1e-4)
scaler = torch.amp.GradScaler("cuda", enabled=False)
t_fused = med(lambda: fused.step())
ref_opt = torch.optim.Adam(params, lr=1e-4)


def torch_step():
    torch.nn.utils.clip_grad_norm_(params, max_norm=2.0)
    ref_opt.step()


t_torch = med(torch_step)
out("f4 optimizer tail: FusedClipAdam.step (3 launches, no host sync) vs clip_grad_norm_ + torch.optim.Adam.step (GPU, foreach)",
    parameters=nparam, fused_ms=round(t_fused, 4), torch_ms=round(t_torch, 4),
    fused_GBps=round(16.0 * nparam / (t_fused * 1e-3) / 1e9, 1), frac_of_hbm_peak=round(16.0 * nparam / (t_fused * 1e-3) / 1e9 / peak, 3))
