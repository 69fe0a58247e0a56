"""Host-side (Python) cost of one eager MGAT training step: cProfile over 20 steps, top functions by cumulative
and own time.  Usage (GPU box): python scripts/profile_host.py"""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import isg_b200  # noqa
from isg_b200 import synth
from isg_b200.isubgvqa import MGAT

dev = torch.device("cuda")
b = synth.make_batch(256, seed=3407)
model = MGAT(channels=300, num_ins=4, heads=4, use_instr=True, masking_thresholds=[1.0, 1.0, 1.0, 0.1], use_topk=True,
             interpretable_mode=False, sampler_type="aimle", sample_k=2).to(dev).train()
model.load_state_dict(synth.make_state_dict(300, 4, 4, 3407))
model.convs[3].mask.sampler_train.target._init[0] = 1.0
t = {k: b[k].to(dev) for k in ("x", "edge_index", "instr_vectors", "global_language_feats", "edge_attr", "batch")}
noise = synth.gumbel_noise(256, b["nmax"], 0.3, 3407).to(dev)

def step():
    model.convs[3].mask.injected_noise = noise
    x = t["x"].detach().requires_grad_(True)
    ea = t["edge_attr"].detach().requires_grad_(True)
    for p in model.parameters():
        p.grad = None
    h, mask, _, _ = model(x, t["edge_index"], t["instr_vectors"], t["global_language_feats"], ea, t["batch"], return_masks=True)
    (h * h).mean().backward()

for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host issue time per step: {(t1 - t0) / 20 * 1e3:.2f} ms; incl. final sync: {(t2 - t0) / 20 * 1e3:.2f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
pr.disable()
torch.cuda.synchronize()
for key in ("cumulative", "tottime"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(22)
    print("\n".join(l[:150] for l in s.getvalue().splitlines()[4:40]))
