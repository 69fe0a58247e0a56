"""Per-shape table of the dense projections of one c3 training step: every isg_linear_{fwd,dgrad,wgrad} call is
timed with CUDA events (eager, synchronised after each call — absolute times include no overlap), grouped by
(product, M, Nout, K, mode).  Usage: python scripts/profile_gemm_shapes.py"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import isg_b200  # noqa: E402,F401
from isg_b200 import lib as L  # noqa: E402
from isg_b200 import ops, synth  # noqa: E402
from isg_b200.isubgvqa.mgat import MGAT  # noqa: E402

dev = torch.device("cuda")
B = 256
b = synth.make_batch(B, seed=3407)
model = MGAT(channels=300, num_ins=4, heads=4, use_instr=True, masking_thresholds=[1.0, 1.0, 1.0, 0.1], use_topk=True,
             interpretable_mode=False, sampler_type="aimle", sample_k=2, nb_samples=1, alpha=1.0, beta=10.0, tau=1.0)
model.load_state_dict(synth.make_state_dict(300, 4, 4, 3407))
model.to(dev).train(True)
model.convs[3].mask.sampler_train.target._init[0] = 1.0
t = {k: b[k].to(dev) for k in ("x", "edge_index", "instr_vectors", "global_language_feats", "edge_attr", "batch")}
ops.allow_side_stream(False)

records = collections.defaultdict(list)
orig_call = L.call
recording = [False]


def timed_call(name, *args):
    if not recording[0] or not name.startswith("isg_linear_"):
        return orig_call(name, *args)
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = orig_call(name, *args)
    e.record()
    torch.cuda.synchronize()
    if name == "isg_linear_fwd":
        M, Nout, K, mode = args[11], args[12], args[13], args[15]
    elif name == "isg_linear_dgrad":
        M, Nout, K, mode = args[9], args[10], args[11], args[12]
    elif name == "isg_linear_wgrad":
        M, Nout, K, mode = args[6], args[7], args[8], args[9]
    else:
        return r
    records[(name[11:], M, Nout, K, mode)].append(a.elapsed_time(e))
    return r


L.call = timed_call
ops.L.call = timed_call


def step():
    x = t["x"].detach().requires_grad_(True)
    ea = t["edge_attr"].detach().requires_grad_(True)
    for p in model.parameters():
        p.grad = None
    h, mask, _, _ = model(x, t["edge_index"], t["instr_vectors"], t["global_language_feats"], ea, t["batch"],
                          return_masks=True)
    (h * h).mean().backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
recording[0] = True
STEPS = 5
for _ in range(STEPS):
    step()
rows = []
for (prod, M, Nout, K, mode), ts in records.items():
    ts.sort()
    med = ts[len(ts) // 2]
    n = len(ts) / STEPS
    fl = 2.0 * M * Nout * K
    rows.append((med * n, prod, M, Nout, K, mode, n, med, fl / (med * 1e-3) / 1e12))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"{'product':6s} {'M':>6s} {'Nout':>5s} {'K':>5s} mode calls/step  med_us  TFLOP/s  ms/step  share")
for ms, prod, M, Nout, K, mode, n, med, tf in rows:
    print(f"{prod:6s} {M:6d} {Nout:5d} {K:5d} {mode:4d} {n:10.1f} {med * 1e3:7.1f} {tf:8.1f} {ms:8.3f} {100 * ms / tot:5.1f}%")
print(f"total {tot:.3f} ms/step over {sum(r[6] for r in rows):.0f} calls")
