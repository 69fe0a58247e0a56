"""Per-tensor errors of the CUDA path against the teacher-forced fp32 oracle and its fp64 replay for one of the
BASELINE-size test configs, per projection mode.  Usage: python scripts/diag_case.py gumbel 0 96"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch
import isg_b200  # noqa
import util
from isg_b200 import ops

sampler, train, B = sys.argv[1], bool(int(sys.argv[2])), int(sys.argv[3])
cfg = dict(sampler=sampler, train=train, channels=300, num_graphs=B, mean_nodes=20, mean_edges=150, k=2,
           seed=900 + B, steps=1, aimle_beta0=2.0 if sampler == "aimle" else None)
free = util.run_oracle_case(cfg, record=True)[0]
for mode in (0, 1):
    ops.set_gemm_mode(mode)
    got = util.run_cuda_case(cfg, capture=True)[0]
    want = util.run_oracle_case(cfg, teacher=[got["teacher"]])[0]
    exact = util.run_oracle_case(cfg, dtype=torch.float64, replay=[free["record"]], teacher=[got["teacher"]])[0]
    rows = []
    for name, ref in exact["param_grads"].items():
        g, w = got["param_grads"].get(name), want["param_grads"].get(name)
        if ref is None or g is None:
            continue
        rows.append((util.rel_err(g, ref), util.rel_err(w, ref), name))
    for key in ("h", "gx", "g_edge_attr", "g_instr", "g_glf"):
        rows.append((util.rel_err(got[key], exact[key]), util.rel_err(want[key], exact[key]), key))
    rows.sort(reverse=True)
    print(f"--- mode {mode}: (cuda vs fp64, oracle32 vs fp64, tensor), worst 10")
    for r in rows[:10]:
        print("  %.2e  %.2e  %s" % r)
