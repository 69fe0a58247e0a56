"""Prints, per golden case and projection mode (0 FFMA / 1 tcgen05 3xTF32), the worst relative error of
the CUDA path against the reference fixture, and (mode-independent) the reference fp32 oracle's own
distance from an fp64 replay — the yardstick for ill-conditioned tensors (Gumbel mask-net gradients).
Usage (GPU box): python scripts/diag_mode_margins.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import torch  # noqa: E402

import isg_b200  # noqa: E402,F401
import util  # noqa: E402
from isg_b200 import ops  # noqa: E402


def worst(got, want):
    w = ("", 0.0)
    for key in ("h", "gx", "g_edge_attr", "g_instr", "g_glf"):
        e = util.rel_err(got[key], want[key])
        if e > w[1]:
            w = (key, e)
    for name, ref in want["param_grads"].items():
        g = got["param_grads"].get(name)
        if ref is None or g is None or isinstance(ref, dict):
            continue
        e = util.rel_err(g, ref)
        if e > w[1]:
            w = (name, e)
    return w


for path in util.golden_files():
    fix = util.load_golden(path)
    cfg = fix["config"]
    name = os.path.basename(path)[:-3]
    o32, o64 = util.run_oracle_fp64_arbiter(cfg)
    ref_self = worst(o32[0], o64[0])
    line = f"{name:22s} ref32-vs-fp64 worst {ref_self[1]:.2e} ({ref_self[0]})"
    for mode in (0, 1):
        ops.set_gemm_mode(mode)
        outs = util.run_cuda_case(cfg)
        w = ("", 0.0)
        wa = ("", 0.0)
        for got, want, exact in zip(outs, fix["steps"], o64):
            c = worst(got, want)
            if c[1] > w[1]:
                w = c
            a = worst(got, exact)
            if a[1] > wa[1]:
                wa = a
        line += f" | mode{mode} vs fixture {w[1]:.2e} ({w[0]}) vs fp64 {wa[1]:.2e} ({wa[0]})"
    print(line, flush=True)
