"""Bring-up diagnostic for the tcgen05 projections (csrc/linear_tc.cu): runs fwd / dgrad / wgrad in
modes 1 (3xTF32) and 2 (1xTF32) against an fp64 reference and the mode-0 FFMA kernel, each product in
its own subprocess (a trapped kernel kills only that process).  Usage: python scripts/diag_tc_gemm.py"""
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = [(128, 32, 128, 0), (128, 300, 1200, 0), (37, 300, 1200, 1), (515, 1200, 600, 1), (130, 600, 300, 1),
          (9600, 300, 1200, 0), (64, 8, 32, 1), (257, 300, 300, 1), (39809, 300, 1200, 0)]


def rel(a, b):
    b = b.to(a.device, torch.float64)
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))


def child(which, mode):
    import torch
    globals()["torch"] = torch
    import isg_b200  # noqa: F401
    from isg_b200 import lib as L
    from isg_b200 import ops

    dev = "cuda"
    for (M, K, Nout, act) in SHAPES:
        g = torch.Generator().manual_seed(M + K)
        x = torch.randn(M, K, generator=g)
        w = torch.randn(Nout, K, generator=g) / math.sqrt(K)
        b = torch.randn(Nout, generator=g) * 0.1
        gy = torch.randn(M, Nout, generator=g)
        xd, wd, bd, gyd = (t.to(dev) for t in (x, w, b, gy))
        x64, w64, b64, gy64 = (t.double() for t in (xd, wd, bd, gyd))
        ops.set_gemm_mode(mode)
        t0 = time.time()
        if which == "fwd":
            y, z = ops.linear_fwd_raw(xd, wd, bd, L.ACT_GELU if act else L.ACT_NONE, want_pre=bool(act))
            torch.cuda.synchronize()
            pre = x64 @ w64.t() + b64
            ref = torch.nn.functional.gelu(pre) if act else pre
            errs = {"y": rel(y, ref)}
            if act:
                errs["z"] = rel(z, pre)
        elif which == "dgrad":
            zprev = torch.randn(M, K, generator=g).to(dev) if act else None
            gx = ops.linear_dgrad_raw(gyd, wd, zprev)
            torch.cuda.synchronize()
            ref = gy64 @ w64
            if act:
                zz = zprev.double().requires_grad_(True)
                torch.nn.functional.gelu(zz).sum().backward()
                ref = ref * zz.grad
            errs = {"gx": rel(gx, ref)}
        else:
            gw = ops.linear_wgrad_raw(gyd, xd)
            torch.cuda.synchronize()
            errs = {"gw": rel(gw, gy64.t() @ x64)}
        # timing (10 reps)
        fn = {"fwd": lambda: ops.linear_fwd_raw(xd, wd, bd, 0, False), "dgrad": lambda: ops.linear_dgrad_raw(gyd, wd),
              "wgrad": lambda: ops.linear_wgrad_raw(gyd, xd)}[which]
        for _ in range(3):
            fn()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        e.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(e) / 10
        tf = 2.0 * M * K * Nout / (ms * 1e-3) / 1e12
        print(f"{which:5s} mode={mode} M={M:6d} K={K:5d} Nout={Nout:5d} act={act} "
              + " ".join(f"{k}={v:.2e}" for k, v in errs.items()) + f"  {ms:.3f} ms  {tf:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    if len(sys.argv) == 3:
        child(sys.argv[1], int(sys.argv[2]))
        sys.exit(0)
    rc = 0
    for mode in [int(m) for m in os.environ.get("DIAG_MODES", "1,2").split(",")]:
        for which in ("fwd", "dgrad", "wgrad"):
            try:
                p = subprocess.run([sys.executable, __file__, which, str(mode)], timeout=120)
                if p.returncode != 0:
                    print(f"!! {which} mode={mode} exited with {p.returncode}", flush=True)
                    rc = 1
            except subprocess.TimeoutExpired:
                print(f"!! {which} mode={mode} timed out", flush=True)
                rc = 1
    sys.exit(rc)
