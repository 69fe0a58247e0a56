"""Times the three products of a projection (isg_linear_fwd / _dgrad / _wgrad) on the shapes of one c3 layer, fp32-grade
mode 1 by default (ISG_TC_PAIR2=0 disables the cta_group::2 pairs; --bf16 times the kind::f16 products instead).
Median of 20 launches, L2 flushed between launches.  Prints one line per shape."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import isg_b200  # noqa: E402,F401
from isg_b200 import lib as L  # noqa: E402

dev = torch.device("cuda")
lib = L.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
bf16 = "--bf16" in sys.argv
mode = 1
SHAPES = [("lin_edge", 39809, 300, 1200), ("lin_l|lin_r", 4910, 300, 2400), ("x_proj[0]", 4910, 1200, 600),
          ("x_proj[2]", 4910, 600, 300)]


def med(fn, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


for name, M, K, Nout in SHAPES:
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(M, K, device=dev, generator=g)
    w = torch.randn(Nout, K, device=dev, generator=g) / K ** 0.5
    gy = torch.randn(M, Nout, device=dev, generator=g)
    st = L.stream()
    fl = 2.0 * M * K * Nout
    if not bf16:
        y, gx, gw = torch.empty(M, Nout, device=dev), torch.empty(M, K, device=dev), torch.empty(Nout, K, device=dev)
        nb = lib.isg_linear_wgrad_workspace_bytes(M, Nout, K)
        ws = L.workspace(nb, dev)
        t_f = med(lambda: L.call("isg_linear_fwd", x.data_ptr(), K, w.data_ptr(), None, None, None, y.data_ptr(), Nout, None,
                                 0, M, Nout, K, 0, mode, 0, st))
        t_d = med(lambda: L.call("isg_linear_dgrad", gy.data_ptr(), Nout, w.data_ptr(), None, None, 0, gx.data_ptr(), K, 0, M,
                                 Nout, K, mode, 0, st))
        t_w = med(lambda: L.call("isg_linear_wgrad", gy.data_ptr(), Nout, x.data_ptr(), K, gw.data_ptr(), None, M, Nout, K,
                                 mode, 0, ws.data_ptr(), nb, st))
    else:
        p8 = lambda n: (n + 7) // 8 * 8
        xb = torch.zeros(M, p8(K), dtype=torch.bfloat16, device=dev)
        xb[:, :K] = x
        wb = torch.zeros(Nout, p8(K), dtype=torch.bfloat16, device=dev)
        wb[:, :K] = w
        wt = torch.zeros(K, p8(Nout), dtype=torch.bfloat16, device=dev)
        wt[:, :Nout] = w.t()
        gyb = gy.to(torch.bfloat16)
        y = torch.empty(M, p8(Nout), dtype=torch.bfloat16, device=dev)
        gx = torch.empty(M, K, device=dev)
        gw = torch.empty(Nout, K, device=dev)
        nb = lib.isg_linear_bf16_wgrad_workspace_bytes(M, Nout, K)
        ws = L.workspace(nb, dev)
        t_f = med(lambda: L.call("isg_linear_bf16_fwd", xb.data_ptr(), p8(K), wb.data_ptr(), p8(K), None, y.data_ptr(), p8(Nout),
                                 None, 0, M, Nout, K, 0, 1, st))
        gyp = torch.zeros(M, p8(Nout), dtype=torch.bfloat16, device=dev)
        gyp[:, :Nout] = gyb
        t_d = med(lambda: L.call("isg_linear_bf16_dgrad", gyp.data_ptr(), p8(Nout), wt.data_ptr(), p8(Nout), None, 0,
                                 gx.data_ptr(), K, 0, M, Nout, K, 0, st))
        t_w = med(lambda: L.call("isg_linear_bf16_wgrad", gyp.data_ptr(), p8(Nout), xb.data_ptr(), p8(K), gw.data_ptr(), M, Nout,
                                 K, ws.data_ptr(), nb, st))
    print(f"{'bf16' if bf16 else 'mode1'} {name:12s} M={M:6d} K={K:5d} Nout={Nout:5d}  fwd {t_f*1e3:7.1f} us {fl/t_f/1e9:7.1f} TF/s | "
          f"dgrad {t_d*1e3:7.1f} us {fl/t_d/1e9:7.1f} | wgrad {t_w*1e3:7.1f} us {fl/t_w/1e9:7.1f}", flush=True)
