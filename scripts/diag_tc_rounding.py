"""Which way does the tcgen05 accumulator round?  Single-pass TF32 (mode 2) on tf32-exact inputs whose
products are exact: any error is accumulate rounding.  Prints signed mean error for all-positive and
all-negative sums as a function of the reduction length."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import isg_b200  # noqa
from isg_b200 import ops

ops.set_gemm_mode(2)
g = torch.Generator().manual_seed(0)
for K in (32, 304, 1216, 4864):
    for sign in (1.0, -1.0):
        # 11-bit mantissa values in [1,2): exact in tf32; products need 22 bits: exact in fp32
        x = (1.0 + torch.randint(0, 1024, (256, K), generator=g).float() / 1024.0).cuda()
        w = sign * (1.0 + torch.randint(0, 1024, (128, K), generator=g).float() / 1024.0).cuda()
        y, _ = ops.linear_fwd_raw(x, w, None, 0, False)
        ref = x.double() @ w.double().t()
        err = (y.double() - ref)
        ulp = torch.finfo(torch.float32).eps * ref.abs().mean()
        print(f"K={K:5d} sign={sign:+.0f} mean signed err = {float(err.mean() / ulp):+9.3f} ulp(result)  "
              f"max |err| = {float(err.abs().max() / ulp):8.3f} ulp   steps={K // 8}")
    # random-sign data (the realistic case)
    x = torch.randn(256, K, generator=g).cuda()
    w = torch.randn(128, K, generator=g).cuda()
    x = (x.view(torch.int32) & -8192).view(torch.float32)
    w = (w.view(torch.int32) & -8192).view(torch.float32)
    y, _ = ops.linear_fwd_raw(x, w, None, 0, False)
    ref = x.double() @ w.double().t()
    err = y.double() - ref
    print(f"K={K:5d} random   mean(err*sign(ref)) = {float((err * ref.sign()).mean() / ref.abs().mean()):+.3e} rel   "
          f"rms err = {float(err.pow(2).mean().sqrt() / ref.abs().mean()):.3e} rel")
