"""One large projection (fwd / dgrad / wgrad of [39809,300]x[300,1200]) in the current GEMM mode, a few
repetitions; a quick runner for ncu captures and A/B timing of the tcgen05 kernel."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import isg_b200  # noqa: E402,F401
from isg_b200 import ops  # noqa: E402

M, K, N = 39809, 300, 1200
g = torch.Generator().manual_seed(1)
x = torch.randn(M, K, generator=g).cuda()
w = (torch.randn(N, K, generator=g) / math.sqrt(K)).cuda()
b = torch.zeros(N).cuda()
gy = torch.randn(M, N, generator=g).cuda()
ops.set_gemm_mode(int(os.environ.get("MODE", "1")))
for name, fn in (("fwd", lambda: ops.linear_fwd_raw(x, w, b, 0, False)), ("dgrad", lambda: ops.linear_dgrad_raw(gy, w)),
                 ("wgrad", lambda: ops.linear_wgrad_raw(gy, x))):
    print("==", name, file=sys.stderr, flush=True)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
