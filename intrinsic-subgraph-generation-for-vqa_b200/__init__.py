"""isg_b200 — B200-native (sm_100a) hot path for ISubGVQA: question-conditioned GATv2 message
passing over batched GQA scene graphs + discrete top-k subgraph samplers (IMLE / AIMLE /
Gumbel / SIMPLE), behind the reference's own nn.Module surface (ISubGVQA/models/mgat.py,
mgat_v2_conv.py, masking.py, ISubGVQA/sampling/**).  Host code is Python/PyTorch plumbing
(device memory, streams, torch.distributed); all arithmetic on the path runs in
hand-written CUDA kernels reached through the C ABI in include/isg.h (csrc/ -> libisg.so).
There is no CPU fallback: importing `isg_b200.lib` raises if libisg.so is missing."""
__version__ = "0.1.0"
