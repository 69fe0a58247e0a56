"""Data-parallel plumbing for the hot path: one process per GPU, whole graphs sharded across ranks,
no collective inside the path (every operator is block-diagonal over graphs — SURVEY.md §8e); the only
exchange is the gradient all-reduce (mean) over parameters, which the reference gets from
DistributedDataParallel(find_unused_parameters=True) (main.py:85-94).

`GradAllReduce` keeps ONE flat fp32 bucket for all parameters of the module (≈42 MB for MGAT) so the
collective is a single NCCL call over NVLink/NVSwitch (NVLS-eligible), issued on a side stream so it can
overlap whatever the caller runs next.  Parameters that received no gradient on this rank (36 MGAT
tensors never do: unused gate_nn / gate_top / node_logits) contribute zeros, exactly like DDP's
unused-parameter handling, so every rank reduces an identically laid-out bucket."""
import torch
import torch.distributed as dist


class GradAllReduce:
    def __init__(self, module, process_group=None):
        from . import ops

        ops.set_side_stream_with_dist(True)  # gradients are read after backward() has returned (and joined)
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self.nbytes = self.numel * 4

    def pack(self):
        """grads -> flat bucket (zeros where a parameter has no gradient)."""
        have = [(v, p.grad) for v, p in zip(self.views, self.params) if p.grad is not None]
        none = [v for v, p in zip(self.views, self.params) if p.grad is None]
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        for v in none:
            v.zero_()

    def unpack(self):
        """flat bucket -> .grad of every parameter that had one (others stay None, like DDP)."""
        have = [(v, p.grad) for v, p in zip(self.views, self.params) if p.grad is not None]
        if have:
            torch._foreach_copy_([g for _, g in have], [v for v, _ in have])

    def all_reduce_mean(self, async_op=False):
        """Averages gradients across ranks.  With async_op=True the collective runs on a side stream;
        call wait() before reading the gradients."""
        self.pack()
        if self.world == 1:
            return
        if self.stream is not None and async_op:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
                self.flat.mul_(1.0 / self.world)
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.mul_(1.0 / self.world)
        self.unpack()

    def wait(self):
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        self.unpack()


def shard_graphs(num_graphs_total, rank, world):
    """Contiguous, balanced split of whole graphs across ranks (DistributedSampler analogue,
    datasets/build.py:44-49, without shuffling): returns (first_graph, num_graphs) of this rank."""
    base, rem = divmod(num_graphs_total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)
