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
    def __init__(self, module, process_group=None, always_exchange=False, params=None):
        from . import ops

        ops.set_side_stream_with_dist(True)  # gradients are read after backward() has returned (and joined)
        self.params = [p for p in (module.parameters() if params is None else params) if p.requires_grad]
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self.nbytes = self.numel * 4
        self._none_seen = None    # this rank's gradient-less parameter indices at the last used-set exchange
        self._global_used = None  # per parameter: did ANY rank produce a gradient (exchanged when the set changes)
        self.always_exchange = bool(always_exchange)
        # NCCL averages inside the collective; other backends (gloo in the CPU tests) sum and scale afterwards
        self._avg = dist.is_initialized() and dist.get_backend(process_group) == "nccl"

    def _local_unused(self):
        return tuple(i for i, p in enumerate(self.params) if p.grad is None)

    def _exchange_used_set(self, none):
        """DDP(find_unused_parameters=True) semantics need the GLOBAL picture: a parameter that got no gradient
        here but did on another rank must still receive the averaged gradient, and one that no rank used keeps
        .grad = None.  The used-bitmap is all-reduced (MAX) whenever this rank's gradient-less set changes — one
        small collective plus a host read on the first step and never again while the set is static.  Every rank
        must take this branch in the same step (true when usage is decided by configuration; data-dependent usage
        needs the exchange on every step: pass always_exchange=True)."""
        used = torch.ones(len(self.params), dtype=torch.float32, device=self.flat.device)
        if none:
            used[list(none)] = 0.0
        if self.world > 1:
            dist.all_reduce(used, op=dist.ReduceOp.MAX, group=self.group)
        self._global_used = tuple(bool(v) for v in used.cpu().tolist())
        self._none_seen = none

    def pack(self):
        """grads -> flat bucket.  A parameter without a gradient on this rank contributes zeros; its slice is
        re-zeroed on EVERY pack (after a reduce it holds the other ranks' average, not zeros)."""
        none = self._local_unused()
        if self.always_exchange or none != self._none_seen:
            self._exchange_used_set(none)
        have = [(v, p.grad) for v, p in zip(self.views, self.params) if p.grad is not None]
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        if none:
            torch._foreach_zero_([self.views[i] for i in none])

    def unpack(self):
        """flat bucket -> .grad.  Parameters some rank used get the averaged gradient (materialised here when this
        rank had none); parameters no rank used keep .grad = None, like DDP."""
        dst, src = [], []
        for i, (v, p) in enumerate(zip(self.views, self.params)):
            if p.grad is not None:
                dst.append(p.grad)
                src.append(v)
            elif self._global_used is not None and self._global_used[i]:
                p.grad = v.clone()
        if dst:
            torch._foreach_copy_(dst, src)

    def all_reduce_mean(self, async_op=False):
        """Averages gradients across ranks.  With async_op=True the collective runs on a side stream;
        call wait() before reading the gradients."""
        self.pack()
        if self.world == 1:
            return
        if self.stream is not None and async_op:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                self._reduce(self.flat)
            return
        self._reduce(self.flat)
        self.unpack()

    def _reduce(self, buf):
        if self._avg:
            dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            buf.mul_(1.0 / self.world)

    def wait(self):
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        self.unpack()


class OverlappedGradAllReduce(GradAllReduce):
    """Bucketed variant that overlaps the collective with the backward pass (what DDP's reducer does for the
    reference, main.py:85-94): the parameters are cut into buckets in the order their gradients are produced;
    a post-accumulate-grad hook packs each finished bucket into its slice of the flat buffer and issues its
    all-reduce on a communication stream while the backward pass of the earlier layers is still running.

    Protocol per step:   backward()  ->  finish()        (finish joins the communication stream and unpacks)
    The first backward is a RECORDING pass (nothing is overlapped): it observes which parameters receive a
    gradient and in which order; parameters that never do (36 MGAT tensors) are left out of the buckets and
    keep contributing zeros, like DDP's unused-parameter handling.  The set must not change afterwards.
    Works under CUDA-graph capture: the communication stream forks from / joins the capturing stream."""

    def __init__(self, module, process_group=None, bucket_bytes=8 << 20):
        super().__init__(module, process_group)
        from . import ops

        ops.forbid_side_stream(True)  # the hooks read gradients DURING backward: no side-stream producers
        self.bucket_bytes = int(bucket_bytes)
        self._index = {p: i for i, p in enumerate(self.params)}
        self._offsets, off = [], 0
        for p in self.params:
            self._offsets.append(off)
            off += p.numel()
        self._order = []        # recording pass: parameter indices in gradient-ready order
        self._buckets = None    # list of dicts: params (indices), lo, hi (flat range)
        self._bucket_of = {}
        self._pending = []
        self._launched = 0
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]

    # ---- recording pass -> buckets
    def _build_buckets(self):
        # the flat buffer is re-laid-out in gradient-ready order so that every bucket is one contiguous slice
        order = self._order + [i for i in range(len(self.params)) if i not in set(self._order)]
        self.views, self._offsets, off = [None] * len(self.params), [0] * len(self.params), 0
        for i in order:
            p = self.params[i]
            self._offsets[i] = off
            self.views[i] = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.flat.zero_()
        self._buckets, cur = [], None
        for i in self._order:
            nbytes = self.params[i].numel() * 4
            if cur is None or cur["bytes"] >= self.bucket_bytes:
                cur = {"params": [], "lo": self._offsets[i], "hi": self._offsets[i], "bytes": 0}
                self._buckets.append(cur)
            cur["params"].append(i)
            cur["hi"] = self._offsets[i] + self.params[i].numel()
            cur["bytes"] += nbytes
        self._bucket_of = {i: b for b, bk in enumerate(self._buckets) for i in bk["params"]}
        self._reset()

    def _reset(self):
        self._pending = [len(bk["params"]) for bk in self._buckets]
        self._launched = 0

    # ---- hooks
    def _on_grad(self, p):
        i = self._index[p]
        if self._buckets is None:
            self._order.append(i)
            return
        b = self._bucket_of.get(i)
        if b is None:
            raise RuntimeError("a parameter outside the recorded set received a gradient; "
                               "OverlappedGradAllReduce needs a static set of used parameters")
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._launch(b)

    def _launch(self, b):
        bk = self._buckets[b]
        torch._foreach_copy_([self.views[i] for i in bk["params"]], [self.params[i].grad for i in bk["params"]])
        self._launched += 1
        if self.world == 1:
            return
        chunk = self.flat[bk["lo"]:bk["hi"]]
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                self._reduce(chunk)
        else:
            self._reduce(chunk)

    def finish(self):
        """Call after backward(): completes the step's reduction and writes the averaged gradients back."""
        if self._buckets is None:  # recording pass: reduce everything now, unoverlapped, and build the buckets
            if not self._order:
                raise RuntimeError("finish() before any backward pass")
            self.all_reduce_mean()
            self._build_buckets()
            return
        if self._launched != len(self._buckets):
            raise RuntimeError(f"{len(self._buckets) - self._launched} gradient bucket(s) never completed: "
                               "the set of used parameters changed after the recording pass")
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        used = [i for bk in self._buckets for i in bk["params"]]
        torch._foreach_copy_([self.params[i].grad for i in used], [self.views[i] for i in used])
        self._reset()

    def remove_hooks(self):
        from . import ops

        if self._hooks:
            ops.forbid_side_stream(False)
        for h in self._hooks:
            h.remove()
        self._hooks = []


class LayerGradAllReduce:
    """Zero-copy, per-layer gradient averaging for an MGAT that runs through the layer executor
    (isubgvqa/executor.py) — the drop-in for DistributedDataParallel(find_unused_parameters=True), main.py:85-94.

    The executor writes every parameter gradient of the model into ONE flat fp32 buffer laid out in backward order
    (layer L-1 first) and hands autograd views of it, so `.grad` already points INTO the bucket: there is no pack
    and no unpack (the flat reducer copies 42 MB in and 42 MB out around its collective).  After each layer's
    backward the executor calls back with that layer's slice; its all-reduce is issued on a communication stream
    right away and runs while the earlier layers are still being differentiated — only layer 0's slice is exposed.
    `finish()` (after backward()) joins the communication stream.  Parameters outside the executor's set (the
    never-used gate_nn / gate_top / node_logits tensors, or anything else hanging off the module) go through a
    small flat GradAllReduce with DDP's unused-parameter semantics.

    The bucket is persistent (re-used every step) and is only handed out while every covered `.grad` is None —
    with gradients being accumulated across backward passes autograd would add a view of the bucket to itself, so
    the executor then falls back to a fresh buffer and the slices are reduced from there.
    Works under CUDA-graph capture (the communication stream forks from / joins the capturing stream)."""

    def __init__(self, model, process_group=None, overlap=True):
        from . import ops
        from .isubgvqa import executor

        ops.set_side_stream_with_dist(True)
        self.model = model
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.overlap = bool(overlap)
        self.covered = executor.flat_params(model)
        covered = {id(p) for p in self.covered}
        rest = [p for p in model.parameters() if p.requires_grad and id(p) not in covered]
        self.rest = GradAllReduce(model, process_group, params=rest) if rest else None
        dev = self.covered[0].device
        self.stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._avg = dist.is_initialized() and dist.get_backend(process_group) == "nccl"
        self._flat = None
        self._slices = []      # (buffer, lo, hi, is-the-bucket) handed over during the current backward
        self._issued = 0
        self._fallback = None
        self.nbytes = 4 * sum(p.numel() for p in self.covered)
        model._isg_grad_bucket = self._bucket
        model._isg_after_layer_backward = self._after_layer

    def detach(self):
        for name in ("_isg_grad_bucket", "_isg_after_layer_backward"):
            if self.model.__dict__.get(name) is not None:
                del self.model.__dict__[name]

    # ---- executor callbacks
    def _bucket(self, numel, device):
        if any(p.grad is not None for p in self.covered):
            return None  # accumulation step: autograd adds into the existing .grad, which may alias the bucket
        if self._flat is None or self._flat.numel() != numel or self._flat.device != device:
            self._flat = torch.empty(numel, dtype=torch.float32, device=device)
        return self._flat

    def _reduce(self, buf):
        if self._avg:
            dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            buf.mul_(1.0 / self.world)

    def _after_layer(self, layer, gflat, lo, hi):
        mine = self._flat is not None and gflat.data_ptr() == self._flat.data_ptr()
        self._slices.append((gflat, lo, hi, mine))
        if self.world == 1 or not self.overlap or not mine:
            return
        chunk = gflat[lo:hi]
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                self._reduce(chunk)
        else:
            self._reduce(chunk)
        self._issued += 1

    def finish(self):
        """Call after backward(): every gradient is averaged when the current stream continues."""
        slices, self._slices = self._slices, []
        issued, self._issued = self._issued, 0
        if not slices:
            raise RuntimeError("finish() without a backward pass through the layer executor (is the executor "
                               "enabled and the configuration supported?)")
        if self.world > 1:
            if not all(s[3] for s in slices):
                # accumulation step: .grad = previous (already averaged, identical on every rank) + this rank's new
                # values; averaging that sum across ranks is previous + mean(new) — the flat reducer does it
                if self._fallback is None:
                    self._fallback = GradAllReduce(self.model, self.group, params=self.covered)
                self._fallback.all_reduce_mean()
            elif issued:
                if self.stream is not None:
                    torch.cuda.current_stream().wait_stream(self.stream)
            else:
                buf = slices[0][0]
                lo, hi = min(s[1] for s in slices), max(s[2] for s in slices)
                self._reduce(buf[lo:hi])  # one collective over the whole (contiguous) bucket, in place
        if self.rest is not None:
            self.rest.all_reduce_mean()

    all_reduce_mean = finish


def shard_graphs(num_graphs_total, rank, world):
    """Contiguous, balanced split of whole graphs across ranks (DistributedSampler analogue,
    datasets/build.py:44-49, without shuffling): returns (first_graph, num_graphs) of this rank."""
    base, rem = divmod(num_graphs_total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def balanced_shards(sizes, world):
    """Size-aware sharding of whole graphs: every rank gets the same NUMBER of graphs and a near-equal total size.
    The reference's DistributedSampler (datasets/build.py:44-49) deals graphs out at random, so a step takes as long as
    the rank that happened to draw the biggest graphs (measured: +0.10 ms of a 4.8 ms step at 2 GPUs, more at 8 — more
    than the gradient all-reduce itself).  Graphs are sorted by size and dealt in snake order (0..W-1, W-1..0, ...);
    returns, per rank, the sorted list of graph ids.  len(sizes) must be a multiple of world."""
    sizes = [float(v) for v in sizes]
    n = len(sizes)
    if n % world:
        raise ValueError(f"{n} graphs do not split evenly over {world} ranks")
    order = sorted(range(n), key=lambda i: (-sizes[i], i))
    shards = [[] for _ in range(world)]
    for pos, g in enumerate(order):
        rnd, k = divmod(pos, world)
        shards[k if rnd % 2 == 0 else world - 1 - k].append(g)
    return [sorted(s) for s in shards]
