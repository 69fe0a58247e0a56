"""Host -> device staging of scene-graph batches, one batch ahead (the loader side of
the path).  The reference moves each batch with `.to("cuda", non_blocking=True)` on the compute stream right
before the forward pass (training/train_epoch.py:66-73) and lets PyG rebuild its gather indices inside every
layer; here the pinned-host -> HBM copies and the CSR / graph_ptr build (csrc/csr.cu) of batch i+1 run on a copy
stream while batch i is being computed, and the number of nodes of the largest graph — known on the host at collate
time — is handed to the GraphIndex so no device -> host read is needed."""
import torch

from .graph import GraphIndex, get_graph_index, register_graph_index

_KEYS = ("x", "edge_index", "instr_vectors", "global_language_feats", "edge_attr", "batch")


class DevicePrefetcher:
    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)

    def stage(self, host_batch, extra=None, nmax=None, build_index=True):
        """Issues the copies (and the CSR build) of one pinned host batch on the copy stream.
        host_batch: dict with the MGAT inputs (_KEYS); extra: dict of further pinned tensors (e.g. sampler noise).
        Returns a handle for `get`."""
        main = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(main)  # buffers freed by the main stream may be reused by these allocations
        with torch.cuda.stream(self.stream):
            dev = {k: host_batch[k].to(self.device, non_blocking=True) for k in _KEYS if k in host_batch}
            ext = {k: v.to(self.device, non_blocking=True) for k, v in (extra or {}).items()}
            gi = None
            if build_index and host_batch.get("host_index") is not None:
                # CSR built at collate time from the per-image cache (isg_b200.collate): upload, no device build
                gi = register_graph_index(GraphIndex.from_host(dev["edge_index"], dev["batch"], host_batch["host_index"]))
            elif build_index:
                gi = get_graph_index(dev["edge_index"], dev["batch"], int(dev["instr_vectors"].shape[1]))
                if nmax is not None:
                    gi.set_nmax(nmax)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return dev, ext, gi, ev

    def get(self, handle):
        """Makes the compute stream wait for a staged batch and returns (device batch, extra tensors)."""
        dev, ext, gi, ev = handle
        main = torch.cuda.current_stream(self.device)
        main.wait_event(ev)
        tensors = list(dev.values()) + list(ext.values())
        if gi is not None:
            tensors += [t for t in (getattr(gi, n) for n in ("dst_ptr", "dst_nbr", "dst_eid", "src_ptr", "src_nbr", "src_eid",
                                                 "dst_order", "src_order", "status", "graph_ptr", "batch32", "_nmax_dev"))
                        if t is not None]
        for t in tensors:
            t.record_stream(main)  # allocated on the copy stream, consumed on the compute stream
        return dev, ext


class HostResults:
    """Device -> host read of each step's results without stalling the step behind it.

    The reference's loop reads the loss with `.item()` right after `backward()` (training/train_epoch.py:120-133),
    which idles the GPU while the host enqueues the next step.  Here every step's results are copied into pinned
    host buffers on the compute stream (`push`), and the host picks them up `lag` steps later (`pop` waits on that
    copy's event only): with lag 1 the host is always one step ahead of the GPU.  `drain()` returns whatever is
    still in flight (call it after the last step)."""

    def __init__(self, lag=1):
        self.lag = int(lag)
        self._inflight = []
        self._free = []

    def push(self, **tensors):
        bufs = self._free.pop() if self._free else {}
        out = {}
        for k, t in tensors.items():
            t = t.detach()
            b = bufs.get(k)
            if b is None or b.shape != t.shape or b.dtype != t.dtype:
                b = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            b.copy_(t, non_blocking=True)
            out[k] = b
        ev = torch.cuda.Event()
        ev.record()
        self._inflight.append((out, ev))

    def pop(self):
        """Results of the step pushed `lag` steps ago, or None while fewer than lag + 1 steps are in flight."""
        if len(self._inflight) <= self.lag:
            return None
        return self._take()

    def drain(self):
        res = []
        while self._inflight:
            res.append(self._take())
        return res

    def _take(self):
        out, ev = self._inflight.pop(0)
        ev.synchronize()
        res = {k: (v.item() if v.dim() == 0 else v.clone()) for k, v in out.items()}
        self._free.append(out)
        return res
