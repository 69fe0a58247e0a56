"""Host -> device staging of scene-graph batches, one batch ahead (the loader side of
the path).  The reference moves each batch with `.to("cuda", non_blocking=True)` on the compute stream right
before the forward pass (training/train_epoch.py:66-73) and lets PyG rebuild its gather indices inside every
layer; here the pinned-host -> HBM copies and the CSR / graph_ptr build (csrc/csr.cu) of batch i+1 run on a copy
stream while batch i is being computed, and the number of nodes of the largest graph — known on the host at collate
time — is handed to the GraphIndex so no device -> host read is needed."""
import torch

from .graph import get_graph_index

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
            if build_index:
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
            tensors += [getattr(gi, n) for n in ("dst_ptr", "dst_nbr", "dst_eid", "src_ptr", "src_nbr", "src_eid",
                                                 "status", "graph_ptr", "batch32", "_nmax_dev")]
        for t in tensors:
            t.record_stream(main)  # allocated on the copy stream, consumed on the compute stream
        return dev, ext
