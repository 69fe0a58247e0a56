"""GraphIndex — device-side index structures shared by every layer of one forward/backward:
dst- and src-sorted CSR of the COO `edge_index` (kernel (a), csrc/csr.cu), `graph_ptr` from the
sorted `batch` vector, and Nmax.  Built once per batch (the reference rebuilds gathers / scatter
indices inside PyG on every one of the 4 layers: models/mgat_v2_conv.py:215)."""
import torch

import os

from . import lib as L

_EDGE_ORDER = os.environ.get("ISG_EDGE_ORDER", "1") != "0"  # A/B switch: "0" keeps the natural node order
# The heavy-first schedule pays where a launch is a handful of CTA waves (c3: 4.7 waves, backward 0.373 -> 0.408 and
# forward 0.518 -> 0.563 of the HBM peak in-step); at batch 4096 (92 waves) the tail is already negligible and the
# moved nodes only cost locality (forward 0.779 -> 0.751), so larger batches keep the natural order.
EDGE_ORDER_MAX_NODES = int(os.environ.get("ISG_EDGE_ORDER_MAX_NODES", "32768"))


class GraphIndex:
    __slots__ = ("edge_index", "batch", "N", "E", "B", "dst_ptr", "dst_nbr", "dst_eid", "src_ptr", "src_nbr",
                 "src_eid", "dst_order", "src_order", "status", "graph_ptr", "batch32", "_nmax_dev", "_nmax", "_closed", "_key",
                 "_keepalive")

    def __init__(self, edge_index, batch, num_graphs, num_nodes=None):
        L.require_cuda(edge_index, batch)
        if edge_index.dtype != torch.int64 or batch.dtype != torch.int64:
            raise TypeError("edge_index and batch must be int64 (PyG layout)")
        lib = L.load()
        dev = edge_index.device
        self.edge_index = edge_index.contiguous()
        self.batch = batch.contiguous()
        self.N = int(batch.numel()) if num_nodes is None else int(num_nodes)
        self.E = int(edge_index.size(1))
        self.B = int(num_graphs)
        N, E, B = self.N, self.E, self.B
        i32 = dict(dtype=torch.int32, device=dev)
        self.dst_ptr = torch.empty(N + 1, **i32)
        self.src_ptr = torch.empty(N + 1, **i32)
        self.dst_nbr = torch.empty(E, **i32)
        self.dst_eid = torch.empty(E, **i32)
        self.src_nbr = torch.empty(E, **i32)
        self.src_eid = torch.empty(E, **i32)
        self.status = torch.empty(1, **i32)
        ws_bytes = lib.isg_csr_workspace_bytes(N, E)
        ws = L.workspace(ws_bytes, dev)
        st = L.stream()
        L.call("isg_csr_build", L.ptr(self.edge_index), E, N, L.ptr(self.dst_ptr), L.ptr(self.dst_nbr),
                                  L.ptr(self.dst_eid), L.ptr(self.src_ptr), L.ptr(self.src_nbr),
                                  L.ptr(self.src_eid), L.ptr(self.status), L.ptr(ws), ws_bytes, st)
        # task order of the edge kernels: longest segments first (csrc/csr.cu, isg_degree_order)
        self.dst_order = self.src_order = None
        if _EDGE_ORDER and N <= EDGE_ORDER_MAX_NODES:
            self.dst_order = torch.empty(max(N, 1), **i32)
            self.src_order = torch.empty(max(N, 1), **i32)
            ows = L.workspace(lib.isg_degree_order_workspace_bytes(N), dev)
            L.call("isg_degree_order", L.ptr(self.dst_ptr), L.ptr(self.src_ptr), N, E, L.ptr(self.dst_order),
                   L.ptr(self.src_order), L.ptr(ows), ows.numel(), st)
        self.graph_ptr = torch.empty(B + 1, **i32)
        self.batch32 = torch.empty(max(N, 1), **i32)
        self._nmax_dev = torch.empty(2, **i32)  # [nmax, number of edges that leave their graph]
        L.call("isg_graph_ptr", L.ptr(self.batch), N, B, L.ptr(self.graph_ptr), L.ptr(self.batch32),
                                  L.ptr(self._nmax_dev), st)
        L.call("isg_graph_closure", L.ptr(self.edge_index), E, L.ptr(self.batch), N,
               self._nmax_dev.data_ptr() + 4, st)
        self._nmax = None
        self._closed = None
        self._key = None
        self._keepalive = None

    @classmethod
    def from_host(cls, edge_index, batch, host_index):
        """GraphIndex from the CSR a data loader built on the host (isg_b200.collate.collate_scene_graphs: per-image
        cached CSR, concatenated at collate time) — uploads the int32 arrays instead of running the device build.
        edge_index / batch are the batch's DEVICE tensors; a per-image cache guarantees that no edge leaves its graph."""
        L.require_cuda(edge_index, batch)
        self = cls.__new__(cls)
        dev = edge_index.device
        self.edge_index, self.batch = edge_index.contiguous(), batch.contiguous()
        self.N, self.E, self.B = int(batch.numel()), int(edge_index.size(1)), int(host_index["num_graphs"])
        for name in ("dst_ptr", "dst_nbr", "dst_eid", "src_ptr", "src_nbr", "src_eid", "graph_ptr", "batch32",
                     "dst_order", "src_order"):
            t = host_index[name]
            setattr(self, name, t.to(dev, non_blocking=True) if not t.is_cuda else t)
        if not _EDGE_ORDER or self.N > EDGE_ORDER_MAX_NODES:
            self.dst_order = self.src_order = None
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self._nmax_dev = None
        self._nmax, self._closed = int(host_index["nmax"]), True
        self._key = self._keepalive = None
        return self

    @property
    def nmax(self):
        """Max nodes per graph as a Python int.  One device->host read per batch (the reference
        syncs here too: to_dense_batch's int(num_nodes.max()), models/masking.py:162)."""
        if self._nmax is None:
            self._read_host()
        return self._nmax

    @property
    def closed(self):
        """True when no edge leaves its graph (always so for PyG batches).  Read together with nmax."""
        if self._closed is None:
            self._read_host()
        return self._closed

    def _read_host(self):
        nmax, crossing = self._nmax_dev.tolist()
        if self._nmax is None:
            self._nmax = int(nmax)
        if self._closed is None:
            self._closed = crossing == 0

    def set_nmax(self, nmax, closed=True):
        """Lets a data loader that already knows the graph sizes (and that its edges stay inside their graphs)
        skip the device->host read."""
        self._nmax = int(nmax)
        self._closed = bool(closed)

    def check_indices(self):
        bad = int(self.status.item())
        if bad:
            raise IndexError(f"edge_index has {bad} endpoints outside [0, {self.N})")


_cache = []  # tiny MRU cache so the 4 conv layers of one forward share one index
_CACHE_SIZE = 4


def get_graph_index(edge_index, batch, num_graphs):
    """The key is the identity of the caller's storage (address, offset, strides, version counter).  The cache
    entry keeps the caller's tensors alive, so the allocator cannot hand the same address to a different batch
    while the entry exists; inputs that had to be copied to become contiguous are indexed but never cached."""
    if not (edge_index.is_contiguous() and batch.is_contiguous()):
        return GraphIndex(edge_index, batch, num_graphs)
    key = (edge_index.data_ptr(), edge_index.storage_offset(), edge_index._version, tuple(edge_index.shape),
           batch.data_ptr(), batch.storage_offset(), batch._version, int(batch.numel()), int(num_graphs),
           edge_index.device)
    for gi in _cache:
        if gi._key == key:
            return gi
    gi = GraphIndex(edge_index, batch, num_graphs)
    gi._key = key
    gi._keepalive = (edge_index, batch)
    _cache.insert(0, gi)
    del _cache[_CACHE_SIZE:]
    return gi


def register_graph_index(gi):
    """Puts a GraphIndex built elsewhere (GraphIndex.from_host) into the cache, so that MGAT.forward — which looks
    the index up by the identity of the edge_index / batch tensors it is given — finds it."""
    edge_index, batch = gi.edge_index, gi.batch
    gi._key = (edge_index.data_ptr(), edge_index.storage_offset(), edge_index._version, tuple(edge_index.shape),
               batch.data_ptr(), batch.storage_offset(), batch._version, int(batch.numel()), int(gi.B), edge_index.device)
    gi._keepalive = (edge_index, batch)
    _cache.insert(0, gi)
    del _cache[_CACHE_SIZE:]
    return gi


def clear_cache():
    del _cache[:]
