"""Batch construction with per-image cached CSR.

The reference caches each image's scene graph as a PyG `Data` in `GQADataset.sg_cache` (datasets/gqa.py:169-177,
built by datasets/scene_graph.py:199-389) and batches with `torch_geometric.data.Batch.from_data_list`
(datasets/gqa.py:260); the gather / scatter index structures are then rebuilt on the GPU inside every PyG layer.
Here the destination- and source-sorted CSR of an image is computed ONCE on the host, when the image first enters the
cache, and `collate_scene_graphs` only concatenates: nodes of a batch are numbered graph by graph and edges keep
their per-graph order, so the stable sort of the batch's edges by endpoint is the concatenation of the per-graph
sorted lists with node / edge offsets added — bit-identical to what csrc/csr.cu (`isg_csr_build`) produces on the
device, which `tests/test_collate.py` checks.  `GraphIndex.from_host` uploads the arrays (pinned, non-blocking) in
place of the device build; Nmax, the degree-ordered task schedule of the edge kernels and the closure flag are known
on the host too, so nothing is read back."""
import numpy as np
import torch


class GraphCsr:
    """Per-image CSR (int32 numpy): *_ptr [n+1], *_nbr [e] (the other endpoint), *_eid [e] (edge id within the image)."""

    __slots__ = ("n", "e", "dst_ptr", "dst_nbr", "dst_eid", "src_ptr", "src_nbr", "src_eid")

    def __init__(self, edge_index, num_nodes):
        ei = edge_index.cpu().numpy() if isinstance(edge_index, torch.Tensor) else np.asarray(edge_index)
        self.n, self.e = int(num_nodes), int(ei.shape[1])
        src, dst = ei[0].astype(np.int64), ei[1].astype(np.int64)
        if self.e and (src.min() < 0 or dst.min() < 0 or src.max() >= self.n or dst.max() >= self.n):
            raise IndexError(f"edge_index has endpoints outside [0, {self.n})")
        for name, key, other in (("dst", dst, src), ("src", src, dst)):
            order = np.argsort(key, kind="stable")  # == isg_csr_build: stable in the original edge id
            ptr = np.zeros(self.n + 1, dtype=np.int32)
            np.cumsum(np.bincount(key, minlength=self.n), out=ptr[1:])
            setattr(self, name + "_ptr", ptr)
            setattr(self, name + "_eid", order.astype(np.int32))
            setattr(self, name + "_nbr", other[order].astype(np.int32))


class SceneGraphCsrCache:
    """image_id -> GraphCsr, filled lazily (the companion of GQADataset.sg_cache, datasets/gqa.py:169-177)."""

    def __init__(self):
        self._cache = {}
        self.hits = self.misses = 0

    def get(self, image_id, edge_index, num_nodes):
        c = self._cache.get(image_id)
        if c is not None and c.n == int(num_nodes) and c.e == int(edge_index.shape[1]):
            self.hits += 1
            return c
        self.misses += 1
        c = self._cache[image_id] = GraphCsr(edge_index, num_nodes)
        return c

    def __len__(self):
        return len(self._cache)


def collate_scene_graphs(graphs, cache=None, pin=False):
    """Batch.from_data_list for scene graphs + the batch CSR from the per-image caches.

    graphs: sequence of dicts with `x` [n,D], `edge_index` [2,e] int64, `edge_attr` [e,De] and optionally `image_id`
    (cache key) or a ready `csr` (GraphCsr), `x_bbox` [n,4], `added_sym_edge` [m] (scene_graph_data.SceneGraphStore).
    Returns a dict: x, edge_index, edge_attr, batch (as PyG lays them out), x_bbox / added_sym_edge when given, and
    `host_index` = dict of int32 tensors (dst_ptr, dst_nbr, dst_eid, src_ptr, src_nbr, src_eid, graph_ptr, batch32,
    dst_order, src_order) + nmax, for GraphIndex.from_host."""
    B = len(graphs)
    ns = np.array([int(g["x"].shape[0]) for g in graphs], dtype=np.int64)
    es = np.array([int(g["edge_index"].shape[1]) for g in graphs], dtype=np.int64)
    n_off = np.concatenate([[0], np.cumsum(ns)])
    e_off = np.concatenate([[0], np.cumsum(es)])
    N, E = int(n_off[-1]), int(e_off[-1])
    csrs = []
    for i, g in enumerate(graphs):
        c = g.get("csr")
        if c is None:
            if cache is not None and g.get("image_id") is not None:
                c = cache.get(g["image_id"], g["edge_index"], ns[i])
            else:
                c = GraphCsr(g["edge_index"], ns[i])
        csrs.append(c)
    out = {
        "x": torch.cat([g["x"] for g in graphs], dim=0),
        "edge_attr": torch.cat([g["edge_attr"] for g in graphs], dim=0),
        "edge_index": torch.cat([g["edge_index"] for g in graphs], dim=1) + torch.from_numpy(np.repeat(n_off[:-1], es)),
        "batch": torch.repeat_interleave(torch.arange(B, dtype=torch.int64), torch.from_numpy(ns)),
    }
    if B and all(g.get("x_bbox") is not None for g in graphs):
        out["x_bbox"] = torch.cat([g["x_bbox"] for g in graphs], dim=0)
    if B and all(g.get("added_sym_edge") is not None for g in graphs):
        # PyG concatenates this attribute WITHOUT adding edge offsets (Data.__inc__ only shifts keys containing "index"),
        # and models/scene_graph_encoder.py:80 applies it to the batch's edge rows as is; `added_sym_edge_global` holds
        # the positions in the batch's edge numbering for callers that want what the name says
        out["added_sym_edge"] = torch.cat([g["added_sym_edge"] for g in graphs], dim=0)
        out["added_sym_edge_global"] = torch.cat([g["added_sym_edge"] + int(e_off[i]) for i, g in enumerate(graphs)], dim=0)
    idx = {}
    node_rep = np.repeat(n_off[:-1], es).astype(np.int32)  # node offset of the graph each edge belongs to
    edge_rep_e = np.repeat(e_off[:-1], es).astype(np.int32)  # edge offset, per edge
    edge_rep_n = np.repeat(e_off[:-1], ns).astype(np.int32)  # edge offset, per node
    for side in ("dst", "src"):
        ptr = np.zeros(N + 1, dtype=np.int32)
        if N:
            ptr[1:] = np.concatenate([getattr(c, side + "_ptr")[1:] for c in csrs]) + edge_rep_n
        nbr = (np.concatenate([getattr(c, side + "_nbr") for c in csrs]) + node_rep) if E else np.zeros(0, dtype=np.int32)
        eid = (np.concatenate([getattr(c, side + "_eid") for c in csrs]) + edge_rep_e) if E else np.zeros(0, dtype=np.int32)
        idx[side + "_ptr"], idx[side + "_nbr"], idx[side + "_eid"] = ptr, nbr.astype(np.int32), eid.astype(np.int32)
        heavy = np.diff(ptr) >= max(2, -(-2 * E // max(N, 1)))  # isg_degree_order's rule: >= twice the mean degree
        light = np.flatnonzero(~heavy)
        if side == "src":  # the src pass walks the light nodes backwards (csrc/csr.cu order_place_kernel)
            light = light[::-1]
        idx[side + "_order"] = np.concatenate([np.flatnonzero(heavy), light]).astype(np.int32)
    idx["graph_ptr"] = n_off.astype(np.int32)
    idx["batch32"] = np.repeat(np.arange(B, dtype=np.int32), ns) if N else np.zeros(1, dtype=np.int32)
    host = {k: torch.from_numpy(v) for k, v in idx.items()}
    if pin and torch.cuda.is_available():
        host = {k: v.pin_memory() for k, v in host.items()}
        out = {k: v.pin_memory() for k, v in out.items()}
    host["nmax"] = int(ns.max()) if B else 0
    host["num_graphs"] = B
    out["host_index"] = host
    return out
