"""Mirror of ISubGVQA/sampling/node_edge_masks.py:5-19."""
import torch

from .. import ops
from ..graph import get_graph_index


class NodeMaskToEdgeMask:
    """`NodeMaskToEdgeMask.apply(mask, edge_index, n_nodes)` — same call contract as the reference's
    autograd.Function (3rd argument is a 0-dim tensor holding N, mgat_v2_conv.py:169-171)."""

    @staticmethod
    def apply(mask, edge_index, n_nodes, gi=None):
        if gi is None:
            n = int(n_nodes) if not torch.is_tensor(n_nodes) else int(n_nodes.item())
            batch = torch.zeros(n, dtype=torch.int64, device=mask.device)
            gi = get_graph_index(edge_index, batch, 1)
        return ops.NodeMaskToEdgeMaskFn.apply(mask, gi)
