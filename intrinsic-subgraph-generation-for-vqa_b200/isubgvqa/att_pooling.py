"""Mirror of ISubGVQA/models/att_pooling.py (GlobalAttention) — SURVEY.md section 8 row f1, the immediate
consumer of MGAT's (h, mask) (models/isubgvqa.py:280-287)."""
import torch

from .. import lib as L
from .. import ops
from ..graph import get_graph_index


class GlobalAttention(torch.nn.Module):
    """att_pooling.py:6-82.  Same constructor, parameters (gate_nn is built and never used by the reference's
    forward either) and forward(x, u, batch, size=None, return_mask=False, node_mask=None) -> out [B,D]
    (and the per-graph softmax gate [N,1] with return_mask=True).  The reference's hard-coded `batch.cuda()`
    calls and the `batch[-1].item()` sync are gone: the graph offsets come from the cached GraphIndex."""

    def __init__(self, num_node_features, num_out_features):
        super().__init__()
        channels = num_out_features
        self.gate_nn = torch.nn.Sequential(torch.nn.Linear(channels, channels), torch.nn.GELU(),
                                           torch.nn.Linear(channels, 1))
        self.node_nn = torch.nn.Sequential(torch.nn.Linear(num_node_features, channels), torch.nn.GELU(),
                                           torch.nn.Linear(channels, channels))
        self.ques_nn = torch.nn.Sequential(torch.nn.Linear(channels, channels), torch.nn.GELU(),
                                           torch.nn.Linear(channels, channels))

    def reset_parameters(self):
        for seq in (self.gate_nn, self.node_nn, self.ques_nn):
            for m in seq:
                if hasattr(m, "reset_parameters"):
                    m.reset_parameters()

    def forward(self, x, u, batch, size=None, return_mask=False, node_mask=None, edge_index=None):
        L.require_cuda(x, u, batch)
        x = x.unsqueeze(-1) if x.dim() == 1 else x
        num_graphs = int(u.shape[0]) if size is None else int(size)
        if edge_index is None:  # only graph_ptr / nmax are needed here; reuse the MGAT index when it is cached
            edge_index = torch.zeros(2, 0, dtype=torch.int64, device=x.device)
        gi = get_graph_index(edge_index, batch, num_graphs)
        x = ops.linear(x, self.node_nn[0].weight, self.node_nn[0].bias, L.ACT_GELU)
        x = ops.linear(x, self.node_nn[2].weight, self.node_nn[2].bias)
        q = ops.linear(u, self.ques_nn[0].weight, self.ques_nn[0].bias, L.ACT_GELU)
        q = ops.linear(q, self.ques_nn[2].weight, self.ques_nn[2].bias)
        out, gate = ops.AttnPool.apply(x, node_mask, q, gi)
        if return_mask:
            return out, gate
        return out

    def __repr__(self):
        return "{}(gate_nn={}, node_nn={}, ques_nn={})".format(self.__class__.__name__, self.gate_nn, self.node_nn,
                                                               self.ques_nn)
