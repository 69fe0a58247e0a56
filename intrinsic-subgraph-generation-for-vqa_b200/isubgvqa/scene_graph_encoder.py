"""Mirror of the graph part of ISubGVQA/models/scene_graph_encoder.py: the
MetaLayer(EdgeModel, NodeModel) built by get_gt_scene_graph_encoding_layer (:107-146) and the GraphNorm that
SceneGraphEncoder.forward evaluates in float64 through a CPU round trip (:99-102).  Same module tree and
state_dict keys as the reference (`edge_model.edge_mlp.{0,2}`, `node_model.node_mlp_{1,2}.{0,2}`; GraphNorm
`weight / bias / mean_scale`), so a reference checkpoint's `scene_graph_encoder.scene_graph_encoding_layer.*` and
`scene_graph_encoder.graph_layer_norm.*` entries load unchanged.  The vocabulary embedding, bbox MLP and
feat_reduc in front of it need GloVe / the GQA vocabulary and stay with the reference (out of scope)."""
import torch

from .. import lib as L
from .. import ops
from ..graph import get_graph_index
from .mgat import GraphNormParams


def _mlp(in_dim, hidden_dim):
    return torch.nn.Sequential(torch.nn.Linear(in_dim, hidden_dim), torch.nn.GELU(),
                               torch.nn.Linear(hidden_dim, hidden_dim))


class EdgeModel(torch.nn.Module):
    """scene_graph_encoder.py:108-120."""

    def __init__(self, num_node_features, num_edge_features, hidden_dim=300):
        super().__init__()
        self.hidden_dim = hidden_dim
        self.edge_mlp = _mlp(2 * num_node_features + num_edge_features, hidden_dim)


class NodeModel(torch.nn.Module):
    """scene_graph_encoder.py:122-143."""

    def __init__(self, num_node_features, hidden_dim=300):
        super().__init__()
        self.hidden_dim = hidden_dim
        self.node_mlp_1 = _mlp(num_node_features + hidden_dim, hidden_dim)
        self.node_mlp_2 = _mlp(num_node_features + hidden_dim, hidden_dim)


class SceneGraphEncodingLayer(torch.nn.Module):
    """torch_geometric.nn.MetaLayer(EdgeModel, NodeModel).forward(x, edge_index, edge_attr, u, batch) ->
    (x', edge_attr', u).  The first Linear of each MLP acts on a concatenation; its weight is split by column
    block so the node blocks run once per node (ops.linear on N rows) and are gathered per edge by one fused
    kernel, instead of materialising [E, 900] / [E, 600] concatenations and multiplying E rows."""

    def __init__(self, num_node_features=300, num_edge_features=300, hidden_dim=300):
        super().__init__()
        self.nf, self.ef, self.hd = num_node_features, num_edge_features, hidden_dim
        self.edge_model = EdgeModel(num_node_features, num_edge_features, hidden_dim)
        self.node_model = NodeModel(num_node_features, hidden_dim)
        self.global_model = None

    def forward(self, x, edge_index, edge_attr=None, u=None, batch=None, gi=None):
        L.require_cuda(x, edge_index, edge_attr)
        if gi is None:
            if batch is None:
                batch = torch.zeros(x.shape[0], dtype=torch.int64, device=x.device)
            gi = get_graph_index(edge_index, batch, int(batch[-1].item()) + 1 if batch.numel() else 0)
        nf = self.nf
        em, n1, n2 = self.edge_model.edge_mlp, self.node_model.node_mlp_1, self.node_model.node_mlp_2
        # EdgeModel: edge_mlp(cat[x[src], x[dst], e])
        w = em[0].weight
        p_src = ops.linear(x, w[:, :nf])
        p_dst = ops.linear(x, w[:, nf:2 * nf])
        q = ops.linear(edge_attr, w[:, 2 * nf:], em[0].bias)
        e2 = ops.linear(ops.GatherAddAct.apply(p_src, p_dst, q, gi, L.ACT_GELU), em[2].weight, em[2].bias)
        # NodeModel: node_mlp_1(cat[x[src], e']) -> scatter_mean by dst -> node_mlp_2(cat[x, agg])
        w = n1[0].weight
        r = ops.linear(e2, w[:, nf:], n1[0].bias)
        m = ops.linear(ops.GatherAddAct.apply(ops.linear(x, w[:, :nf]), None, r, gi, L.ACT_GELU), n1[2].weight,
                       n1[2].bias)
        agg = ops.SegmentMeanByDst.apply(m, gi)
        hcat = torch.cat([x, agg], dim=1)
        x2 = ops.linear(ops.linear(hcat, n2[0].weight, n2[0].bias, L.ACT_GELU), n2[2].weight, n2[2].bias)
        return x2, e2, u


def get_gt_scene_graph_encoding_layer(num_node_features, num_edge_features, hidden_dim):
    """Same factory name and arguments as scene_graph_encoder.py:107."""
    return SceneGraphEncodingLayer(num_node_features, num_edge_features, hidden_dim)


class GraphNorm64(GraphNormParams):
    """`self.graph_layer_norm` of SceneGraphEncoder applied the way its forward does (:99-102): in float64, result
    cast back — on the device, without the PCIe round trip."""

    def forward(self, x, batch=None, batch_size=None, gi=None):
        if gi is None:
            B = int(batch_size) if batch_size is not None else int(batch[-1].item()) + 1
            gi = get_graph_index(torch.zeros(2, 0, dtype=torch.int64, device=x.device), batch, B)
        return ops.GraphNorm64.apply(x.float(), self.weight, self.bias, self.mean_scale, gi, float(self.eps))


def encode_scene_graph(layer, graph_layer_norm, x, edge_index, edge_attr, batch, num_graphs=None):
    """The tail of SceneGraphEncoder.forward (:91-104): (x_encoded, edge_attr_encoded)."""
    if num_graphs is None:
        num_graphs = int(batch[-1].item()) + 1
    gi = get_graph_index(edge_index, batch, num_graphs)
    x2, e2, _ = layer(x, edge_index, edge_attr, None, batch, gi=gi)
    return graph_layer_norm(x2, batch, gi=gi), e2
