import torch
from . import inits, conv, dense, norm  # noqa: F401
from .conv import MessagePassing  # noqa: F401
from .norm import GraphNorm  # noqa: F401


class _SelectTopK(torch.nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.weight = torch.nn.Parameter(torch.empty(1, in_channels))
        self.reset_parameters()

    def reset_parameters(self):
        inits.uniform(self.in_channels, self.weight)


class TopKPooling(torch.nn.Module):
    """Parameters only (masking.py:89-90 builds it, nothing calls it): state_dict key
    `select.weight` [1, in_channels] as in PyG 2.6.1."""

    def __init__(self, in_channels, ratio=0.5, min_score=None, multiplier=1.0, nonlinearity="tanh"):
        super().__init__()
        self.in_channels = in_channels
        self.ratio = ratio
        self.select = _SelectTopK(in_channels)

    def reset_parameters(self):
        self.select.reset_parameters()


class MetaLayer(torch.nn.Module):
    """torch_geometric.nn.MetaLayer (PyG 2.6.1) as used by models/scene_graph_encoder.py:145:
        row, col = edge_index
        edge_attr = edge_model(x[row], x[col], edge_attr, u, batch[row])      (if edge_model)
        x = node_model(x, edge_index, edge_attr, u, batch)                    (if node_model)
        u = global_model(x, edge_index, edge_attr, u, batch)                  (if global_model)
        return x, edge_attr, u"""

    def __init__(self, edge_model=None, node_model=None, global_model=None):
        super().__init__()
        self.edge_model = edge_model
        self.node_model = node_model
        self.global_model = global_model
        self.reset_parameters()

    def reset_parameters(self):
        for item in (self.node_model, self.edge_model, self.global_model):
            if hasattr(item, "reset_parameters"):
                item.reset_parameters()

    def forward(self, x, edge_index, edge_attr=None, u=None, batch=None):
        row, col = edge_index[0], edge_index[1]
        if self.edge_model is not None:
            edge_attr = self.edge_model(x[row], x[col], edge_attr, u, batch if batch is None else batch[row])
        if self.node_model is not None:
            x = self.node_model(x, edge_index, edge_attr, u, batch)
        if self.global_model is not None:
            u = self.global_model(x, edge_index, edge_attr, u, batch)
        return x, edge_attr, u
