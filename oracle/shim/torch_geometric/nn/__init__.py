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
