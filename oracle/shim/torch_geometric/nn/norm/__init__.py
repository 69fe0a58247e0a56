import torch
from .. import inits


class GraphNorm(torch.nn.Module):
    """PyG GraphNorm(eps=1e-5) (mgat.py:93-95,171):
    out = x - mean_g * mean_scale ; y = weight * out / sqrt(mean_g(out^2) + eps) + bias."""

    def __init__(self, in_channels, eps=1e-5):
        super().__init__()
        self.in_channels = in_channels
        self.eps = eps
        self.weight = torch.nn.Parameter(torch.empty(in_channels))
        self.bias = torch.nn.Parameter(torch.empty(in_channels))
        self.mean_scale = torch.nn.Parameter(torch.empty(in_channels))
        self.reset_parameters()

    def reset_parameters(self):
        inits.ones(self.weight)
        inits.zeros(self.bias)
        inits.ones(self.mean_scale)

    @staticmethod
    def _mean(x, batch, B):
        s = torch.zeros(B, x.size(1), dtype=x.dtype, device=x.device).index_add_(0, batch, x)
        c = torch.zeros(B, dtype=x.dtype, device=x.device).index_add_(
            0, batch, torch.ones(x.size(0), dtype=x.dtype, device=x.device)).clamp(min=1)
        return s / c.unsqueeze(-1)

    def forward(self, x, batch=None, batch_size=None):
        if batch is None:
            batch = x.new_zeros(x.size(0), dtype=torch.long)
            batch_size = 1
        if batch_size is None:
            batch_size = int(batch.max()) + 1
        mean = self._mean(x, batch, batch_size)
        out = x - mean.index_select(0, batch) * self.mean_scale
        var = self._mean(out.pow(2), batch, batch_size)
        std = (var + self.eps).sqrt().index_select(0, batch)
        return self.weight * out / std + self.bias
