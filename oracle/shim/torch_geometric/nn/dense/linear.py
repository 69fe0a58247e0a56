import torch
import torch.nn.functional as F
from .. import inits


class Linear(torch.nn.Module):
    """PyG dense Linear: weight [out, in]; glorot weight init; bias U(+-1/sqrt(in))."""

    def __init__(self, in_channels, out_channels, bias=True, weight_initializer=None,
                 bias_initializer=None):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.weight_initializer = weight_initializer
        self.bias_initializer = bias_initializer
        self.weight = torch.nn.Parameter(torch.empty(out_channels, in_channels))
        if bias:
            self.bias = torch.nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        if self.weight_initializer == "glorot":
            inits.glorot(self.weight)
        else:
            torch.nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        if self.bias is not None:
            if self.bias_initializer == "zeros":
                inits.zeros(self.bias)
            else:
                inits.uniform(self.in_channels, self.bias)

    def forward(self, x):
        return F.linear(x, self.weight, self.bias)
