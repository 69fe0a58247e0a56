import math
import torch


def uniform(size, value):
    if isinstance(value, torch.Tensor):
        bound = 1.0 / math.sqrt(size)
        value.data.uniform_(-bound, bound)


def glorot(value):
    if isinstance(value, torch.Tensor):
        stdv = math.sqrt(6.0 / (value.size(-2) + value.size(-1)))
        value.data.uniform_(-stdv, stdv)


def zeros(value):
    if isinstance(value, torch.Tensor):
        value.data.fill_(0.0)


def ones(value):
    if isinstance(value, torch.Tensor):
        value.data.fill_(1.0)


def reset(value):
    if hasattr(value, "reset_parameters"):
        value.reset_parameters()
    else:
        for child in value.children() if hasattr(value, "children") else []:
            reset(child)
