"""TEST INFRASTRUCTURE — pure-torch restatement of the torch_scatter==2.1.2 symbols the
reference hot path imports (requirements.txt:15).  The real wheel is not installable
offline; only the semantics used by ISubGVQA are restated:

  scatter            <- sampling/node_edge_masks.py:2,16 ; (GraphNorm / to_dense_batch via PyG)
  scatter_add        <- models/masking.py:5 ; utils/topk.py:1 ; models/att_pooling.py:3
  scatter_max        <- utils/topk.py:1
  scatter_softmax    <- utils/scatter_scaled_dot_product.py:1,7
  scatter_mean       <- models/scene_graph_encoder.py

Semantics (torch_scatter 2.1.2, `dim`-wise scatter with broadcasting of a 1-D index):
sum = zeros.scatter_add_; mean = sum / clamp(count, 1); max returns (values, argmax);
softmax = exp(src - max[index]) / sum(exp)[index]  (no epsilon, max not detached).
Never imported by the product package.
"""
import torch


def _broadcast(index, src, dim):
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(dim):
            index = index.unsqueeze(0)
    for _ in range(index.dim(), src.dim()):
        index = index.unsqueeze(-1)
    return index.expand(src.size())


def _out_size(src, index, dim, dim_size):
    size = list(src.size())
    if dim_size is not None:
        size[dim] = int(dim_size)
    elif index.numel() == 0:
        size[dim] = 0
    else:
        size[dim] = int(index.max()) + 1
    return size


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    index = _broadcast(index, src, dim)
    if out is None:
        out = torch.zeros(_out_size(src, index, dim, dim_size), dtype=src.dtype, device=src.device)
        return out.scatter_add_(dim, index, src)
    return out.scatter_add_(dim, index, src)


def scatter_add(src, index, dim=-1, out=None, dim_size=None):
    return scatter_sum(src, index, dim, out, dim_size)


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    out = scatter_sum(src, index, dim, out, dim_size)
    dim_size = out.size(dim)
    index_dim = dim
    if index_dim < 0:
        index_dim = index_dim + src.dim()
    if index.dim() <= index_dim:
        index_dim = index.dim() - 1
    ones = torch.ones(index.size(), dtype=src.dtype, device=src.device)
    count = scatter_sum(ones, index, index_dim, None, dim_size)
    count = count.clamp(min=1)
    count = _broadcast(count, out, dim)
    if out.is_floating_point():
        return out / count
    return torch.div(out, count, rounding_mode="floor")


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    index_b = _broadcast(index, src, dim)
    size = _out_size(src, index_b, dim, dim_size)
    if src.is_floating_point():
        fill = torch.finfo(src.dtype).min
    else:
        fill = torch.iinfo(src.dtype).min
    res = torch.full(size, fill, dtype=src.dtype, device=src.device)
    res = res.scatter_reduce(dim, index_b, src, reduce="amax", include_self=True)
    # argmax: first position attaining the max (torch_scatter returns an arg index; ties unspecified)
    hit = src == res.gather(dim, index_b)
    pos = torch.arange(src.size(dim), device=src.device)
    shape = [1] * src.dim()
    shape[dim] = -1
    pos = pos.view(shape).expand_as(src)
    big = src.size(dim)
    cand = torch.where(hit, pos, torch.full_like(pos, big))
    arg = torch.full(size, big, dtype=torch.long, device=src.device)
    arg = arg.scatter_reduce(dim, index_b, cand, reduce="amin", include_self=True)
    return res, arg


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    if reduce in ("sum", "add"):
        return scatter_sum(src, index, dim, out, dim_size)
    if reduce == "mean":
        return scatter_mean(src, index, dim, out, dim_size)
    if reduce == "max":
        return scatter_max(src, index, dim, out, dim_size)[0]
    raise ValueError(reduce)


def scatter_softmax(src, index, dim=-1, eps=1e-12, dim_size=None):
    if not torch.is_floating_point(src):
        raise ValueError("scatter_softmax needs floating point input")
    index = _broadcast(index, src, dim)
    max_value_per_index = scatter_max(src, index, dim=dim, dim_size=dim_size)[0]
    max_per_src_element = max_value_per_index.gather(dim, index)
    recentered = src - max_per_src_element
    recentered_exp = recentered.exp()
    sum_per_index = scatter_sum(recentered_exp, index, dim, dim_size=dim_size)
    normalizing = sum_per_index.gather(dim, index)
    return recentered_exp.div(normalizing)
