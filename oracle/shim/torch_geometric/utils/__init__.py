"""torch_geometric.utils symbols used on the hot path (PyG 2.6.1 semantics)."""
import torch


def _scatter(src, index, dim_size, reduce):
    shape = [dim_size] + list(src.shape[1:])
    idx = index.view([-1] + [1] * (src.dim() - 1)).expand_as(src)
    if reduce == "sum":
        return torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add_(0, idx, src)
    if reduce == "max":
        out = torch.full(shape, torch.finfo(src.dtype).min, dtype=src.dtype, device=src.device)
        return out.scatter_reduce(0, idx, src, reduce="amax", include_self=True)
    raise ValueError(reduce)


def softmax(src, index=None, ptr=None, num_nodes=None, dim=0):
    """PyG utils.softmax for an `index` vector (mgat_v2_conv.py:272, att_pooling.py:71):
    max is taken on src.detach(); out = exp(src-max) / (segment_sum + 1e-16)."""
    assert index is not None and dim == 0
    N = int(index.max()) + 1 if num_nodes is None else int(num_nodes)
    src_max = _scatter(src.detach(), index, N, "max")
    out = src - src_max.index_select(0, index)
    out = out.exp()
    out_sum = _scatter(out, index, N, "sum") + 1e-16
    return out / out_sum.index_select(0, index)


def to_dense_batch(x, batch=None, fill_value=0.0, max_num_nodes=None, batch_size=None):
    """masking.py:145,162.  Returns ([B, Nmax, *], bool [B, Nmax])."""
    if batch is None:
        mask = torch.ones(1, x.size(0), dtype=torch.bool, device=x.device)
        return x.unsqueeze(0), mask
    if batch_size is None:
        batch_size = int(batch.max()) + 1
    num_nodes = torch.zeros(batch_size, dtype=batch.dtype, device=x.device).scatter_add_(
        0, batch, batch.new_ones(x.size(0)))
    cum_nodes = torch.cat([batch.new_zeros(1), num_nodes.cumsum(dim=0)])
    if max_num_nodes is None:
        max_num_nodes = int(num_nodes.max())
    tmp = torch.arange(batch.size(0), device=x.device) - cum_nodes[batch]
    idx = tmp + (batch * max_num_nodes)
    size = [batch_size * max_num_nodes] + list(x.size())[1:]
    out = torch.as_tensor(fill_value, device=x.device).to(x.dtype).repeat(size)
    out[idx] = x
    out = out.view([batch_size, max_num_nodes] + list(x.size())[1:])
    mask = torch.zeros(batch_size * max_num_nodes, dtype=torch.bool, device=x.device)
    mask[idx] = 1
    mask = mask.view(batch_size, max_num_nodes)
    return out, mask


def remove_self_loops(edge_index, edge_attr=None):
    mask = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, mask]
    return edge_index, (None if edge_attr is None else edge_attr[mask])


def add_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
    raise NotImplementedError("dead on the hot path: mgat.py:63 passes add_self_loops=False")


def index_sort(inputs, max_value=None, stable=False):
    return inputs.sort(stable=stable)
