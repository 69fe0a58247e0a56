import inspect
import torch


class MessagePassing(torch.nn.Module):
    """PyG MessagePassing restricted to what MaskingGATv2Conv uses (mgat_v2_conv.py:47,215):
    Tensor COO edge_index, flow source_to_target, aggr='add', node_dim=0.
      x_j = x[0].index_select(0, edge_index[0]); x_i = x[1].index_select(0, edge_index[1])
      index = edge_index[1]; ptr = None; size_i = x[1].size(0)
      out = zeros(size_i, ...).index_add_(0, index, message(...))
    Extra kwargs are forwarded to message() by name."""

    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2, **kwargs):
        super().__init__()
        assert aggr in ("add", "sum") and flow == "source_to_target"
        self.aggr = aggr
        self.flow = flow
        self.node_dim = node_dim
        self._msg_params = None

    def propagate(self, edge_index, size=None, **kwargs):
        assert isinstance(edge_index, torch.Tensor) and self.node_dim == 0
        if self._msg_params is None:
            self._msg_params = list(inspect.signature(self.message).parameters)
        x = kwargs.get("x")
        x_src, x_dst = (x if isinstance(x, (tuple, list)) else (x, x))
        src, dst = edge_index[0], edge_index[1]
        avail = dict(kwargs)
        avail.update(
            x_j=x_src.index_select(0, src),
            x_i=x_dst.index_select(0, dst),
            index=dst, ptr=None, size_i=x_dst.size(0), size_j=x_src.size(0),
            edge_index=edge_index,
        )
        msg = self.message(**{k: avail[k] for k in self._msg_params})
        out = torch.zeros((x_dst.size(0),) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device)
        return out.index_add_(0, dst, msg)

    def message(self, x_j):
        return x_j
