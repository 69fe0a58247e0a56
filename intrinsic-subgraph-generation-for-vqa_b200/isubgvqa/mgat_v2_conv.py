"""Mirror of ISubGVQA/models/mgat_v2_conv.py (MaskingGATv2Conv)."""
import math

import torch

from .. import lib as L
from .. import ops
from ..graph import get_graph_index
from .masking import MaskingModel


class PygLinear(torch.nn.Module):
    """Parameter container with torch_geometric.nn.dense.linear.Linear's layout and init
    (weight [out,in] glorot, bias U(+-1/sqrt(in)))."""

    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = torch.nn.Parameter(torch.empty(out_channels, in_channels))
        if bias:
            self.bias = torch.nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        a = math.sqrt(6.0 / (self.in_channels + self.out_channels))
        torch.nn.init.uniform_(self.weight, -a, a)
        if self.bias is not None:
            b = 1.0 / math.sqrt(self.in_channels)
            torch.nn.init.uniform_(self.bias, -b, b)

    def forward(self, x):
        return ops.linear(x, self.weight, self.bias)


class MaskingGATv2Conv(torch.nn.Module):
    """mgat_v2_conv.py:18-285.  forward(x, edge_index, batch, edge_attr, instruction, imle_att,
    return_attention_weights, return_masks, all_instrs) -> (out [N,H*C], mask, (edge_index, alpha [E,H]))."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2, dropout=0.0,
                 add_self_loops=True, edge_dim=None, fill_value="mean", bias=True, share_weights=False,
                 masking_threshold=None, use_instr=False, use_topk=False, concat_instr=False,
                 use_all_instrs=False, sampler_type=None, sample_k=None, nb_samples=1, alpha=1.0, beta=10.0,
                 tau=1.0, **kwargs):
        super().__init__()
        if add_self_loops:
            raise NotImplementedError("MGAT builds the conv with add_self_loops=False (mgat.py:63)")
        if not concat or dropout != 0.0 or use_all_instrs or not isinstance(in_channels, int):
            raise NotImplementedError("only concat=True, dropout=0, use_all_instrs=False and an int in_channels are on "
                                      "the ISubGVQA path")
        if edge_dim is None:
            raise NotImplementedError("MGAT always passes edge_dim (mgat.py:61)")
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.negative_slope, self.dropout = concat, negative_slope, dropout
        self.add_self_loops, self.edge_dim, self.fill_value = add_self_loops, edge_dim, fill_value
        self.share_weights, self.use_instr = share_weights, use_instr
        self.concat_instr, self.use_all_instrs = concat_instr, use_all_instrs
        self.lin_l = PygLinear(in_channels, heads * out_channels, bias=bias)
        self.lin_r = self.lin_l if share_weights else PygLinear(in_channels, heads * out_channels, bias=bias)
        self.att = torch.nn.Parameter(torch.empty(1, heads, out_channels))
        self.lin_edge = PygLinear(edge_dim, heads * out_channels, bias=False)
        if bias:
            self.bias = torch.nn.Parameter(torch.empty(heads * out_channels))
        else:
            self.register_parameter("bias", None)
        self.mask = MaskingModel(in_channels, out_channels, masking_threshold, use_topk=use_topk,
                                 sampler_type=sampler_type, sample_k=sample_k, nb_samples=nb_samples, alpha=alpha,
                                 beta=beta, tau=tau)
        self.reset_parameters()

    def reset_parameters(self):
        self.lin_l.reset_parameters()
        self.lin_r.reset_parameters()
        self.lin_edge.reset_parameters()
        a = math.sqrt(6.0 / (self.att.size(-2) + self.att.size(-1)))
        torch.nn.init.uniform_(self.att, -a, a)
        if self.bias is not None:
            torch.nn.init.zeros_(self.bias)

    def forward(self, x, edge_index, batch, edge_attr=None, instruction=None, imle_att=None,
                return_attention_weights=None, return_masks=None, all_instrs=None, gi=None, e_proj=None):
        H = self.heads
        if gi is None:
            num_graphs = instruction.shape[0] if instruction is not None else int(batch[-1].item()) + 1
            gi = get_graph_index(edge_index, batch, num_graphs)
        if self.use_instr and self.concat_instr:
            x = ops.ConcatInstr.apply(x, instruction, gi)  # :153-154
        elif self.use_instr:
            x = ops.InstrGate.apply(x, instruction, gi)  # :156-157
        mask, edge_mask = None, None
        if self.mask.masking_threshold != 1.0:  # :161
            mask = self.mask.forward_fused(x, imle_att, gi)  # :166-168 (double gather handled in-kernel)
            edge_mask = ops.NodeMaskToEdgeMaskFn.apply(mask, gi)  # :169-171
        # :177,181 — lin_l and lin_r as ONE projection with the stacked weights [W_l; W_r] (their parameters
        # stay separate tensors; torch.cat's backward hands each its slice of the fused weight gradient)
        w_r, b_r = (self.lin_l.weight, self.lin_l.bias) if self.share_weights else (self.lin_r.weight, self.lin_r.bias)
        w_lr = torch.cat([self.lin_l.weight, w_r], dim=0)
        b_lr = torch.cat([self.lin_l.bias, b_r], dim=0) if self.lin_l.bias is not None else None
        xlr = ops.linear(x, w_lr, b_lr)
        if e_proj is None:
            e_proj = ops.linear(edge_attr, self.lin_edge.weight, None)  # :259 (inside message() in the reference)
        dbg = getattr(self, "debug_tensors", None)
        if dbg is not None:
            HC = H * self.out_channels
            for name, t in (("xg", x), ("xlr", xlr), ("e_proj", e_proj)):
                if t.requires_grad:
                    t.retain_grad()
                dbg[name] = t
            dbg["x_l"], dbg["x_r"] = xlr[:, :HC], xlr[:, HC:]
        out, alpha = ops.GatEdge.apply(xlr, e_proj, self.att, self.bias, edge_mask, gi, H,
                                       float(self.negative_slope))
        if isinstance(return_attention_weights, bool):
            return out, mask, (edge_index, alpha)
        return out, mask

    def __repr__(self):
        return f"{self.__class__.__name__}({self.in_channels}, {self.out_channels}, heads={self.heads})"
