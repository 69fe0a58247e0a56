"""Mirror of ISubGVQA/models/mgat.py (MGAT)."""
import torch

import os

from .. import lib as L
from .. import ops
from ..graph import get_graph_index
from . import executor
from .mgat_v2_conv import MaskingGATv2Conv

# 1 (default): MGAT.forward runs the layer executor (one C call per layer and direction, isubgvqa/executor.py);
# 0: the per-operator path below (one autograd Function per kernel family), kept as the cross-check.
_USE_EXECUTOR = os.environ.get("ISG_EXECUTOR", "1") != "0"


def set_executor(on):
    global _USE_EXECUTOR
    _USE_EXECUTOR = bool(on)


class GraphNormParams(torch.nn.Module):
    """Parameter container with torch_geometric.nn.norm.GraphNorm's layout (weight, bias, mean_scale)."""

    def __init__(self, in_channels, eps=1e-5):
        super().__init__()
        self.in_channels, self.eps = in_channels, eps
        self.weight = torch.nn.Parameter(torch.ones(in_channels))
        self.bias = torch.nn.Parameter(torch.zeros(in_channels))
        self.mean_scale = torch.nn.Parameter(torch.ones(in_channels))

    def reset_parameters(self):
        torch.nn.init.ones_(self.weight)
        torch.nn.init.zeros_(self.bias)
        torch.nn.init.ones_(self.mean_scale)


class MGAT(torch.nn.Module):
    """mgat.py:8-184: num_ins x [MaskingGATv2Conv -> x_proj (Linear GELU Linear GELU) -> scatter-SDPA ->
    GraphNorm -> residual].  Same constructor / forward signature / 4-tuple return / state_dict."""

    def __init__(self, channels, num_ins, dropout=0.0, heads=4, use_instr=False, masking_thresholds=None,
                 use_topk=False, interpretable_mode=True, concat_instr=False, use_all_instrs=False,
                 use_global_mask=False, node_classification=False, sampler_type=None, sample_k=None, nb_samples=1,
                 alpha=1.0, beta=10.0, tau=1.0):
        super().__init__()
        if not use_instr:
            raise NotImplementedError("the reference MGAT itself only works with use_instr=True (mgat.py:46-53)")
        self.masking_thresholds = masking_thresholds
        self.use_global_mask, self.node_classification = use_global_mask, node_classification
        self.heads, self.use_instr, self.use_topk = heads, use_instr, use_topk
        self.interpretable_mode, self.use_all_instrs = interpretable_mode, use_all_instrs
        self.in_channels = channels * 2 if concat_instr else channels
        self.convs = torch.nn.ModuleList([
            MaskingGATv2Conv(in_channels=self.in_channels, out_channels=channels, heads=heads, edge_dim=channels,
                             masking_threshold=self.masking_thresholds[i], add_self_loops=False, use_instr=True,
                             use_topk=use_topk, concat_instr=concat_instr, use_all_instrs=use_all_instrs,
                             sampler_type=sampler_type, sample_k=sample_k, nb_samples=nb_samples, alpha=alpha,
                             beta=beta, tau=tau) for i in range(num_ins)])
        self.x_proj = torch.nn.ModuleList([
            torch.nn.Sequential(torch.nn.Linear(heads * channels, channels * int(heads / 2)), torch.nn.GELU(),
                                torch.nn.Linear(channels * int(heads / 2), channels), torch.nn.GELU())
            for _ in range(num_ins)])
        self.bns = torch.nn.ModuleList([GraphNormParams(channels) for _ in range(num_ins)])
        self.dropout = dropout
        self.node_logits = torch.nn.Sequential(torch.nn.Linear(channels, 512), torch.nn.GELU(),
                                               torch.nn.Linear(512, 2577))

    def reset_parameters(self):
        for conv in self.convs:
            conv.reset_parameters()
        for bn in self.bns:
            bn.reset_parameters()

    def forward(self, x, edge_index, instr_vectors, global_language_feats, edge_attr, batch, return_masks=False,
                explainer=False, explainer_stage=False, expl_bypass_x=False):
        L.require_cuda(x, edge_index, batch, edge_attr)
        ops.join_side_stream()  # no-op unless a previous backward pass was interrupted before its join ran
        params = self.__dict__.get("_param_list")
        if params is None:  # nn.Module.parameters() walks the module tree (~0.4 ms per call); the set is static
            params = self.__dict__["_param_list"] = list(self.parameters())
        ops.allow_side_stream(all(p.grad is None for p in params))
        gi = get_graph_index(edge_index, batch, instr_vectors.shape[1])
        if _USE_EXECUTOR and executor.supported(self, explainer, params):
            masked = [c for c in self.convs if c.mask.masking_threshold != 1.0]
            if masked and gi.B > x.shape[0]:  # batch[batch[n]] (quirk Q1): the reference raises here too
                raise IndexError(f"double gather needs num_graphs <= num_nodes, got B={gi.B}, N={x.shape[0]}")
            if not any(c.mask.sampler_type == "simple" for c in masked) or gi.nmax > 1:
                for conv in self.convs:
                    executor.ensure_stacked(conv)
                specs = [c.mask.executor_spec(gi, x.device, x.shape[0]) if c.mask.masking_threshold != 1.0 else None
                         for c in self.convs]
                h, mask = executor.MgatFunction.apply(self, gi, specs, ops.gemm_mode(), x, edge_attr, instr_vectors,
                                                      global_language_feats, *executor.flat_params(self))
                return h, mask, [], []
        if ops.gemm_mode() == executor.BF16_MODE:
            raise NotImplementedError("the bf16 configuration (gemm mode 3) runs through the layer executor only; this "
                                      "model / call is outside what the executor covers (executor.supported)")
        h = x
        mask = None
        if self.use_global_mask:
            global_mask = torch.ones((h.size(0), 1), device=h.device, dtype=h.dtype)
        for i, conv in enumerate(self.convs):
            ins = instr_vectors[i]
            if explainer:
                h = expl_bypass_x if (explainer_stage - 1) == i else h
            conv_res, mask, _edge_att = conv(x=h, edge_index=edge_index, edge_attr=edge_attr, instruction=ins,
                                             batch=batch, return_masks=return_masks, return_attention_weights=True,
                                             imle_att=global_language_feats, all_instrs=instr_vectors, gi=gi)
            dbg = getattr(self, "debug_tensors", None)
            if dbg is not None:
                conv_res.retain_grad()
                dbg[f"conv_out.{i}"] = conv_res
            p = self.x_proj[i]
            conv_res = ops.linear(conv_res, p[0].weight, p[0].bias, L.ACT_GELU)
            conv_res = ops.linear(conv_res, p[2].weight, p[2].bias, L.ACT_GELU)
            if dbg is not None:
                conv_res.retain_grad()
                dbg[f"proj.{i}"] = conv_res
            if self.use_global_mask:
                global_mask = mask * global_mask
            bn = self.bns[i]
            h = ops.SdpaGraphNormResidual.apply(conv_res, ins, h, bn.weight, bn.bias, bn.mean_scale, gi,
                                                float(bn.eps))
            if dbg is not None:
                h.retain_grad()
                dbg[f"h.{i}"] = h
            if self.use_global_mask:
                h = global_mask * h
            elif self.interpretable_mode and mask is not None:
                h = mask * h
        return h, mask, [], []
