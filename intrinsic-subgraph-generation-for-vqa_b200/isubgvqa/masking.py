"""Mirror of ISubGVQA/models/masking.py (MaskingModel, get_imle_samplers, get_aimle_samplers)."""
import math

import torch

from .. import lib as L
from .. import ops
from ..graph import get_graph_index
from .samplers import (AdaptiveTargetDistribution, EdgeSIMPLEBatched, GumbelDistribution, GumbelSampler,
                       IMLEScheme, TargetDistribution, aimle, imle)


class _TopKSelect(torch.nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.empty(1, in_channels))
        bound = 1.0 / math.sqrt(in_channels)
        torch.nn.init.uniform_(self.weight, -bound, bound)


class _TopKPoolingParams(torch.nn.Module):
    """Parameter container with the state_dict layout of torch_geometric 2.6.1 TopKPooling
    (`gate_top.select.weight` [1, D]); the reference builds it and never calls it (masking.py:89-90)."""

    def __init__(self, in_channels):
        super().__init__()
        self.select = _TopKSelect(in_channels)


def get_imle_samplers(sample_k, beta=10, alpha=1.0, tau=1.0, noise_scale=0.3, nb_samples=1, device=None,
                      noise_source="host"):
    """masking.py:214-245."""
    sched = IMLEScheme("edge_candid", sample_k, 1, 1)
    train = imle(sched.torch_sample_scheme, target_distribution=TargetDistribution(alpha=alpha, beta=beta),
                 noise_distribution=GumbelDistribution(0.0, noise_scale, device, noise_source),
                 nb_samples=nb_samples, input_noise_temperature=tau, target_noise_temperature=tau)
    val = imle(sched.torch_sample_scheme, target_distribution=None,
               noise_distribution=GumbelDistribution(0.0, noise_scale, device, noise_source),
               nb_samples=nb_samples, input_noise_temperature=tau if nb_samples > 1 else 0.0,
               target_noise_temperature=tau)
    return train, val


def get_aimle_samplers(sample_k, alpha=1.0, tau=1.0, noise_scale=0.3, nb_samples=1, device=None,
                       noise_source="host"):
    """masking.py:248-283."""
    sched = IMLEScheme("edge_candid", sample_k, 1, 1)
    train = aimle(sched.torch_sample_scheme,
                  target_distribution=AdaptiveTargetDistribution(initial_alpha=alpha, initial_beta=0.0),
                  noise_distribution=GumbelDistribution(0.0, noise_scale, device, noise_source),
                  nb_samples=nb_samples, theta_noise_temperature=tau, target_noise_temperature=tau,
                  symmetric_perturbation=True)
    val = aimle(sched.torch_sample_scheme, target_distribution=None,
                noise_distribution=GumbelDistribution(0.0, noise_scale, device, noise_source),
                nb_samples=nb_samples, theta_noise_temperature=1.0 if nb_samples > 1 else tau,
                target_noise_temperature=tau, symmetric_perturbation=True)
    return train, val


class MaskingModel(torch.nn.Module):
    """masking.py:53-199.  Same constructor, attributes (`masking_threshold` is read by the conv)
    and state_dict keys; forward(x, u, batch, edge_index, size=None, use_all_instrs=True) -> [N,1].

    `injected_noise` / `injected_dropout_mask` (attributes, default None) replace the random draws
    of the next forward — used by the parity tests to feed both sides the same randomness."""

    # The two 300x300 gate projections feed a DISCRETE decision (top-k) and, for Gumbel, a tau = 0.1 softmax
    # that amplifies logit perturbations 10x per round: they always run in strict fp32 (FFMA), whatever
    # the global projection mode is.  Cost: 2 x N x 300 x 300 FMAs per step.
    GATE_GEMM_MODE = 0

    def __init__(self, dim_nodes, dim_questions, masking_threshold=0.3, use_topk=False, sample_k=None,
                 sampler_type=None, nb_samples=1, alpha=1.0, beta=10.0, tau=1.0, noise_source="host"):
        super().__init__()
        self.use_topk = use_topk
        self.sample_k = sample_k
        self.sampler_type = sampler_type
        self.masking_threshold = int(masking_threshold) if masking_threshold > 1 else masking_threshold
        self.dim_nodes, self.dim_questions = dim_nodes, dim_questions
        self.gate_nn = torch.nn.Sequential(torch.nn.Linear(dim_questions, dim_questions), torch.nn.GELU(),
                                           torch.nn.Linear(dim_questions, 1))
        self.node_nn = torch.nn.Sequential(torch.nn.Linear(dim_nodes, dim_questions), torch.nn.GELU())
        self.ques_nn = torch.nn.Sequential(torch.nn.Linear(dim_questions, dim_questions), torch.nn.GELU())
        if use_topk:
            self.gate_top = _TopKPoolingParams(dim_questions)
        if sampler_type == "imle":
            self.sampler_train, self.sampler_val = get_imle_samplers(
                sample_k=sample_k, device=None, nb_samples=nb_samples, alpha=alpha, beta=beta, tau=tau,
                noise_source=noise_source)
        elif sampler_type == "aimle":
            self.sampler_train, self.sampler_val = get_aimle_samplers(
                sample_k=sample_k, device=None, nb_samples=nb_samples, alpha=alpha, tau=tau,
                noise_source=noise_source)
            # adaptive beta / gradient-norm EMA as a persistent buffer: checkpointed with the model (the reference
            # loses it on resume, target_aimle.py:101-109) and moved by .to(device)
            self.sampler_train.target.bind(self, "aimle_state")
        elif sampler_type == "simple":
            self.sampler = EdgeSIMPLEBatched(k=sample_k, device="cuda", policy="edge_candid")
        elif sampler_type == "gumbel":
            self.sampler = GumbelSampler(k=sample_k, policy="edge_candid", train_ensemble=1, val_ensemble=1)
        self.injected_noise = None
        self.injected_dropout_mask = None
        self.last_theta = None

    OPTIONAL_STATE_KEYS = ("aimle_state",)  # absent from checkpoints written by the reference

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        before = len(missing_keys)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)
        optional = {prefix + k for k in self.OPTIONAL_STATE_KEYS}
        missing_keys[before:] = [k for k in missing_keys[before:] if k not in optional]

    def reset_parameters(self):
        for seq in (self.gate_nn, self.node_nn, self.ques_nn):
            for m in seq:
                if hasattr(m, "reset_parameters"):
                    m.reset_parameters()

    # -- fused path used by MaskingGATv2Conv: u_graph is [B,D] (imle_att BEFORE the [batch] gather)
    def forward_fused(self, x, u_graph, gi):
        xn = ops.linear(x, self.node_nn[0].weight, self.node_nn[0].bias, L.ACT_GELU, self.GATE_GEMM_MODE)
        q = ops.linear(u_graph, self.ques_nn[0].weight, self.ques_nn[0].bias, L.ACT_GELU, self.GATE_GEMM_MODE)
        theta = ops.GateTheta.apply(xn, q, gi, True, self._keep_mask(x.shape[0], x.device))
        return self._sample(theta, gi)

    def forward(self, x, u, batch, edge_index, size=None, use_all_instrs=True):
        if use_all_instrs:
            raise NotImplementedError("use_all_instrs yields a 2-D dense tensor the reference's own IMLE wrapper "
                                      "rejects (wrapper.py:115-120); it is dead code on the path")
        x = x.unsqueeze(-1) if x.dim() == 1 else x
        num_graphs = size if size is not None else int(batch[-1].item()) + 1  # masking.py:135
        gi = get_graph_index(edge_index, batch, num_graphs)
        xn = ops.linear(x, self.node_nn[0].weight, self.node_nn[0].bias, L.ACT_GELU, self.GATE_GEMM_MODE)
        q = ops.linear(u, self.ques_nn[0].weight, self.ques_nn[0].bias, L.ACT_GELU, self.GATE_GEMM_MODE)  # u [N,D]
        theta = ops.GateTheta.apply(xn, q, gi, False, self._keep_mask(x.shape[0], x.device))  # q[batch] (:152)
        return self._sample(theta, gi)

    DROPOUT_P = 0.2  # masking.py:159 / :196 — dropout on theta, training only

    def _keep_mask(self, N, device):
        """The keep-mask F.dropout(theta, p=0.2) would draw — Bernoulli(1-p) / (1-p) on the device generator —
        or the injected one; it is applied inside the gate-logit kernel.  None in eval mode."""
        drop, self.injected_dropout_mask = self.injected_dropout_mask, None
        if not self.training:
            return None
        if drop is not None:
            return drop
        keep = torch.empty(N, 1, dtype=torch.float32, device=device).bernoulli_(1.0 - self.DROPOUT_P)
        return keep.mul_(1.0 / (1.0 - self.DROPOUT_P))

    def executor_spec(self, gi, device, N):
        """What the layer executor (csrc/executor.cu) needs to run this layer's sampler: the random draws of this
        step (noise, dropout keep-mask — injected ones take precedence) and the sampler's constants."""
        from .executor import SAMPLER_CODE

        noise, self.injected_noise = self.injected_noise, None
        keep = self._keep_mask(N, device)
        if keep is not None:
            keep = keep.to(device=device, dtype=torch.float32).contiguous()
        st = self.sampler_type
        spec = dict(code=SAMPLER_CODE[st], keep=keep)
        if st in ("imle", "aimle"):
            sampler = self.sampler_train if self.training else self.sampler_val
            nz = ops._noise2d(sampler._noise(gi.B, gi.nmax, device, noise), gi)
            spec.update(k=int(sampler.k), noise=nz, tau_in=sampler.tau_in, tau_tgt=sampler.tau_tgt)
            if st == "imle":
                spec.update(alpha=float(sampler.target.alpha), beta=float(sampler.target.beta))
            else:
                state, adaptive = sampler._state(device)
                spec.update(state=state, adaptive=adaptive)
        elif st == "gumbel":
            g = noise if noise is not None else self.sampler._gumbel(gi.B, gi.nmax, device)
            spec.update(k=max(1, min(int(self.sampler.k), gi.nmax)), noise=ops._noise2d(g, gi),
                        gumbel_tau=float(self.sampler.tau))
        else:  # simple
            npad = L.load().isg_simple_npad(gi.nmax)
            g = noise if noise is not None else self.sampler._gumbel(gi.B, npad, device)
            g = g.reshape(gi.B, -1)
            if g.shape[1] != npad:
                raise ValueError(f"SIMPLE noise has {g.shape[1]} slots per graph, expected n_pad = {npad}")
            spec.update(k=max(1, min(int(self.sampler.k), gi.nmax)), noise=g.to(torch.float32).contiguous())
        return spec

    def _sample(self, theta, gi):
        noise, self.injected_noise = self.injected_noise, None
        self.last_theta = theta
        if not self.use_topk:  # masking.py:195-198
            return (torch.sigmoid(theta) > 0.5).to(dtype=theta.dtype)
        if self.sampler_type in ("imle", "aimle"):
            sampler = self.sampler_train if self.training else self.sampler_val
            return sampler.ragged(theta, gi, noise)
        if self.sampler_type in ("gumbel", "simple"):
            return self.sampler.ragged(theta, gi, noise)
        raise ValueError(f"unknown sampler_type {self.sampler_type!r}")
