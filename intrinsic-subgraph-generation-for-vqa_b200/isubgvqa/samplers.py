"""Host-side mirror of ISubGVQA/sampling/methods/** on top of the CUDA sampler kernels
(csrc/sampler.cu).  Same class / function names, argument meaning and return shapes as the
reference, so models/masking.py-style call sites keep working:

  GumbelDistribution            <- sampling/methods/noise.py:71-89
  TargetDistribution            <- sampling/methods/target.py:22-48
  AdaptiveTargetDistribution    <- sampling/methods/target_aimle.py:87-162 (state lives on the device)
  select_from_edge_candidates   <- sampling/methods/deterministic_scheme.py:36-43
  IMLEScheme                    <- sampling/methods/imle_scheme.py:8-29
  imle / aimle                  <- sampling/methods/wrapper.py:16-176 / aimle.py:16-247
  GumbelSampler                 <- sampling/methods/gumbel_scheme.py:12-123
  EdgeSIMPLEBatched             <- sampling/methods/simple_scheme.py:24-162

Dense entry points take theta as [B, Nmax, 1] exactly like the reference; the fused ragged entry
points (`.ragged(theta [N,1], gi, ...)`) skip to_dense_batch altogether and are what
MaskingModel uses.  Only nb_samples == 1 is supported (the reference's own call site breaks for
S > 1: models/masking.py:169-173)."""
import torch

from .. import lib as L
from .. import ops


class _DenseIndex:
    """Uniform graph_ptr so a dense [B, Nmax] tensor can go through the ragged kernels."""

    def __init__(self, B, nmax, device):
        self.B, self.nmax = int(B), int(nmax)
        self.graph_ptr = torch.arange(0, (self.B + 1) * self.nmax, max(self.nmax, 1), dtype=torch.int32,
                                      device=device)[: self.B + 1].contiguous()
        if self.nmax == 0:
            self.graph_ptr = torch.zeros(self.B + 1, dtype=torch.int32, device=device)


class GumbelDistribution:
    """Gumbel(loc, scale) noise.  source='host' reproduces the reference bit-for-bit: samples are
    drawn by torch.distributions on the CPU generator and copied to the device (noise.py:86-89).
    source='device' draws on the CUDA generator instead (no H2D copy, different stream)."""

    def __init__(self, loc=0.0, scale=1.0, device="cpu", source="host"):
        self.loc, self._scale, self.device, self.source = loc, scale, device, source

    @property
    def scale(self):
        return self._scale

    @scale.setter
    def scale(self, value):
        self._scale = value

    def sample(self, shape, device=None):
        dev = self.device if device is None else device
        if self.source == "host":
            g = torch.distributions.gumbel.Gumbel(loc=self.loc, scale=self.scale)
            return g.sample(torch.Size(shape)).to(dev)
        e = torch.empty(tuple(shape), dtype=torch.float32, device=dev).exponential_()
        return self.loc - self.scale * torch.log(e)


class TargetDistribution:
    def __init__(self, alpha=1.0, beta=1.0, do_gradient_scaling=False, eps=1e-7):
        if do_gradient_scaling:
            raise NotImplementedError("do_gradient_scaling is never enabled on the ISubGVQA path")
        self.alpha, self.beta = alpha, beta

    def params(self, theta, dy):  # target.py:44-48 — kept for API parity; the kernel fuses it
        return self.alpha * theta - self.beta * dy


class AdaptiveTargetDistribution:
    """target_aimle.py:87-162.  The reference keeps beta / grad_norm as Python attributes and pays
    three .item() syncs per backward; here they live in an 8-double device vector that
    isg_aimle_bwd updates in place: [beta, grad_norm, previous_beta_update, alpha, beta_update_step,
    grad_norm_decay_rate, target_norm, beta_update_momentum].  When the distribution belongs to a
    MaskingModel the vector is that module's persistent buffer `aimle_state` (bind()), so it follows
    .to(device) and is part of state_dict(); standalone it is created lazily on first use."""

    def __init__(self, initial_alpha=1.0, initial_beta=1.0, initial_grad_norm=1.0, beta_update_step=0.0001,
                 beta_update_momentum=0.0, grad_norm_decay_rate=0.9, target_norm=1.0):
        self._init = [float(initial_beta), float(initial_grad_norm), 0.0, float(initial_alpha),
                      float(beta_update_step), float(grad_norm_decay_rate), float(target_norm),
                      float(beta_update_momentum)]
        self._state = None
        self._owner = None  # (module, buffer name)

    def bind(self, module, name):
        """Keep the state in `module`'s registered buffer `name` (float64[8])."""
        module.register_buffer(name, torch.tensor(self._init, dtype=torch.float64), persistent=True)
        self._owner = (module, name)
        self._state = None

    def _buffer(self):
        module, name = self._owner
        buf = getattr(module, name)
        if buf.dtype != torch.float64:  # module.float()/.half() casts floating buffers; the kernel needs doubles
            buf = buf.double()
            setattr(module, name, buf)
        return buf

    def state(self, device):
        if self._owner is not None:
            buf = self._buffer()
            if buf.device != torch.device(device):
                raise RuntimeError(f"AIMLE state lives on {buf.device} but theta is on {device}: move the module")
            return buf
        if self._state is None or self._state.device != torch.device(device):
            init = self._init if self._state is None else self._state.cpu().tolist()
            self._state = torch.tensor(init, dtype=torch.float64, device=device)
        return self._state

    def _get(self, i):
        if self._owner is not None:
            return float(self._buffer()[i].item())
        return self._init[i] if self._state is None else float(self._state[i].item())

    def _set(self, i, value):
        self._init[i] = float(value)
        if self._owner is not None:
            self._buffer()[i] = float(value)
        elif self._state is not None:
            self._state[i] = float(value)

    alpha = property(lambda self: self._get(3), lambda self, v: self._set(3, v))
    beta = property(lambda self: self._get(0), lambda self, v: self._set(0, v))
    grad_norm = property(lambda self: self._get(1), lambda self, v: self._set(1, v))
    previous_beta_update = property(lambda self: self._get(2), lambda self, v: self._set(2, v))

    def state_dict(self):
        """Not part of the reference (its AIMLE state is lost on resume); offered so
        callers of a standalone distribution can checkpoint it."""
        if self._owner is not None:
            return {"state": self._buffer().cpu().tolist()}
        return {"state": self._init if self._state is None else self._state.cpu().tolist()}

    def load_state_dict(self, sd):
        self._init = list(sd["state"])
        if self._owner is not None:
            self._buffer().copy_(torch.tensor(self._init, dtype=torch.float64))
        self._state = None


def select_from_edge_candidates(scores, k):
    """deterministic_scheme.py:36-43 on the device.  scores [B, Nmax, 1]."""
    B, nmax, ens = scores.shape
    if ens != 1:
        raise NotImplementedError("ensemble > 1 is not used on the ISubGVQA path")
    if k >= nmax:
        return scores.new_ones(scores.shape)
    di = _DenseIndex(B, nmax, scores.device)
    flat = scores.detach().to(torch.float32).reshape(B * nmax, 1).contiguous()
    mask = torch.empty_like(flat)
    zd = torch.empty(B, nmax, dtype=torch.float32, device=scores.device)
    L.call("isg_topk_mask_fwd", L.ptr(flat), None, L.ptr(di.graph_ptr), B, nmax, int(k), 0.0, L.ptr(mask),
                                       L.ptr(zd), L.stream())
    return zd.view(B, nmax, 1)


class IMLEScheme:
    def __init__(self, imle_sample_policy, sample_k, train_ensemble, val_ensemble):
        if imle_sample_policy != "edge_candid":
            raise NotImplementedError("only the 'edge_candid' policy is on the ISubGVQA path (masking.py:215-220)")
        self.policy, self.k = imle_sample_policy, sample_k
        self.adj = None
        self.train_ensemble, self.val_ensemble = train_ensemble, val_ensemble

    @torch.no_grad()
    def torch_sample_scheme(self, logits):
        return select_from_edge_candidates(logits.detach(), self.k), None


def _solver_k(function):
    k = getattr(function, "isg_topk_k", None)
    if k is None:
        bound = getattr(function, "__self__", None)
        if isinstance(bound, IMLEScheme):
            k = bound.k
    if k is None:
        raise NotImplementedError(
            "isg_b200 fuses the MAP solver into the kernel: only top-k ('edge_candid') solvers are supported; "
            "set `function.isg_topk_k = k` or pass IMLEScheme(...).torch_sample_scheme")
    return int(k)


class _ImleSampler:
    def __init__(self, k, target, noise_distribution, nb_samples, tau_in, tau_tgt):
        if nb_samples != 1:
            raise NotImplementedError("nb_samples must be 1 (the reference call site requires it too)")
        self.k, self.target, self.noise_distribution = k, target, noise_distribution
        self.tau_in, self.tau_tgt = float(tau_in), float(tau_tgt)

    def _noise(self, B, nmax, device, noise):
        if noise is not None:
            return noise
        if self.noise_distribution is None:
            return None
        return self.noise_distribution.sample((B, 1, nmax, 1), device=device)

    def ragged(self, theta, gi, noise=None):
        """theta [N,1] -> mask [N,1]; noise [B,1,Nmax,1] (injected) or drawn from noise_distribution."""
        noise = self._noise(gi.B, gi.nmax, theta.device, noise)
        return ops.TopkImle.apply(theta, noise, gi, self.k, float(self.target.alpha), float(self.target.beta),
                                  self.tau_in, self.tau_tgt)

    def __call__(self, input, noise=None):
        """Reference shape contract (wrapper.py:112-121): [B,Nmax,1] -> ([1,B,Nmax,1], None)."""
        B, nmax = input.shape[0], input.shape[1]
        di = _DenseIndex(B, nmax, input.device)
        z = self.ragged(input.reshape(B * nmax, 1), di, noise)
        return z.view(1, B, nmax, 1), None


def imle(function=None, target_distribution=None, noise_distribution=None, nb_samples=1,
         input_noise_temperature=1.0, target_noise_temperature=1.0):
    """wrapper.py:16-176.  Usable directly or as a decorator, like the reference."""
    if target_distribution is None:
        target_distribution = TargetDistribution(alpha=1.0, beta=1.0)
    if function is None:
        return lambda f: imle(f, target_distribution, noise_distribution, nb_samples, input_noise_temperature,
                              target_noise_temperature)
    return _ImleSampler(_solver_k(function), target_distribution, noise_distribution, nb_samples,
                        input_noise_temperature, target_noise_temperature)


class _AimleSampler(_ImleSampler):
    def __init__(self, k, target, noise_distribution, nb_samples, tau_in, tau_tgt, symmetric):
        super().__init__(k, target, noise_distribution, nb_samples, tau_in, tau_tgt)
        if not symmetric:
            raise NotImplementedError("ISubGVQA always uses symmetric_perturbation=True (masking.py:266,282)")
        self._fixed_state = None

    def _state(self, device):
        if isinstance(self.target, AdaptiveTargetDistribution):
            return self.target.state(device), True
        if self._fixed_state is None or self._fixed_state.device != torch.device(device):
            self._fixed_state = torch.tensor(
                [float(self.target.beta), 1.0, 0.0, float(self.target.alpha), 0.0, 0.0, 1.0, 0.0],
                dtype=torch.float64, device=device)
        return self._fixed_state, False

    def ragged(self, theta, gi, noise=None):
        noise = self._noise(gi.B, gi.nmax, theta.device, noise)
        state, adaptive = self._state(theta.device)
        return ops.TopkAimle.apply(theta, noise, gi, self.k, state, adaptive, self.tau_in, self.tau_tgt)

    def __call__(self, theta, noise=None):
        """aimle.py:138 returns z as [B*S, Nmax, 1]."""
        B, nmax = theta.shape[0], theta.shape[1]
        di = _DenseIndex(B, nmax, theta.device)
        return self.ragged(theta.reshape(B * nmax, 1), di, noise).view(B, nmax, 1)


def aimle(function=None, target_distribution=None, noise_distribution=None, nb_samples=1, nb_marginal_samples=1,
          theta_noise_temperature=1.0, target_noise_temperature=1.0, symmetric_perturbation=False,
          _is_minimization=False):
    """aimle.py:16-247."""
    if nb_marginal_samples != 1 or _is_minimization:
        raise NotImplementedError("nb_marginal_samples != 1 / minimisation are not on the ISubGVQA path")
    if target_distribution is None:
        target_distribution = TargetDistribution(alpha=1.0, beta=1.0)
    if function is None:
        return lambda f: aimle(f, target_distribution, noise_distribution, nb_samples, nb_marginal_samples,
                               theta_noise_temperature, target_noise_temperature, symmetric_perturbation)
    return _AimleSampler(_solver_k(function), target_distribution, noise_distribution, nb_samples,
                         theta_noise_temperature, target_noise_temperature, symmetric_perturbation)


class GumbelSampler(torch.nn.Module):
    """gumbel_scheme.py:12-107 ('edge_candid').  forward(scores [B,Nmax,1], train) -> ([1,B,Nmax,1], None)."""

    def __init__(self, k, train_ensemble, val_ensemble, tau=0.1, hard=True, policy=None):
        super().__init__()
        if policy != "edge_candid" or not hard or train_ensemble != 1 or val_ensemble != 1:
            raise NotImplementedError("ISubGVQA builds GumbelSampler(policy='edge_candid', ensembles 1, hard=True)")
        self.policy, self.k, self.hard, self.tau = policy, k, hard, tau
        self.adj = None
        self.train_ensemble, self.val_ensemble = train_ensemble, val_ensemble

    @staticmethod
    def _gumbel(B, nmax, device):
        e = torch.empty(B, nmax, dtype=torch.float32, device=device).exponential_()
        return -torch.log(e)  # Gumbel(0,1) on the device generator, as gumbel_scheme.py:65-70 does

    def ragged(self, theta, gi, gumbel=None):
        if gumbel is None:
            gumbel = self._gumbel(gi.B, gi.nmax, theta.device)
        return ops.GumbelTopk.apply(theta, gumbel, gi, int(self.k), float(self.tau))

    def forward(self, scores, train=True, gumbel=None):
        B, nmax, ens = scores.shape
        if ens != 1:
            raise NotImplementedError("ensemble > 1 is not used on the ISubGVQA path")
        di = _DenseIndex(B, nmax, scores.device)
        out = self.ragged(scores.reshape(B * nmax, 1), di, gumbel)
        return out.view(1, B, nmax, 1), None

    @torch.no_grad()
    def validation(self, scores):
        return select_from_edge_candidates(scores, self.k)[None], None


class EdgeSIMPLEBatched(torch.nn.Module):
    """simple_scheme.py:24-162 — SIMPLE exact k-subset marginals ('edge_candid', ensembles 1, no logits
    activation).  forward(scores [B,Nmax,1], train) -> (mask [1,B,Nmax,1], marginals [B,Nmax,1]).
    The SDD circuit is evaluated by csrc/simple.cu straight from its closed form (a balanced tree of
    count nodes), so no ./simple_configs/*.pkl is written or read (simple.py:114-122 does both)."""

    def __init__(self, k, device, policy, val_ensemble=1, train_ensemble=1, logits_activation=None):
        super().__init__()
        if policy != "edge_candid" or val_ensemble != 1 or train_ensemble != 1 or \
                logits_activation not in (None, "None"):
            raise NotImplementedError("ISubGVQA builds EdgeSIMPLEBatched(policy='edge_candid'), ensembles 1, no "
                                      "logits activation (models/masking.py:109-113)")
        self.k, self.device, self.policy = k, device, policy
        self.layer_configs = dict()
        self.adj = None
        self.val_ensemble, self.train_ensemble = val_ensemble, train_ensemble
        self.logits_activation = logits_activation

    @staticmethod
    def _gumbel(B, npad, device):
        u = torch.rand(B, npad, dtype=torch.float32, device=device)  # simple.py:94-96 (device generator)
        return -torch.log(-torch.log(u))

    def ragged(self, theta, gi, gumbel=None):
        if gumbel is None:
            gumbel = self._gumbel(gi.B, L.load().isg_simple_npad(gi.nmax), theta.device)
        mask, _marg = ops.SimpleTopk.apply(theta, gumbel, gi, int(self.k))
        return mask

    def forward(self, scores, train=True, gumbel=None):
        B, nmax, ens = scores.shape
        if ens != 1:
            raise NotImplementedError("ensemble > 1 is not used on the ISubGVQA path")
        di = _DenseIndex(B, nmax, scores.device)
        if gumbel is None:
            gumbel = self._gumbel(B, L.load().isg_simple_npad(nmax), scores.device)
        mask, marg = ops.SimpleTopk.apply(scores.reshape(B * nmax, 1), gumbel, di, int(self.k))
        return mask.view(1, B, nmax, 1), marg.view(B, nmax, 1)

    @torch.no_grad()
    def validation(self, scores):
        return select_from_edge_candidates(scores, self.k)[None], None
