"""SURVEY.md section 8 row f4 — the optimizer tail of the reference's training step
(training/train_epoch.py:111-118) as three kernel launches and no host synchronisation:

    gradscaler.scale(loss).backward()
    gradscaler.unscale_(optimizer)                                   |
    torch.nn.utils.clip_grad_norm_(model_params, max_norm=2.0)       |  ->  optimizer.step(grad_scaler=gradscaler)
    gradscaler.step(optimizer)     # torch.optim.Adam, main.py:106   |
    gradscaler.update()                                              (unchanged)

`FusedClipAdam` subclasses torch.optim.Adam: same constructor, same state layout (`step`, `exp_avg`,
`exp_avg_sq` per parameter), so `optimizer.load_state_dict(checkpoint["optimizer"])` (main.py:132-133) and
`optimizer.state_dict()` (training/train_loop.py:91) interoperate with checkpoints written by the reference.
Only what the reference uses is supported: no weight decay, no amsgrad, fp32 CUDA parameters."""
import ctypes

import numpy as np
import torch

from . import lib as L


class FusedClipAdam(torch.optim.Adam):
    _step_supports_amp_scaling = True  # GradScaler.step hands us the scaler instead of unscaling / syncing itself

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, max_norm=2.0):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False)
        self.max_norm = float(max_norm)
        self._dev = None  # device scalars: state[4] (norm, clip coefficient, found_inf, steps), nonfinite flag

    # ---- device-side bookkeeping
    def _device_state(self, device):
        if self._dev is None or self._dev["state"].device != device:
            steps = 0.0
            for group in self.param_groups:
                for p in group["params"]:
                    st = self.state.get(p)
                    if st and "step" in st:
                        steps = max(steps, float(st["step"]))
            self._dev = dict(state=torch.tensor([0.0, 1.0, 0.0, steps], dtype=torch.float32, device=device),
                             nonfinite=torch.zeros(1, dtype=torch.int32, device=device), partial=None)
        return self._dev

    @property
    def last_grad_norm(self):
        """Total (unscaled, pre-clip) gradient norm of the last step — a device scalar, no sync until read."""
        return None if self._dev is None else self._dev["state"][0]

    @torch.no_grad()
    def step(self, closure=None, grad_scaler=None):
        if closure is not None:
            raise NotImplementedError("FusedClipAdam.step does not take a closure")
        lib = L.load()
        items = []
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise TypeError("FusedClipAdam handles fp32 CUDA parameters only (no CPU fallback)")
                st = self.state[p]
                if not st:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                items.append((group, p, g, st))
        if not items:
            return None
        dev = items[0][1].device
        ds = self._device_state(dev)
        inv_scale = found_inf_out = None
        if grad_scaler is not None and grad_scaler.is_enabled():
            scale = grad_scaler._get_scale_async()
            inv_scale = scale.double().reciprocal().float()
            # the per-device found_inf tensor GradScaler.update() reads
            opt_state = grad_scaler._per_optimizer_states[id(self)]
            found_inf_out = opt_state["found_inf_per_device"].setdefault(dev, torch.zeros(1, dtype=torch.float32, device=dev))
        st_ = L.stream()
        maxt = lib.isg_opt_max_tensors()
        groups = [items[i:i + maxt] for i in range(0, len(items), maxt)]
        metas, total_blocks = [], 0
        for grp in groups:
            n = len(grp)
            numel = np.array([it[1].numel() for it in grp], dtype=np.int64)
            arr = lambda f: np.array([f(it) for it in grp], dtype=np.uint64)  # noqa: E731
            meta = dict(n=n, numel=numel, p=arr(lambda it: it[1].data_ptr()), g=arr(lambda it: it[2].data_ptr()),
                        m=arr(lambda it: it[3]["exp_avg"].data_ptr()), v=arr(lambda it: it[3]["exp_avg_sq"].data_ptr()),
                        off=total_blocks)
            total_blocks += int(lib.isg_opt_blocks(numel.ctypes.data_as(ctypes.c_void_p), n))
            metas.append(meta)
        if ds["partial"] is None or ds["partial"].numel() < total_blocks:
            ds["partial"] = torch.empty(max(total_blocks, 1), dtype=torch.float64, device=dev)
        ds["nonfinite"].zero_()
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        for m in metas:
            L.call("isg_grad_sq_partials", vp(m["g"]), vp(m["numel"]), m["n"], L.ptr(inv_scale), L.ptr(ds["partial"]),
                   m["off"], L.ptr(ds["nonfinite"]), st_)
        L.call("isg_clip_finalize", L.ptr(ds["partial"]), total_blocks, L.ptr(ds["nonfinite"]), self.max_norm,
               L.ptr(ds["state"]), L.ptr(found_inf_out), st_)
        for grp, m in zip(groups, metas):
            group = grp[0][0]
            if any(it[0] is not group for it in grp):
                raise NotImplementedError("parameter groups with different hyper-parameters inside one launch group")
            lr = group["lr"]
            lr_dev = lr if isinstance(lr, torch.Tensor) else None
            b1, b2 = group["betas"]
            L.call("isg_adam_update", vp(m["p"]), vp(m["g"]), vp(m["m"]), vp(m["v"]), vp(m["numel"]), m["n"],
                   L.ptr(inv_scale), L.ptr(ds["state"]), L.ptr(lr_dev), 0.0 if lr_dev is not None else float(lr),
                   float(b1), float(b2), float(group["eps"]), st_)
        return None

    def state_dict(self):
        """torch.optim.Adam's layout; the per-parameter `step` entries are refreshed from the device counter (one
        read, only when a checkpoint is written)."""
        if self._dev is not None:
            steps = float(self._dev["state"][3].item())
            for st in self.state.values():
                if "step" in st:
                    st["step"] = torch.tensor(steps, dtype=torch.float32)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._dev = None  # re-read the step count from the loaded state on the next step
