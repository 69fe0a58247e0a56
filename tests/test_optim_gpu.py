"""SURVEY.md section 8 row f4 — FusedClipAdam against the reference's optimizer tail
(training/train_epoch.py:111-118: GradScaler.unscale_ + clip_grad_norm_(2.0) + GradScaler.step(torch.optim.Adam) +
update) on identical gradients, including a step with a non-finite gradient (both skip it and halve the scale) and
a checkpoint round trip between the two optimizers (main.py:132-133, training/train_loop.py:91)."""
import copy
import warnings

import pytest
import torch

from isg_b200 import synth
from isg_b200.optim import FusedClipAdam

pytestmark = pytest.mark.gpu


def _params(seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = [(1200, 300), (1200,), (1, 4, 300), (300, 600), (5,), (70000,), (2577, 512), (1,)]
    return [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in shapes] + \
           [torch.nn.Parameter(torch.zeros(3, 3).cuda())]  # never receives a gradient


def _grads(params, step, scale, poison=False):
    g = torch.Generator().manual_seed(100 + step)
    out = []
    for i, p in enumerate(params[:-1]):
        t = torch.randn(p.shape, generator=g).cuda() * (3.0 if i % 2 else 0.01) * scale
        if poison and i == 3:
            t[0, 0] = float("inf")
        out.append(t)
    return out + [None]


def test_matches_gradscaler_clip_adam_sequence():
    ref_p, our_p = _params(), _params()
    ref_opt = torch.optim.Adam(ref_p, lr=3e-3)
    our_opt = FusedClipAdam(our_p, lr=3e-3, max_norm=2.0)
    ref_sc, our_sc = torch.amp.GradScaler("cuda", init_scale=1024.0), torch.amp.GradScaler("cuda", init_scale=1024.0)
    for step in range(6):
        poison = step == 2
        for params, opt, sc, fused in ((ref_p, ref_opt, ref_sc, False), (our_p, our_opt, our_sc, True)):
            scale = float(sc.get_scale())
            for p, g in zip(params, _grads(params, step, scale, poison)):
                p.grad = g
            sc._lazy_init_scale_growth_tracker(torch.device("cuda")) if sc._scale is None else None
            if fused:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore", FutureWarning)
                    sc.step(opt)
            else:
                sc.unscale_(opt)
                torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], max_norm=2.0)
                sc.step(opt)
            sc.update()
        assert float(ref_sc.get_scale()) == float(our_sc.get_scale()), step
        for a, b in zip(our_p, ref_p):
            err = float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)
            assert err <= 2e-6, (step, tuple(a.shape), err)
    assert float(our_sc.get_scale()) == 512.0  # the poisoned step halved it once
    assert float(our_opt._dev["state"][3]) == 5.0  # five steps taken, one skipped


def test_without_scaler_and_state_dict_interchange_with_torch_adam():
    ref_p, our_p = _params(1), _params(1)
    ref_opt = torch.optim.Adam(ref_p, lr=1e-2)
    our_opt = FusedClipAdam(our_p, lr=1e-2, max_norm=2.0)
    for step in range(3):
        for params, opt, fused in ((ref_p, ref_opt, False), (our_p, our_opt, True)):
            for p, g in zip(params, _grads(params, step, 1.0)):
                p.grad = g
            if not fused:
                torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], max_norm=2.0)
            opt.step()
    total = torch.sqrt(sum((g.float() ** 2).sum() for g in _grads(our_p, 2, 1.0) if g is not None))
    assert abs(float(our_opt.last_grad_norm) - float(total)) <= 1e-5 * float(total)
    # continue the fused run in a torch.optim.Adam loaded from its state_dict, and vice versa
    sd_ours, sd_ref = copy.deepcopy(our_opt.state_dict()), copy.deepcopy(ref_opt.state_dict())
    assert float(sd_ours["state"][0]["step"]) == 3.0
    cont_ref = torch.optim.Adam(our_p, lr=1e-2)
    cont_ref.load_state_dict(sd_ours)
    cont_ours = FusedClipAdam(ref_p, lr=1e-2, max_norm=2.0)
    cont_ours.load_state_dict(sd_ref)
    for params, opt, fused in ((our_p, cont_ref, False), (ref_p, cont_ours, True)):
        for p, g in zip(params, _grads(params, 3, 1.0)):
            p.grad = g
        if not fused:
            torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], max_norm=2.0)
        opt.step()
    for a, b in zip(our_p, ref_p):
        assert float((a - b).abs().max()) <= 2e-6 * max(float(b.abs().max()), 1e-30)


def test_trains_mgat_end_to_end_without_host_sync():
    """One real training step of the drop-in MGAT with the fused optimizer: parameters move, unused ones don't."""
    from isg_b200.isubgvqa import MGAT

    b = synth.make_batch(8, mean_nodes=10, mean_edges=50, seed=4)
    m = MGAT(channels=300, num_ins=4, heads=4, use_instr=True, masking_thresholds=[1.0, 1.0, 1.0, 0.1],
             use_topk=True, interpretable_mode=False, sampler_type="imle", sample_k=2)
    m.load_state_dict(synth.make_state_dict(seed=4))
    m.cuda().train()
    before = {k: v.detach().clone() for k, v in m.named_parameters()}
    opt = FusedClipAdam(m.parameters(), lr=1e-3, max_norm=2.0)
    sc = torch.amp.GradScaler("cuda")
    h, _, _, _ = m(*[b[k].cuda() for k in ("x", "edge_index", "instr_vectors", "global_language_feats", "edge_attr",
                                           "batch")])
    opt.zero_grad()
    sc.scale((h * h).mean()).backward()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", FutureWarning)
        sc.step(opt)
    sc.update()
    moved = [k for k, v in m.named_parameters() if not torch.equal(v, before[k])]
    assert "convs.0.lin_l.weight" in moved and "convs.3.mask.node_nn.0.weight" in moved
    assert "node_logits.0.weight" not in moved and "convs.0.mask.gate_nn.0.weight" not in moved
    assert torch.isfinite(opt.last_grad_norm)
