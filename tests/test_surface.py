"""CPU: the drop-in keeps the reference's call signatures.  For every class and function the hot path's callers touch
(models/isubgvqa.py:159-176,267-287; models/masking.py:21-75), the reference's parameters appear in the drop-in at the
same positions with the same names, kinds and defaults; anything the drop-in adds comes AFTER them and is optional
(e.g. `gi=` a prebuilt graph index, `noise_source=` the RNG placement).  Compared against the unmodified reference,
imported live — skipped where /root/reference is absent."""
import importlib
import inspect
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import reference_loader as rl  # noqa: E402

pytestmark = pytest.mark.skipif(not rl.available(), reason="/root/reference is not present on this machine")

CLASSES = [  # (reference module, class, drop-in module attribute, methods)
    ("ISubGVQA.models.mgat", "MGAT", "mgat", ("__init__", "forward")),
    ("ISubGVQA.models.mgat_v2_conv", "MaskingGATv2Conv", "mgat_v2_conv", ("__init__", "forward")),
    ("ISubGVQA.models.masking", "MaskingModel", "masking", ("__init__", "forward")),
    ("ISubGVQA.models.att_pooling", "GlobalAttention", "att_pooling", ("__init__", "forward")),
    ("ISubGVQA.sampling.methods.gumbel_scheme", "GumbelSampler", "samplers", ("__init__", "forward")),
    ("ISubGVQA.sampling.methods.simple_scheme", "EdgeSIMPLEBatched", "samplers", ("__init__", "forward")),
    ("ISubGVQA.sampling.methods.imle_scheme", "IMLEScheme", "samplers", ("__init__",)),
    ("ISubGVQA.sampling.methods.noise", "GumbelDistribution", "samplers", ("__init__", "sample")),
    ("ISubGVQA.sampling.methods.target", "TargetDistribution", "samplers", ("__init__", "params")),
    ("ISubGVQA.sampling.methods.target_aimle", "AdaptiveTargetDistribution", "samplers", ("__init__",)),
]
FUNCTIONS = [
    ("ISubGVQA.models.masking", "get_imle_samplers", "masking"),
    ("ISubGVQA.models.masking", "get_aimle_samplers", "masking"),
    ("ISubGVQA.sampling.methods.wrapper", "imle", "samplers"),
    ("ISubGVQA.sampling.methods.aimle", "aimle", "samplers"),
    ("ISubGVQA.sampling.methods.deterministic_scheme", "select_from_edge_candidates", "samplers"),
]


def _params(fn):
    return [(p.name, p.kind, p.default) for p in inspect.signature(fn).parameters.values()]


def _check(ref_fn, our_fn, what):
    ref, ours = _params(ref_fn), _params(our_fn)
    assert len(ours) >= len(ref), f"{what}: the drop-in takes fewer parameters than the reference"
    for (rn, rk, rd), (on, ok, od) in zip(ref, ours):
        assert rn == on and rk == ok, f"{what}: parameter {rn!r} ({rk.name}) became {on!r} ({ok.name})"
        assert (rd is inspect.Parameter.empty) == (od is inspect.Parameter.empty) and (rd is inspect.Parameter.empty or rd == od), \
            f"{what}: default of {rn!r} is {od!r}, the reference has {rd!r}"
    for name, kind, default in ours[len(ref):]:
        assert default is not inspect.Parameter.empty or kind in (inspect.Parameter.VAR_POSITIONAL, inspect.Parameter.VAR_KEYWORD), \
            f"{what}: added parameter {name!r} must be optional"


@pytest.fixture(scope="module")
def reference():
    with rl.scratch_cwd():
        return rl.load()


@pytest.mark.parametrize("ref_mod,cls,attr,methods", CLASSES, ids=[c[1] for c in CLASSES])
def test_class_signatures_extend_the_reference(reference, ref_mod, cls, attr, methods):
    from isg_b200 import isubgvqa as ours

    rc, oc = getattr(importlib.import_module(ref_mod), cls), getattr(getattr(ours, attr), cls)
    for m in methods:
        _check(getattr(rc, m), getattr(oc, m), f"{cls}.{m}")


@pytest.mark.parametrize("ref_mod,fn,attr", FUNCTIONS, ids=[f[1] for f in FUNCTIONS])
def test_function_signatures_extend_the_reference(reference, ref_mod, fn, attr):
    from isg_b200 import isubgvqa as ours

    _check(getattr(importlib.import_module(ref_mod), fn), getattr(getattr(ours, attr), fn), fn)


def test_forward_returns_the_reference_arity(reference):
    """MGAT.forward returns a 4-tuple like the reference's (models/mgat.py:179-184: h, mask, node_logits_layers,
    hidden_states) — read off the syntax trees of both (no device needed)."""
    import ast
    import textwrap

    from isg_b200.isubgvqa import mgat as ours

    def arities(fn):
        tree = ast.parse(textwrap.dedent(inspect.getsource(fn)))
        top = tree.body[0]
        return [len(n.value.elts) if isinstance(n.value, ast.Tuple) else 1
                for n in ast.walk(top) if isinstance(n, ast.Return) and n.value is not None]

    assert arities(importlib.import_module("ISubGVQA.models.mgat").MGAT.forward) == [4]
    ours_ret = arities(ours.MGAT.forward)
    assert ours_ret and set(ours_ret) == {4}, ours_ret


def test_fresh_modules_are_initialised_like_the_reference(reference):
    """Training from scratch starts from the same distribution: every parameter of a freshly built MGAT has the
    reference's shape; the constant initialisations (GraphNorm weight / bias / mean_scale, conv bias, ...) are equal
    bit for bit; every randomly initialised tensor of >= 1024 elements has the reference's spread (uniform bound within
    3 %, standard deviation within 8 % — PyG glorot for lin_l / lin_r / lin_edge / att, torch's Linear default for the
    MLPs).  The draws themselves differ (the modules call the RNG in a different order)."""
    import torch

    from ISubGVQA.models.mgat import MGAT as Ref
    from isg_b200.isubgvqa import MGAT

    kw = dict(channels=64, num_ins=4, heads=4, use_instr=True, masking_thresholds=[1.0, 1.0, 1.0, 0.1], use_topk=True,
              interpretable_mode=False, sampler_type="aimle", sample_k=2, nb_samples=1, alpha=1.0, beta=10.0, tau=1.0)
    with rl.scratch_cwd():
        torch.manual_seed(123)
        ref = Ref(**kw).state_dict()
    torch.manual_seed(123)
    ours = MGAT(**kw).state_dict()
    assert set(ref) <= set(ours) and all(k.endswith("aimle_state") for k in set(ours) - set(ref))
    constant = random = 0
    for k, r in ref.items():
        o = ours[k]
        assert o.shape == r.shape and o.dtype == r.dtype, k
        if r.numel() > 1 and float(r.float().std()) == 0.0:
            assert torch.equal(o, r), f"{k}: constant initialisation differs"
            constant += 1
        elif r.numel() >= 1024:
            rb, ob = float(r.abs().max()), float(o.abs().max())
            rs_, os_ = float(r.std()), float(o.std())
            assert abs(ob - rb) <= 0.03 * rb and abs(os_ - rs_) <= 0.08 * rs_, (k, rb, ob, rs_, os_)
            random += 1
    assert constant >= 12 and random >= 30, (constant, random)


def test_widened_modules_are_initialised_like_the_reference(reference):
    """The same for the widened rows: GlobalAttention (f1) and the scene-graph encoding MetaLayer (f2) — state_dict
    keys and shapes equal the reference's, random tensors have its spread."""
    import torch

    from ISubGVQA.models.att_pooling import GlobalAttention as RefGA
    from isg_b200.isubgvqa import GlobalAttention, SceneGraphEncodingLayer

    with rl.scratch_cwd():
        torch.manual_seed(7)
        ref_ga = RefGA(num_node_features=256, num_out_features=256).state_dict()
        ref_layer = rl.load_scene_graph_encoding_layer(256, 256, 256).state_dict()
    torch.manual_seed(7)
    our_ga = GlobalAttention(num_node_features=256, num_out_features=256).state_dict()
    our_layer = SceneGraphEncodingLayer(256, 256, 256).state_dict()
    for ref, ours, what in ((ref_ga, our_ga, "GlobalAttention"), (ref_layer, our_layer, "SceneGraphEncodingLayer")):
        assert list(ref) == list(ours), what
        for k, r in ref.items():
            o = ours[k]
            assert o.shape == r.shape and o.dtype == r.dtype, (what, k)
            if r.numel() >= 1024:
                rb, ob, rs_, os_ = float(r.abs().max()), float(o.abs().max()), float(r.std()), float(o.std())
                assert abs(ob - rb) <= 0.03 * rb and abs(os_ - rs_) <= 0.08 * rs_, (what, k, rb, ob, rs_, os_)
