"""Live pinning of the oracle: runs the UNMODIFIED reference (imported from /root/reference through
oracle/reference_loader.py on the pure-torch shim) next to oracle/isg_oracle.py on freshly seeded inputs that
are NOT among the committed fixtures.  Skipped wherever the reference tree is absent (the GPU box)."""
import math
import os
import sys

import pytest
import torch

import util

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import reference_loader as rl  # noqa: E402

pytestmark = pytest.mark.skipif(not rl.available(), reason="/root/reference is not present on this machine")


@pytest.fixture(scope="module")
def golden_mod():
    import make_golden

    with rl.scratch_cwd():
        rl.load()
    return make_golden


@pytest.mark.parametrize("sampler,train,C,B,k,seed", [("imle", True, 16, 5, 2, 4100), ("aimle", True, 16, 5, 3, 4101),
                                                      ("gumbel", True, 16, 6, 2, 4102), ("simple", True, 16, 6, 2, 4103),
                                                      ("simple", False, 32, 4, 3, 4104), ("imle", False, 16, 3, 2, 4105)])
def test_oracle_matches_live_reference(golden_mod, sampler, train, C, B, k, seed):
    steps = 2 if sampler == "aimle" else 1
    torch.manual_seed(seed)
    ref = golden_mod.run_reference(sampler, train, C, B, 9, 40, k, seed, steps)
    cfg = dict(sampler=sampler, train=train, channels=C, num_graphs=B, mean_nodes=9, mean_edges=40, k=k, seed=seed,
               steps=steps, aimle_beta0=golden_mod.AIMLE_BETA0 if sampler == "aimle" else None)
    got = util.run_oracle_case(cfg)
    for g, w in zip(got, ref):
        # two CPU fp32 evaluations of the Gumbel tau=0.1 mask-network gradient already differ by ~6e-5
        util.compare_step(g, w, sampler, rtol=util.RTOL if (sampler == "gumbel" and train) else 2e-5)


@pytest.mark.parametrize("sampler,train,B,k,mn,me,seed", [
    ("imle", True, 1, 2, 9, 40, 5000),    # a single graph
    ("imle", True, 4, 3, 2, 3, 5001),     # k = Nmax: every node of the largest graph is selected
    ("aimle", True, 4, 3, 2, 3, 5002),    # k > Nmax (local_k = min(k, Nmax), imle_scheme.py)
    ("gumbel", True, 4, 3, 2, 3, 5003),   # k > Nmax through the relaxed top-k
    ("imle", True, 3, 5, 3, 6, 5007),     # k far above every graph size
    ("simple", False, 1, 2, 2, 2, 5005),  # one 3-node graph, padded to 4 leaves
    ("gumbel", False, 1, 2, 9, 40, 5006)])
def test_oracle_matches_live_reference_on_degenerate_batches(golden_mod, sampler, train, B, k, mn, me, seed):
    """Edge cases the reference handles implicitly: one graph, graphs smaller than k, k above Nmax.  Tolerance is the
    1e-4 bar itself: on 6-10 node batches two fp32 evaluations of the same sums differ by up to ~3e-5."""
    steps = 2 if sampler == "aimle" else 1
    torch.manual_seed(seed)
    ref = golden_mod.run_reference(sampler, train, 16, B, mn, me, k, seed, steps)
    cfg = dict(sampler=sampler, train=train, channels=16, num_graphs=B, mean_nodes=mn, mean_edges=me, k=k, seed=seed,
               steps=steps, aimle_beta0=golden_mod.AIMLE_BETA0 if sampler == "aimle" else None)
    got = util.run_oracle_case(cfg)
    assert len(got) == len(ref)
    for g, w in zip(got, ref):
        assert bool(torch.isfinite(w["h"]).all()) and bool(torch.isfinite(w["mask"]).all())
        util.compare_step(g, w, sampler, rtol=util.RTOL)


@pytest.mark.parametrize("sampler", ["imle", "aimle", "gumbel", "simple"])
def test_oracle_matches_live_reference_on_unselected_seeds(golden_mod, sampler):
    """Ten CONSECUTIVE seeds per sampler (6000-6009), training and evaluation mode — no scan, no selection (the
    committed fixtures' seeds were picked away from leaky-ReLU kinks, oracle/make_golden.py:30-34; this sweep is not).
    Bound: the 1e-4 bar, except the Gumbel tau = 0.1 relaxation in training mode, whose gradient is ill-conditioned
    in fp32 (log(1 - onehot) with onehot -> 1, gumbel_scheme.py:76-80): at seed 6003 the two fp32 evaluations differ
    by 2.8e-4 on every gradient while the oracle in fp64 is within 1.1e-5 of the reference's fp32 — same function,
    different rounding of theta by one ulp; at seed 6009 it is the reference's fp32 that sits 1.6e-4 from the fp64
    value — so in that one mode both fp32 evaluations are held to 1e-3 of each other and of the oracle's fp64 replay
    (the arbiter), and `h` to the 1e-4 bar."""
    for seed in range(6000, 6010):
        for train in ((True,) if sampler == "aimle" else (True, False)):
            steps = 2 if sampler == "aimle" else 1
            torch.manual_seed(seed)
            ref = golden_mod.run_reference(sampler, train, 16, 5, 9, 40, 2, seed, steps)
            cfg = dict(sampler=sampler, train=train, channels=16, num_graphs=5, mean_nodes=9, mean_edges=40, k=2,
                       seed=seed, steps=steps, aimle_beta0=golden_mod.AIMLE_BETA0 if sampler == "aimle" else None)
            if sampler == "gumbel" and train:
                got, exact = util.run_oracle_fp64_arbiter(cfg)
                util.compare_step(got[0], ref[0], sampler, rtol=1e-3)
                assert util.rel_err(got[0]["h"], ref[0]["h"]) <= util.RTOL
                for key in ("h", "gx", "g_edge_attr", "g_instr", "g_glf"):
                    assert util.rel_err(ref[0][key], exact[0][key]) <= 1e-3, (seed, key)
                    assert util.rel_err(got[0][key], exact[0][key]) <= 1e-3, (seed, key)
                continue
            got = util.run_oracle_case(cfg)
            for g, w in zip(got, ref):
                util.compare_step(g, w, sampler, rtol=util.RTOL)


def test_simple_with_graphs_smaller_than_k_is_nan_in_the_reference_and_in_the_oracle(golden_mod):
    """SIMPLE in training mode on a batch whose smallest graphs have fewer than k nodes: the reference's exact-k circuit
    (simple.py:214-244) has no assignment with k ones among the real leaves, its marginals come out NaN for those graphs
    — the restatement reproduces the NaNs in the same positions and the finite values elsewhere (the drop-in therefore
    owes nothing for this input; GQA scene graphs have >= 2 = k objects)."""
    seed, B, k = 5004, 4, 3
    torch.manual_seed(seed)
    ref = golden_mod.run_reference("simple", True, 16, B, 2, 3, k, seed, 1)[0]
    cfg = dict(sampler="simple", train=True, channels=16, num_graphs=B, mean_nodes=2, mean_edges=3, k=k, seed=seed, steps=1)
    got = util.run_oracle_case(cfg)[0]
    nan_ref, nan_got = torch.isnan(ref["mask"]), torch.isnan(got["mask"])
    assert bool(nan_ref.any()), "expected the reference to produce NaN marginals here"
    assert torch.equal(nan_ref, nan_got)
    assert torch.allclose(ref["mask"][~nan_ref], got["mask"][~nan_got], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("sampler,train,seed", [("imle", True, 4200), ("gumbel", False, 4201)])
def test_oracle_matches_live_reference_concat_instr(golden_mod, sampler, train, seed):
    """The `--concat_instr 1` variant (models/mgat_v2_conv.py:153-154, mgat.py:41-44): the conv sees
    [x, instruction[batch]] and its lin_l / lin_r / mask.node_nn take 2C inputs."""
    C, B, k = 16, 5, 2
    torch.manual_seed(seed)
    ref = golden_mod.run_reference(sampler, train, C, B, 9, 40, k, seed, 1, concat_instr=True)
    cfg = dict(sampler=sampler, train=train, channels=C, num_graphs=B, mean_nodes=9, mean_edges=40, k=k, seed=seed,
               steps=1, concat_instr=True)
    got = util.run_oracle_case(cfg)
    util.compare_step(got[0], ref[0], sampler, rtol=2e-5)


def test_simple_circuit_restatement_matches_live_layer(golden_mod):
    """Layer.log_pr (simple.py:214-244) vs oracle.simple_marginals on logits with exact zeros (the -1000 dummy
    pad regime), incl. the gradient that the reference obtains by differentiating through both passes."""
    import isg_oracle as O
    from ISubGVQA.sampling.methods.simple_scheme import EdgeSIMPLEBatched

    g = torch.Generator().manual_seed(5)
    for nmax, k in [(9, 2), (20, 2), (33, 4), (64, 5)]:
        with rl.scratch_cwd():
            s = EdgeSIMPLEBatched(k=k, device="cpu", policy="edge_candid")
            theta = torch.randn(6, nmax, 1, generator=g)
            theta[torch.rand(6, nmax, 1, generator=g) < 0.25] = 0.0
            theta[3, nmax // 2:, 0] = 0.0
            t1 = theta.clone().requires_grad_(True)
            _, marg = s(t1, train=True)
        w = torch.randn(6, nmax, 1, generator=g)
        (marg * w).sum().backward()
        t2 = theta.clone().requires_grad_(True)
        npad = 2 ** math.ceil(math.log2(nmax))
        flat = torch.cat([t2[..., 0], t2.new_full((6, npad - nmax), -1.0e10)], dim=1)
        m2 = O.simple_marginals(flat, k)[:, :nmax]
        (m2 * w[..., 0]).sum().backward()
        assert torch.allclose(m2, marg[..., 0], rtol=0, atol=2e-7)
        assert util.rel_err(t2.grad, t1.grad) <= 1e-5


def test_global_attention_pool_matches_live_reference(golden_mod):
    """oracle.global_attention_pool vs the reference GlobalAttention
    (models/att_pooling.py), whose hard-coded `.cuda()` calls are neutralised on this CPU-only machine."""
    import isg_oracle as O
    from ISubGVQA.models.att_pooling import GlobalAttention

    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        g = torch.Generator().manual_seed(77)
        counts = torch.tensor([5, 1, 9, 3, 7])
        batch = torch.repeat_interleave(torch.arange(5), counts)
        N, D = int(counts.sum()), 48
        ref = GlobalAttention(num_node_features=D, num_out_features=D)
        x = torch.randn(N, D, generator=g, requires_grad=True)
        u = torch.randn(5, D, generator=g, requires_grad=True)
        mask = (torch.rand(N, 1, generator=g) > 0.4).float().requires_grad_(True)
        wo, wg = torch.randn(5, D, generator=g), torch.randn(N, 1, generator=g)
        out, gate = ref(x, u, batch, size=None, return_mask=True, node_mask=mask)
        ((out * wo).sum() + (gate * wg).sum()).backward()
        want = (out.detach(), gate.detach(), x.grad.clone(), u.grad.clone(), mask.grad.clone())
        params = {k: v.detach() for k, v in ref.state_dict().items()}
        x2, u2, m2 = (t.detach().clone().requires_grad_(True) for t in (x, u, mask))
        o2, g2 = O.global_attention_pool(x2, u2, batch, 5, params, node_mask=m2)
        ((o2 * wo).sum() + (g2 * wg).sum()).backward()
    finally:
        torch.Tensor.cuda = orig_cuda
    for a, b in zip((o2, g2, x2.grad, u2.grad, m2.grad), want):
        assert util.rel_err(a, b) <= 1e-6


def test_scene_graph_encoding_layer_restatement_matches_live_reference():
    """oracle.scene_graph_encode vs the reference's own MetaLayer (built by
    get_gt_scene_graph_encoding_layer, models/scene_graph_encoder.py:107-146) followed by the float64 GraphNorm
    of SceneGraphEncoder.forward (:99-102), values and every gradient."""
    import isg_oracle as O
    from isg_b200 import synth

    torch.manual_seed(11)
    with rl.scratch_cwd():
        layer = rl.load_scene_graph_encoding_layer(24, 24, 24)
    import torch_geometric  # the shim (on sys.path once the reference is loaded)

    gn = torch_geometric.nn.norm.GraphNorm(24)
    with torch.no_grad():
        gn.weight.uniform_(0.5, 1.5), gn.bias.normal_(0, 0.1), gn.mean_scale.uniform_(0.5, 1.5)
    b = synth.make_batch(5, channels=24, mean_nodes=8, mean_edges=30, seed=9)
    x = b["x"].clone().requires_grad_(True)
    ea = b["edge_attr"].clone().requires_grad_(True)
    xe, ee, _ = layer(x=x, edge_index=b["edge_index"], edge_attr=ea, u=None, batch=b["batch"])
    save = xe.dtype
    xn = gn(xe.type(torch.DoubleTensor), b["batch"]).type(save)  # scene_graph_encoder.py:99-102
    w1, w2 = torch.randn_like(xn), torch.randn_like(ee)
    ((xn * w1).sum() + (ee * w2).sum()).backward()
    ref = dict(xn=xn.detach(), ee=ee.detach(), gx=x.grad.clone(), gea=ea.grad.clone(),
               params={k: p.grad.clone() for k, p in layer.named_parameters()},
               gn={k: p.grad.clone() for k, p in gn.named_parameters()})

    p = {k: v.detach().clone().requires_grad_(True) for k, v in layer.state_dict().items()}
    g = {k: v.detach().clone().requires_grad_(True) for k, v in gn.state_dict().items()}
    x2 = b["x"].clone().requires_grad_(True)
    ea2 = b["edge_attr"].clone().requires_grad_(True)
    yn, ye = O.scene_graph_encode(x2, b["edge_index"], ea2, b["batch"], p, g["weight"], g["bias"], g["mean_scale"], 5)
    ((yn * w1).sum() + (ye * w2).sum()).backward()
    assert util.rel_err(yn, ref["xn"]) <= 1e-6 and util.rel_err(ye, ref["ee"]) <= 1e-6
    assert util.rel_err(x2.grad, ref["gx"]) <= 1e-5 and util.rel_err(ea2.grad, ref["gea"]) <= 1e-5
    for k, w in ref["params"].items():
        assert util.rel_err(p[k].grad, w) <= 1e-5, k
    for k, w in ref["gn"].items():
        assert util.rel_err(g[k].grad, w) <= 1e-5, k
