"""GPU parity at the REAL BASELINE.json sizes (not scaled-down stand-ins), through the nn.Module surface:

  c3  training step, AIMLE, 256 graphs      (config 3/4)  — forward + every gradient
  c2  inference, Gumbel, 1024 graphs        (config 2)    — forward
  AIMLE from the reference's own beta0 = 0  (models/masking.py:258) for three consecutive steps

The oracle port runs ~250 graphs/s on a few host cores, so each case costs seconds.

Gradients are checked twice.  (1) Teacher-forced: the oracle is re-run with the edge kernel's forward inputs
(x_l, x_r, e_proj of every layer) replaced by the CUDA values, so both sides evaluate leaky_relu on identical
pre-activations: every tensor must then agree to 1e-4 (or agree with the fp64 replay of the same step).
(2) Un-forced, every input and parameter gradient, against the EXACT gradient: the free-running oracle in fp64
(discrete sampler decisions replayed).  leaky_relu's derivative jumps at 0; of the ~1.9e8 GATv2 pre-activations
s = x_r[dst] + x_l[src] + e_proj of a 256-graph step, those within the forward rounding error of 0 get the other
slope than exact arithmetic gives them, and each such flip moves whole rows of the weight gradients by 1e-3..1e-2
of their scale.  The number of flips is proportional to the forward error of s: ~2e-7 for the CPU oracle's fp32
GEMMs (its gradients stay within 1e-5 of fp64 at this size), 1.4-2e-6 for the tensor-core 3xTF32 projections here
 — hence ~10x more flips and the measured un-forced errors below, largest on the masked layer whose
gradient is concentrated on k nodes per graph.  Measured (B200, seed 3407): relative L2 error vs fp64 <= 2.5e-4 on
57 of the 70 tensors, 6e-4..1.9e-3 on convs.3.{lin_l,lin_r,lin_edge} and 8e-4 on bns.2.{bias,mean_scale}.
The test holds every tensor to UNFORCED_RTOL_L2 and reports the table; the 1e-4 bar of BASELINE.json is met in the
teacher-forced comparison, i.e. for identical leaky_relu branch decisions."""
import pytest
import torch

import util
from isg_b200 import synth

pytestmark = pytest.mark.gpu

UNFORCED_RTOL_L2 = 3e-3   # every gradient tensor, relative L2 error against the free-running fp64 oracle


def _rel_l2(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a - b).norm()) / (float(b.norm()) or 1.0)


def _check_unforced(got, free32, exact, names=("gx", "g_edge_attr", "g_instr", "g_glf")):
    """got: CUDA; free32: free-running fp32 oracle; exact: free-running fp64 oracle (sampler decisions replayed)."""
    rows = {}
    for key in names:
        rows[key] = (_rel_l2(got[key], exact[key]), _rel_l2(free32[key], exact[key]))
    for name, w in exact["param_grads"].items():
        g = got["param_grads"].get(name)
        if w is None:
            assert g is None or float(g.abs().max()) == 0.0, name
            continue
        assert g is not None, name
        rows[name] = (_rel_l2(g, w), _rel_l2(free32["param_grads"][name], w))
    bad = {k: v for k, v in rows.items() if v[0] > UNFORCED_RTOL_L2}
    assert not bad, f"un-forced gradients, relative L2 error vs fp64 (cuda, fp32 oracle): {bad}"
    return rows


def test_c3_training_step_matches_oracle_at_256_graphs():
    cfg = dict(sampler="aimle", train=True, channels=300, num_graphs=256, mean_nodes=20, mean_edges=150, k=2,
               seed=3407, steps=1, aimle_beta0=1.0)
    got = util.run_cuda_case(cfg, capture=True)[0]
    free = util.run_oracle_case(cfg, record=True)[0]
    assert util.rel_err(got["h"], free["h"]) <= util.RTOL
    assert torch.equal(got["mask"], free["mask"])
    assert abs(got["loss"] - free["loss"]) <= util.RTOL * abs(free["loss"])
    want = util.run_oracle_case(cfg, teacher=[got["teacher"]])[0]
    exact = util.run_oracle_case(cfg, dtype=torch.float64, replay=[free["record"]], teacher=[got["teacher"]])[0]
    util.compare_step(got, want, "aimle", exact=exact)
    truth = util.run_oracle_case(cfg, dtype=torch.float64, replay=[free["record"]])[0]
    rows = _check_unforced(got, free, truth)
    worst = max(rows.items(), key=lambda kv: kv[1][0])
    print(f"c3 un-forced vs fp64: worst tensor {worst[0]}: cuda {worst[1][0]:.2e}, fp32 oracle {worst[1][1]:.2e}; "
          f"median cuda {sorted(v[0] for v in rows.values())[len(rows) // 2]:.2e}, "
          f"median fp32 oracle {sorted(v[1] for v in rows.values())[len(rows) // 2]:.2e}; "
          f"tensors above 2.5e-4: {sorted(k for k, v in rows.items() if v[0] > 2.5e-4)}")


def test_c2_inference_matches_oracle_at_1024_graphs():
    import isg_oracle as O
    from isg_b200.isubgvqa import MGAT

    B, seed = 1024, 3407
    b = synth.make_batch(B, seed=seed)
    gum = util.case_noise("gumbel", B, b["nmax"], seed)
    sd = synth.make_state_dict(seed=seed)
    oracle = O.OracleMGAT(channels=300, sampler_type="gumbel", sample_k=2)
    oracle.load_state_dict(sd)
    oracle.eval()
    with torch.no_grad():
        h_o, m_o, _, _ = oracle(b["x"], b["edge_index"], b["instr_vectors"], b["global_language_feats"],
                                b["edge_attr"], b["batch"], noise=gum)
    model = MGAT(channels=300, num_ins=4, heads=4, use_instr=True, masking_thresholds=[1.0, 1.0, 1.0, 0.1],
                 use_topk=True, interpretable_mode=False, sampler_type="gumbel", sample_k=2)
    model.load_state_dict(sd)
    model.cuda().eval()
    model.convs[3].mask.injected_noise = gum.cuda()
    with torch.no_grad():
        h, m, _, _ = model(*[b[k].cuda() for k in ("x", "edge_index", "instr_vectors", "global_language_feats",
                                                    "edge_attr", "batch")], return_masks=True)
    assert util.rel_err(h, h_o) <= util.RTOL
    assert util.rel_err(m, m_o) <= util.RTOL
    assert torch.equal(m.cpu() > 0.5, m_o > 0.5)  # the same nodes are selected


def test_aimle_from_reference_initial_beta_zero():
    """The reference starts AIMLE at beta = 0 (models/masking.py:258): pm = beta*|theta|/|dy| = 0, both target
    MAPs coincide and the sampler's gradient is EXACTLY zero on the first step (target_aimle.py:111-115,
    aimle.py:186-226), then beta grows by 1e-4 per step.  Three consecutive steps, state carried across them."""
    cfg = dict(sampler="aimle", train=True, channels=300, num_graphs=24, mean_nodes=20, mean_edges=150, k=2,
               seed=77, steps=3, aimle_beta0=0.0)
    got = util.run_cuda_case(cfg, capture=True)
    free = util.run_oracle_case(cfg, record=True)
    gate = ("convs.3.mask.node_nn.0.weight", "convs.3.mask.node_nn.0.bias", "convs.3.mask.ques_nn.0.weight",
            "convs.3.mask.ques_nn.0.bias")
    for name in gate:  # step 0: nothing flows through the sampler
        g = got[0]["param_grads"][name]
        assert g is not None and float(g.abs().max()) == 0.0, name
        w = free[0]["param_grads"][name]
        assert w is None or float(w.abs().max()) == 0.0, name
    want = util.run_oracle_case(cfg, teacher=[g["teacher"] for g in got])
    exact = util.run_oracle_case(cfg, dtype=torch.float64, replay=[f["record"] for f in free],
                                 teacher=[g["teacher"] for g in got])
    for s in range(3):
        assert torch.equal(got[s]["mask"], free[s]["mask"]), s
        assert util.rel_err(got[s]["h"], free[s]["h"]) <= util.RTOL
        util.compare_step(got[s], want[s], "aimle", exact=exact[s])


def test_aimle_state_is_checkpointed_and_resumes_bit_identically():
    """AIMLE beta / gradient-norm EMA live in MaskingModel's persistent buffer `aimle_state`: a run that is
    checkpointed after step 1 and resumed in a fresh model continues exactly like the uninterrupted run."""
    from isg_b200.isubgvqa import MGAT

    cfg = dict(sampler="aimle", train=True, channels=64, num_graphs=8, mean_nodes=10, mean_edges=50, k=2,
               seed=11, steps=3, aimle_beta0=3.0)
    full = util.run_cuda_case(cfg)

    def make():
        m = MGAT(channels=64, num_ins=4, heads=4, use_instr=True, masking_thresholds=[1.0, 1.0, 1.0, 0.1],
                 use_topk=True, interpretable_mode=False, sampler_type="aimle", sample_k=2)
        m.load_state_dict(synth.make_state_dict(64, 4, 4, 11))  # reference-layout dict: no aimle_state key
        return m.cuda().train()

    b = synth.make_batch(8, channels=64, mean_nodes=10, mean_edges=50, seed=11)
    args = [b[k].cuda() for k in ("x", "edge_index", "instr_vectors", "global_language_feats", "edge_attr", "batch")]

    def step(m, s):
        m.zero_grad()
        m.convs[3].mask.injected_noise = util.case_noise("aimle", 8, b["nmax"], 11 + s).cuda()
        m.convs[3].mask.injected_dropout_mask = util.case_dropout(b["x"].shape[0], True, 11 + s).cuda()
        h, mask, _, _ = m(*args, return_masks=True)
        util.loss_fn(h).backward()
        return h.detach().cpu(), {k: p.grad.cpu() for k, p in m.named_parameters() if p.grad is not None}

    m1 = make()
    m1.convs[3].mask.sampler_train.target.beta = 3.0
    step(m1, 0)
    step(m1, 1)
    sd = {k: v.cpu().clone() for k, v in m1.state_dict().items()}
    assert "convs.3.mask.aimle_state" in sd and sd["convs.3.mask.aimle_state"].dtype == torch.float64
    assert abs(float(sd["convs.3.mask.aimle_state"][0]) - 3.0) > 1e-5  # beta moved
    m2 = make()
    m2.load_state_dict(sd, strict=True)
    h2, g2 = step(m2, 2)
    assert torch.equal(h2, full[2]["h"])
    for k, g in g2.items():
        assert torch.equal(g, full[2]["param_grads"][k]), k
