"""Shared helpers for the parity tests: rebuild a golden case's inputs from its config, run the
oracle / the CUDA drop-in on it, compare."""
import glob
import math
import os

import torch

from isg_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-4  # BASELINE.json north_star: logits, answers and gradients within 1e-4 relative in fp32


def golden_files():
    """MGAT fixtures (the hot path proper)."""
    return sorted(p for p in glob.glob(os.path.join(GOLDEN_DIR, "*.pt")) if not os.path.basename(p).startswith("sgenc_"))


def sgenc_golden_files():
    """Scene-graph encoding layer fixtures."""
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "sgenc_*.pt")))


def sgenc_loss(xn, ee):
    w1 = torch.sin(torch.arange(xn.numel(), dtype=torch.float32, device=xn.device)).view_as(xn)
    w2 = torch.cos(torch.arange(ee.numel(), dtype=torch.float32, device=ee.device)).view_as(ee)
    return (xn * w1.to(xn.dtype)).sum() / xn.shape[0] + (ee * w2.to(ee.dtype)).sum() / ee.shape[0]


def run_oracle_sgenc(cfg, dtype=torch.float32):
    import isg_oracle as O

    C, B, seed = cfg["channels"], cfg["num_graphs"], cfg["seed"]
    sd, gnp = synth.make_sgenc_state_dict(C, C, C, seed)
    b = synth.make_batch(B, channels=C, mean_nodes=cfg["mean_nodes"], mean_edges=cfg["mean_edges"], seed=seed)
    p = {k: v.to(dtype).requires_grad_(True) for k, v in sd.items()}
    g = {k: v.to(dtype).requires_grad_(True) for k, v in gnp.items()}
    x = b["x"].to(dtype).requires_grad_(True)
    ea = b["edge_attr"].to(dtype).requires_grad_(True)
    xn, ee = O.scene_graph_encode(x, b["edge_index"], ea, b["batch"], p, g["weight"], g["bias"], g["mean_scale"], B)
    sgenc_loss(xn, ee).backward()
    pg = {k: v.grad for k, v in p.items()}
    pg.update({"graph_layer_norm." + k: v.grad for k, v in g.items()})
    return dict(x_encoded=xn.detach(), edge_attr_encoded=ee.detach(), gx=x.grad, g_edge_attr=ea.grad, param_grads=pg)


def run_cuda_sgenc(cfg, device="cuda"):
    from isg_b200.isubgvqa import GraphNorm64, SceneGraphEncodingLayer, encode_scene_graph

    C, B, seed = cfg["channels"], cfg["num_graphs"], cfg["seed"]
    sd, gnp = synth.make_sgenc_state_dict(C, C, C, seed)
    b = synth.make_batch(B, channels=C, mean_nodes=cfg["mean_nodes"], mean_edges=cfg["mean_edges"], seed=seed)
    layer = SceneGraphEncodingLayer(C, C, C)
    layer.load_state_dict(sd, strict=True)
    gn = GraphNorm64(C)
    gn.load_state_dict(gnp, strict=True)
    layer.to(device), gn.to(device)
    x = b["x"].to(device).requires_grad_(True)
    ea = b["edge_attr"].to(device).requires_grad_(True)
    xn, ee = encode_scene_graph(layer, gn, x, b["edge_index"].to(device), ea, b["batch"].to(device), B)
    sgenc_loss(xn, ee).backward()
    pg = {k: p.grad.detach().cpu() for k, p in layer.named_parameters()}
    pg.update({"graph_layer_norm." + k: p.grad.detach().cpu() for k, p in gn.named_parameters()})
    return dict(x_encoded=xn.detach().cpu(), edge_attr_encoded=ee.detach().cpu(), gx=x.grad.cpu(),
                g_edge_attr=ea.grad.cpu(), param_grads=pg)


def compare_sgenc(got, want, rtol=RTOL):
    for key in ("x_encoded", "edge_attr_encoded", "gx", "g_edge_attr"):
        assert rel_err(got[key], want[key]) <= rtol, (key, rel_err(got[key], want[key]))
    for name, w in want["param_grads"].items():
        check_param_grad(name, got["param_grads"].get(name), w, rtol)


def load_golden(path):
    return torch.load(path, weights_only=False)


def case_noise(sampler, B, nmax, seed):
    if sampler in ("imle", "aimle"):
        return synth.gumbel_noise(B, nmax, 0.3, seed=seed)
    if sampler == "gumbel":
        return synth.gumbel_noise(B, nmax, 1.0, seed=seed)[:, 0, :, 0].contiguous()
    npad = 2 ** math.ceil(math.log2(nmax))
    return synth.gumbel_noise(B, npad, 1.0, seed=seed)[:, 0, :, 0].contiguous()


def case_dropout(N, train, seed):
    if not train:
        return None
    g = torch.Generator().manual_seed(seed + 17)
    return (torch.rand(N, 1, generator=g) > 0.2).float() / 0.8


def loss_fn(h):
    w = torch.sin(torch.arange(h.numel(), dtype=torch.float32, device=h.device)).to(h.dtype).view_as(h)
    return (h * w).sum() / h.shape[0] + (h * h).mean()


def rel_err(a, b):
    """max |a-b| / max |b|  (relative to the tensor's scale, the metric the 1e-4 bound is read in)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = float(b.abs().max())
    if denom == 0.0:
        return float((a - b).abs().max())
    return float((a - b).abs().max()) / denom


def digest(t):
    t = t.detach().flatten().to(torch.float64).cpu()
    return dict(sum=float(t.sum()), abssum=float(t.abs().sum()), head=t[:64].to(torch.float32).clone(),
                numel=t.numel())


def check_param_grad(name, got, want, rtol=RTOL):
    if want is None:
        assert got is None or float(got.abs().max()) == 0.0, f"{name}: reference has no grad"
        return
    assert got is not None, f"{name}: missing grad"
    if isinstance(want, dict):
        d = digest(got)
        assert d["numel"] == want["numel"], name
        scale = max(want["abssum"], 1e-30)
        assert abs(d["abssum"] - want["abssum"]) <= rtol * scale, (name, d["abssum"], want["abssum"])
        assert abs(d["sum"] - want["sum"]) <= rtol * scale, (name, d["sum"], want["sum"])
        hscale = max(float(want["head"].abs().max()), want["abssum"] / want["numel"])
        assert float((d["head"] - want["head"]).abs().max()) <= rtol * hscale * 4, name
    else:
        assert rel_err(got, want) <= rtol, (name, rel_err(got, want))


def run_oracle_case(cfg, step_count=None, dtype=torch.float32, replay=None, record=False, teacher=None,
                    kink_margin=False):
    """Runs oracle/isg_oracle.py::OracleMGAT on a golden config; returns per-step result dicts.
    dtype=float64 + replay=<records of an fp32 run> gives the high-precision arbiter: the discrete
    sampler decisions (mask, perturbation gradient) are replayed, everything else is fp64."""
    import isg_oracle as O

    C, B, seed, sampler, train = cfg["channels"], cfg["num_graphs"], cfg["seed"], cfg["sampler"], cfg["train"]
    b = synth.make_batch(B, channels=C, mean_nodes=cfg["mean_nodes"], mean_edges=cfg["mean_edges"], seed=seed)
    cat = bool(cfg.get("concat_instr"))
    model = O.OracleMGAT(channels=C, sampler_type=sampler, sample_k=cfg["k"], concat_instr=cat).to(dtype)
    model.load_state_dict({k: v.to(dtype) for k, v in synth.make_state_dict(C, 4, 4, seed, concat_instr=cat).items()})
    model.train(train)
    if cfg.get("aimle_beta0") is not None:
        for st in model.aimle_state:
            st.beta = cfg["aimle_beta0"]
    N = b["x"].shape[0]
    outs = []
    for step in range(step_count or cfg["steps"]):
        noise = case_noise(sampler, B, b["nmax"], seed + step).to(dtype)
        drop = case_dropout(N, train, seed + step)
        drop = drop.to(dtype) if drop is not None else None
        x = b["x"].to(dtype).clone().requires_grad_(True)
        ea = b["edge_attr"].to(dtype).clone().requires_grad_(True)
        iv = b["instr_vectors"].to(dtype).clone().requires_grad_(True)
        gl = b["global_language_feats"].to(dtype).clone().requires_grad_(True)
        model.zero_grad()
        model.record = {} if record else None
        model.replay = replay[step] if replay is not None else None
        model.teacher = teacher[step] if teacher is not None else None
        model.kink_margin = [] if kink_margin else None
        h, mask, _, _ = model(x, b["edge_index"], iv, gl, ea, b["batch"], noise=noise, theta_dropout_mask=drop)
        loss = loss_fn(h)
        loss.backward()
        pg = {k: (p.grad.clone() if p.grad is not None else None) for k, p in model.named_ref_parameters()}
        outs.append(dict(h=h.detach(), mask=mask.detach(), loss=float(loss.detach()), gx=x.grad, g_edge_attr=ea.grad,
                         g_instr=iv.grad, g_glf=gl.grad, param_grads=pg, record=model.record,
                         kink_margin=model.kink_margin))
    return outs


def run_oracle_fp64_arbiter(cfg):
    """fp32 oracle run (records the discrete sampler decisions) followed by an fp64 replay."""
    o32 = run_oracle_case(cfg, record=True)
    o64 = run_oracle_case(cfg, dtype=torch.float64, replay=[o["record"] for o in o32])
    return o32, o64


def run_cuda_case(cfg, step_count=None, device="cuda", capture=False, executor=True):
    """Runs the CUDA drop-in (isg_b200.isubgvqa.MGAT) on a golden config.  capture=True also returns the
    edge kernel's forward inputs of every layer (`teacher` dict for run_oracle_case).  executor=True (the
    default configuration of the package) runs MGAT.forward through the layer executor, False through the
    per-operator autograd Functions."""
    from isg_b200.isubgvqa import MGAT
    from isg_b200.isubgvqa import mgat as mgat_mod

    prev_exec = mgat_mod._USE_EXECUTOR
    mgat_mod.set_executor(executor)
    try:
        return _run_cuda_case(cfg, step_count, device, capture, executor, MGAT)
    finally:
        mgat_mod.set_executor(prev_exec)


def _run_cuda_case(cfg, step_count, device, capture, executor, MGAT):

    C, B, seed, sampler, train = cfg["channels"], cfg["num_graphs"], cfg["seed"], cfg["sampler"], cfg["train"]
    b = synth.make_batch(B, channels=C, mean_nodes=cfg["mean_nodes"], mean_edges=cfg["mean_edges"], seed=seed)
    model = MGAT(channels=C, num_ins=4, heads=4, use_instr=True, masking_thresholds=[1.0, 1.0, 1.0, 0.1],
                 use_topk=True, interpretable_mode=False, sampler_type=sampler, sample_k=cfg["k"], nb_samples=1,
                 alpha=1.0, beta=10.0, tau=1.0, concat_instr=bool(cfg.get("concat_instr")))
    model.load_state_dict(synth.make_state_dict(C, 4, 4, seed, concat_instr=bool(cfg.get("concat_instr"))))
    model.to(device)
    model.train(train)
    if cfg.get("aimle_beta0") is not None:
        model.convs[3].mask.sampler_train.target.beta = float(cfg["aimle_beta0"])
    N = b["x"].shape[0]
    ei, batch = b["edge_index"].to(device), b["batch"].to(device)
    outs = []
    for step in range(step_count or cfg["steps"]):
        noise = case_noise(sampler, B, b["nmax"], seed + step).to(device)
        drop = case_dropout(N, train, seed + step)
        x = b["x"].to(device).requires_grad_(True)
        ea = b["edge_attr"].to(device).requires_grad_(True)
        iv = b["instr_vectors"].to(device).requires_grad_(True)
        gl = b["global_language_feats"].to(device).requires_grad_(True)
        model.zero_grad()
        model.capture_activations = {} if (capture and executor) else None
        for conv in model.convs:
            conv.debug_tensors = {} if (capture and not executor) else None
            if conv.mask.masking_threshold != 1.0:
                conv.mask.injected_noise = noise
                conv.mask.injected_dropout_mask = drop.to(device) if drop is not None else None
        h, mask, _, _ = model(x, ei, iv, gl, ea, batch, return_masks=True)
        loss = loss_fn(h)
        loss.backward()
        pg = {k: (p.grad.detach().cpu() if p.grad is not None else None) for k, p in model.named_parameters()}
        teacher = None
        if capture and executor:
            teacher = {k: v.detach().cpu() for k, v in model.capture_activations.items()}
        elif capture:
            teacher = {f"{n}.{i}": conv.debug_tensors[n].detach().cpu() for i, conv in enumerate(model.convs)
                       for n in ("x_l", "x_r", "e_proj")}
        outs.append(dict(h=h.detach().cpu(), mask=mask.detach().cpu(), loss=float(loss.detach()), gx=x.grad.cpu(),
                         g_edge_attr=ea.grad.cpu(), g_instr=iv.grad.cpu(), g_glf=gl.grad.cpu(), param_grads=pg,
                         teacher=teacher))
    return outs


def compare_step(got, want, sampler, rtol=RTOL, exact=None):
    """got vs want (the fp32 reference / oracle).  `exact`, when given, is the fp64 arbiter replay of the same
    step (run_oracle_case(dtype=float64, replay=...)): a float tensor that misses rtol against the fp32
    reference still passes if it is within rtol of the exact value — two fp32 evaluations of an
    ill-conditioned quantity can legitimately differ by more than either differs from the truth."""
    if sampler in ("imle", "aimle"):
        assert torch.equal(got["mask"], want["mask"]), "top-k node mask must be bit-exact given the same noise"
    else:
        assert rel_err(got["mask"], want["mask"]) <= rtol
    for key in ("h", "gx", "g_edge_attr", "g_instr", "g_glf"):
        e = rel_err(got[key], want[key])
        if e > rtol and exact is not None:
            e = min(e, rel_err(got[key], exact[key]))
        assert e <= rtol, (key, e)
    for name, w in want["param_grads"].items():
        try:
            check_param_grad(name, got["param_grads"].get(name), w, rtol)
        except AssertionError:
            if exact is None or exact["param_grads"].get(name) is None:
                raise
            check_param_grad(name, got["param_grads"].get(name), exact["param_grads"][name], rtol)


def heavy_first_order(ptr, N, E, side="dst"):
    """The task schedule isg_degree_order / isg_b200.collate produce: nodes with degree >= max(2, ceil(2E/N)) first in
    increasing node id, then the others — increasing for the dst ordering, decreasing for the src ordering."""
    deg = (ptr[1:N + 1] - ptr[:N]).long()
    thr = max(2, -(-2 * E // max(N, 1)))
    ids = torch.arange(N)
    light = ids[deg < thr]
    return torch.cat([ids[deg >= thr], light.flip(0) if side == "src" else light])
