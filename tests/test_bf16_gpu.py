"""bf16 configuration (BASELINE.json config 3, second half): the tcgen05 kind::f16 projections against float64
products of the SAME bf16-rounded operands.  With identical operands the only differences are fp32 accumulation
order and the rounding of the output to its storage type, so the bars are: fp32 outputs <= 2e-5 of the tensor's
scale, bf16 outputs <= 2^-8 (one bf16 ulp of the largest element: 3.9e-3)."""
import ctypes

import numpy as np
import pytest
import torch

import util
from isg_b200 import lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF16_ULP = 2.0 ** -8


def _pad8(n):
    return (n + 7) // 8 * 8


def _bf16_padded(t):
    """[rows, cols] fp32 -> (bf16 device tensor [rows, pad8(cols)] with zero pad columns, the rounded values as f64)."""
    rows, cols = t.shape
    out = torch.zeros(rows, _pad8(cols), dtype=torch.bfloat16, device=DEV)
    out[:, :cols] = t.to(DEV).to(torch.bfloat16)
    return out, out[:, :cols].double().cpu()


def _gelu_grad(z):
    return 0.5 * (1 + torch.erf(z / 2 ** 0.5)) + z * torch.exp(-0.5 * z * z) / (2 * torch.pi) ** 0.5


@pytest.mark.parametrize("M,K,Nout,act,out_bf16", [
    (1, 300, 1200, 0, True), (37, 300, 2400, 0, True), (4910, 300, 1200, 0, True), (515, 1200, 600, 1, True),
    (4910, 600, 300, 1, False), (129, 300, 300, 0, False), (3000, 64, 48, 1, True), (40000, 300, 1200, 0, True),
    # resident-weight-tile form (reduction <= 320, >= 2 row blocks per CTA): forward with a ragged last column block,
    # dgrad through Nout = 300, fp32 output with GELU
    (4910, 300, 2400, 0, True), (20000, 1200, 300, 1, True), (9000, 300, 2400, 1, False),
])
def test_linear_bf16_fwd_dgrad_wgrad(M, K, Nout, act, out_bf16):
    lib = L.load()
    g = torch.Generator().manual_seed(M + K + Nout)
    x, w = torch.randn(M, K, generator=g), torch.randn(Nout, K, generator=g) / K ** 0.5
    b, gy = torch.randn(Nout, generator=g) * 0.1, torch.randn(M, Nout, generator=g)
    xb, xr = _bf16_padded(x)
    gyb, gyr = _bf16_padded(gy)
    # weights through isg_weights_to_bf16 (W and W^T in one launch)
    w_d = w.to(DEV)
    ldw, ldt = _pad8(K), _pad8(Nout)
    wb = torch.full((Nout, ldw), float("nan"), dtype=torch.bfloat16, device=DEV)
    wt = torch.full((K, ldt), float("nan"), dtype=torch.bfloat16, device=DEV)
    arr = lambda v, ct: (ct * 1)(v)
    L.call("isg_weights_to_bf16", 1, arr(w_d.data_ptr(), ctypes.c_void_p), arr(Nout, ctypes.c_int), arr(K, ctypes.c_int),
           arr(wb.data_ptr(), ctypes.c_void_p), arr(ldw, ctypes.c_int), arr(wt.data_ptr(), ctypes.c_void_p),
           arr(ldt, ctypes.c_int), L.stream())
    wr = w.to(torch.bfloat16).double()
    assert torch.equal(wb[:, :K].double().cpu(), wr) and torch.equal(wt[:, :Nout].double().cpu(), wr.t())
    assert float(wb[:, K:].abs().sum()) == 0.0 and float(wt[:, Nout:].abs().sum()) == 0.0  # pad columns are zero
    # ---- forward
    odt = torch.bfloat16 if out_bf16 else torch.float32
    code = L.BF16 if out_bf16 else L.F32
    ldy = _pad8(Nout) if out_bf16 else Nout
    y = torch.zeros(M, ldy, dtype=odt, device=DEV)
    z = torch.zeros(M, ldy, dtype=odt, device=DEV) if act else None
    b_d = b.to(DEV)
    L.call("isg_linear_bf16_fwd", L.ptr(xb), xb.stride(0), L.ptr(wb), ldw, L.ptr(b_d), L.ptr(y), ldy, L.ptr(z), ldy, M,
           Nout, K, act, code, L.stream())
    z_ref = xr @ wr.t() + b.double()
    tol = BF16_ULP if out_bf16 else 2e-5
    if act:
        assert util.rel_err(z[:, :Nout].double().cpu(), z_ref) <= tol
        z_used = z[:, :Nout].double().cpu()  # the kernel applies GELU to the STORED pre-activation
        y_ref = torch.nn.functional.gelu(z_used)
    else:
        y_ref = z_ref
    assert util.rel_err(y[:, :Nout].double().cpu(), y_ref) <= tol, util.rel_err(y[:, :Nout].double().cpu(), y_ref)
    # ---- dgrad (x gelu'(z_prev) when a pre-activation of the PREVIOUS layer is given; it has K columns)
    ldgx = _pad8(K) if out_bf16 else K
    gx = torch.zeros(M, ldgx, dtype=odt, device=DEV)
    zp = (torch.randn(M, K, generator=g)).to(DEV).to(odt) if act else None
    zp_pad = None
    if zp is not None:
        zp_pad = torch.zeros(M, ldgx, dtype=odt, device=DEV)
        zp_pad[:, :K] = zp
    L.call("isg_linear_bf16_dgrad", L.ptr(gyb), gyb.stride(0), L.ptr(wt), ldt, L.ptr(zp_pad), ldgx, L.ptr(gx), ldgx, 0, M,
           Nout, K, code, L.stream())
    gx_ref = gyr @ wr
    if zp is not None:
        gx_ref = gx_ref * _gelu_grad(zp.double().cpu())
    assert util.rel_err(gx[:, :K].double().cpu(), gx_ref) <= tol, util.rel_err(gx[:, :K].double().cpu(), gx_ref)
    if not out_bf16:  # accumulate (fp32 outputs only)
        L.call("isg_linear_bf16_dgrad", L.ptr(gyb), gyb.stride(0), L.ptr(wt), ldt, L.ptr(zp_pad), ldgx, L.ptr(gx), ldgx, 1,
               M, Nout, K, code, L.stream())
        assert util.rel_err(gx[:, :K].double().cpu(), 2 * gx_ref) <= 2 * tol
    # ---- wgrad (fp32, deterministic split over M)
    gw = torch.empty(Nout, K, device=DEV)
    nb = lib.isg_linear_bf16_wgrad_workspace_bytes(M, Nout, K)
    ws = L.workspace(nb, DEV)
    for _ in range(2):
        L.call("isg_linear_bf16_wgrad", L.ptr(gyb), gyb.stride(0), L.ptr(xb), xb.stride(0), L.ptr(gw), M, Nout, K, L.ptr(ws),
               nb, L.stream())
        if _ == 0:
            first = gw.clone()
    assert torch.equal(first, gw)  # run-to-run identical
    gw_ref = gyr.t() @ xr
    assert util.rel_err(gw.double().cpu(), gw_ref) <= 2e-5, util.rel_err(gw.double().cpu(), gw_ref)


def test_to_bf16_pads_with_zeros():
    x = torch.randn(77, 300, device=DEV)
    out = torch.full((77, 304), float("nan"), dtype=torch.bfloat16, device=DEV)
    L.call("isg_to_bf16", L.ptr(x), 300, 77, 300, L.ptr(out), 304, L.stream())
    assert torch.equal(out[:, :300], x.to(torch.bfloat16))
    assert float(out[:, 300:].abs().sum()) == 0.0


# ---------------------------------------------------------------------------------------------------------------
# end to end: MGAT forward + backward in the bf16 configuration (gemm mode 3, layer executor) against the fp32
# configuration of the same CUDA path (itself within 1e-4 of the oracle: tests/test_mgat_gpu.py) on the same inputs
# and the same injected noise.  Stated tolerances (bf16 has 8 significand bits: 2^-8 = 3.9e-3 per rounding, compounded
# over four layers of projections):
#   * layer output h: ||a-b||_2 / ||b||_2 <= 2e-2 (measured 1e-2) and max|a-b| / max|b| <= 1e-1 (measured 2-6e-2);
#   * every gradient tensor: l2 error <= 0.2, i.e. cosine similarity >= 0.98 (measured 1e-2 ... 0.16), and all
#     parameter gradients taken as one vector: l2 error <= 0.1.
# Why the gradients are looser than the output: the attention backward forms a * (m t - sum a m t) with
# t = <g_out, x_l[src]> — a difference of nearly equal terms whenever the softmax is not saturated — so the 2^-9
# rounding of the bf16-stored g_out / x_l / e_proj is amplified (lin_r and its bias, which receive ONLY that term,
# are the worst: 0.07-0.11); and AIMLE's perturbation gradient is a difference of two top-k MAP states, which flips
# discretely under any perturbation of dy (gate-projection gradients 0.15).  The fp32 configuration is the parity
# configuration; this one trades that accuracy for HBM traffic and is reported separately (north star).  The gate logits that feed the discrete sampler stay fp32, but they
# are computed from bf16-perturbed activations, so a near-tie may flip: masks must agree on >= 95 % of the nodes
# and the numeric bars apply when they agree everywhere.
# ---------------------------------------------------------------------------------------------------------------
L2_TOL = 0.2  # ||a-b||_2 / ||b||_2 of every gradient tensor (cosine similarity >= 0.98)


def _l2_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm()) / max(float(b.norm()), 1e-300)


def _run_mode(cfg, mode, steps=1):
    from isg_b200 import ops

    prev = ops.gemm_mode()
    ops.set_gemm_mode(mode)
    try:
        return util.run_cuda_case(cfg, step_count=steps)
    finally:
        ops.set_gemm_mode(prev)


@pytest.mark.parametrize("sampler,train,B", [("imle", True, 12), ("aimle", True, 64), ("gumbel", False, 16)])
def test_mgat_bf16_configuration_tracks_fp32(sampler, train, B):
    cfg = dict(sampler=sampler, train=train, channels=300, num_graphs=B, mean_nodes=14, mean_edges=90, k=2,
               seed=77 + B, steps=1)
    if sampler == "aimle":
        cfg["aimle_beta0"] = 1.0
    want = _run_mode(cfg, 1)[0]
    got = _run_mode(cfg, 3)[0]
    for key in ("h", "gx", "g_edge_attr", "g_instr", "g_glf"):
        assert torch.isfinite(got[key]).all(), key
    agree = float((got["mask"] == want["mask"]).float().mean()) if sampler != "gumbel" else 1.0
    assert agree >= 0.95, agree
    if sampler == "gumbel":
        assert util.rel_err(got["mask"], want["mask"]) <= 5e-2
    if agree == 1.0:
        report = {}
        keys = ["h"] + (["gx", "g_edge_attr", "g_instr", "g_glf"] if train else [])
        for key in keys:
            report[key] = (util.rel_err(got[key], want[key]), _l2_err(got[key], want[key]))
        if train:
            for name, w in want["param_grads"].items():
                g = got["param_grads"].get(name)
                assert (g is None) == (w is None), name
                if w is not None:
                    report["param:" + name] = (util.rel_err(g, w), _l2_err(g, w))
        print("bf16 vs fp32 (max-norm, l2):", {k: (round(a, 4), round(b, 4)) for k, (a, b) in report.items()})
        assert report["h"][1] <= 2e-2 and report["h"][0] <= 1e-1, report["h"]
        for key, (emax, el2) in report.items():
            assert el2 <= L2_TOL, (key, emax, el2)
        if train:
            names = [n for n, w in want["param_grads"].items() if w is not None]
            a = torch.cat([got["param_grads"][n].flatten().double() for n in names])
            b = torch.cat([want["param_grads"][n].flatten().double() for n in names])
            assert float((a - b).norm() / b.norm()) <= 0.1, float((a - b).norm() / b.norm())


def test_bf16_mode_refuses_the_per_operator_path():
    from isg_b200 import ops
    from isg_b200.isubgvqa import mgat as mgat_mod

    cfg = dict(sampler="imle", train=True, channels=300, num_graphs=4, mean_nodes=8, mean_edges=30, k=2, seed=5, steps=1)
    prev = ops.gemm_mode()
    ops.set_gemm_mode(3)
    try:
        with pytest.raises(NotImplementedError):
            util.run_cuda_case(cfg, executor=False)
    finally:
        ops.set_gemm_mode(prev)
        mgat_mod.set_executor(True)
