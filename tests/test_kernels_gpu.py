"""GPU parity tests, kernel by kernel: every libisg.so entry point (called through the C ABI via the
ctypes/autograd wrappers) against oracle/isg_oracle.py on the same seeded inputs.
Bar: bit-exact for integer/index work and top-k masks; <= 1e-4 relative for fp32 values/gradients."""
import math

import numpy as np
import pytest
import torch

import util
from isg_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda"
RTOL = util.RTOL


def _gi(edge_index, batch, B):
    from isg_b200.graph import GraphIndex

    return GraphIndex(edge_index.to(DEV), batch.to(DEV), B)


# ------------------------------------------------------------------------------------------ (a) CSR
def _check_csr(ei, N):
    import isg_oracle as O

    B = 1
    gi = _gi(ei, torch.zeros(N, dtype=torch.int64), B)
    want = O.csr_build(ei, N)
    assert int(gi.status.item()) == 0
    for key in ("dst_ptr", "dst_eid", "dst_nbr", "src_ptr", "src_eid", "src_nbr"):
        assert torch.equal(getattr(gi, key).cpu(), want[key]), key


@pytest.mark.parametrize("B,mn,me", [(1, 5, 12), (7, 9, 40), (64, 20, 150), (3, 150, 3000)])
def test_csr_matches_stable_sort(B, mn, me):
    t = synth.make_topology(B, mn, me, seed=B * 7 + 1, max_nodes=None)
    _check_csr(t["edge_index"], t["batch"].numel())


def test_csr_shuffled_and_degenerate_inputs():
    g = torch.Generator().manual_seed(0)
    t = synth.make_topology(5, 12, 80, seed=3)
    ei = t["edge_index"]
    perm = torch.randperm(ei.size(1), generator=g)
    _check_csr(ei[:, perm].contiguous(), t["batch"].numel())  # arbitrary COO order
    _check_csr(torch.zeros(2, 0, dtype=torch.int64), 6)  # no edges
    hub = torch.stack([torch.arange(300) % 7, torch.zeros(300, dtype=torch.int64)])  # one node, 300 in-edges
    _check_csr(hub, 7)
    dup = torch.tensor([[0, 0, 0, 1, 1], [1, 1, 1, 0, 0]])  # duplicate edges keep their order
    _check_csr(dup, 2)


def test_csr_large_properties():
    """B = 4096 x (20 nodes, 150 edges): too big for a python oracle loop -> check invariants."""
    t = synth.make_topology(4096, 20, 150, seed=9)
    ei = t["edge_index"]
    N = t["batch"].numel()
    gi = _gi(ei, t["batch"], 4096)
    eid = gi.dst_eid.cpu().long()
    ptr = gi.dst_ptr.cpu().long()
    assert ptr[0] == 0 and ptr[-1] == ei.size(1)
    assert torch.equal(torch.sort(eid).values, torch.arange(ei.size(1)))  # a permutation
    dst_sorted = ei[1][eid]
    assert bool((dst_sorted[1:] >= dst_sorted[:-1]).all())  # sorted by destination
    same = dst_sorted[1:] == dst_sorted[:-1]
    assert bool((eid[1:][same] > eid[:-1][same]).all())  # stable inside a segment
    assert torch.equal(torch.bincount(ei[1], minlength=N), ptr[1:] - ptr[:-1])
    assert torch.equal(gi.dst_nbr.cpu().long(), ei[0][eid])
    seid = gi.src_eid.cpu().long()
    assert torch.equal(gi.src_nbr.cpu().long(), ei[1][seid])
    assert gi.nmax == int(t["num_nodes"].max())


def test_graph_ptr_with_empty_graphs():
    batch = torch.tensor([0, 0, 2, 2, 2, 5], dtype=torch.int64)
    gi = _gi(torch.zeros(2, 0, dtype=torch.int64), batch, 7)
    assert gi.graph_ptr.cpu().tolist() == [0, 2, 2, 5, 5, 5, 6, 6]
    assert gi.nmax == 3
    assert gi.batch32.cpu().tolist()[:6] == batch.tolist()


def test_csr_reports_out_of_range_indices():
    ei = torch.tensor([[0, 1, 9], [1, 0, 0]])
    gi = _gi(ei, torch.zeros(3, dtype=torch.int64), 1)
    assert int(gi.status.item()) == 1
    with pytest.raises(IndexError):
        gi.check_indices()


# ------------------------------------------------------------------------------------------ (b) edge kernel
def _edge_case(B, mn, me, C, H, masked, seed, general_mask=False):
    t = synth.make_topology(B, mn, me, seed=seed, max_nodes=None)
    g = torch.Generator().manual_seed(seed)
    N, E = t["batch"].numel(), t["edge_index"].size(1)
    d = dict(t)
    d["x_l"] = torch.randn(N, H * C, generator=g)
    d["x_r"] = torch.randn(N, H * C, generator=g)
    d["e_proj"] = torch.randn(E, H * C, generator=g)
    d["att"] = torch.randn(1, H, C, generator=g) * 0.3
    d["bias"] = torch.randn(H * C, generator=g) * 0.1
    d["g_out"] = torch.randn(N, H * C, generator=g)
    if masked:
        if general_mask:
            d["mask"] = torch.rand(E, 1, generator=g) * 1.2
        else:
            d["mask"] = (torch.rand(E, 1, generator=g) > 0.4).float()
    else:
        d["mask"] = None
    return d


def _run_edge_oracle(d, H, C):
    import isg_oracle as O

    N = d["x_l"].shape[0]
    leaves = {k: d[k].clone().double().requires_grad_(True) for k in ("x_l", "x_r", "e_proj", "att", "bias")}
    m = d["mask"].clone().double().requires_grad_(True) if d["mask"] is not None else None
    out, alpha = O.gat_edge(leaves["x_l"].view(N, H, C), leaves["x_r"].view(N, H, C),
                            leaves["e_proj"].view(-1, H, C), leaves["att"], d["edge_index"], m)
    out = out.reshape(N, H * C) + leaves["bias"]
    out.backward(d["g_out"].double())
    res = dict(out=out.detach(), alpha=alpha.detach())
    for k, v in leaves.items():
        res["g_" + k] = v.grad
    res["g_mask"] = m.grad if m is not None else None
    return res


def _run_edge_cuda(d, H, C, fused_pitch=False):
    from isg_b200 import ops

    gi = _gi(d["edge_index"], d["batch"], int(d["batch"].max()) + 1)
    # the op takes [x_l | x_r] as one [N, 2HC] tensor (the output of the fused lin_l/lin_r projection)
    xlr = torch.cat([d["x_l"], d["x_r"]], dim=1).to(DEV).requires_grad_(True)
    ep = d["e_proj"].to(DEV).requires_grad_(True)
    att = d["att"].to(DEV).requires_grad_(True)
    bias = d["bias"].to(DEV).requires_grad_(True)
    m = d["mask"].to(DEV).requires_grad_(True) if d["mask"] is not None else None
    out, alpha = ops.GatEdge.apply(xlr, ep, att, bias, m, gi, H, 0.2)
    out.backward(d["g_out"].to(DEV))
    res = dict(out=out.detach(), alpha=alpha.detach(), g_e_proj=ep.grad, g_att=att.grad, g_bias=bias.grad,
               g_mask=m.grad if m is not None else None)
    res["g_x_l"], res["g_x_r"] = xlr.grad[:, : H * C], xlr.grad[:, H * C:]
    return res


@pytest.mark.parametrize("B,mn,me,C,H,masked,gen", [
    (4, 8, 40, 300, 4, False, False),
    (4, 8, 40, 300, 4, True, False),
    (4, 8, 40, 300, 4, True, True),
    (6, 10, 60, 8, 4, True, False),
    (3, 12, 70, 64, 2, False, False),
    (2, 6, 30, 512, 1, True, True),
    (2, 40, 900, 300, 4, True, False),  # in-degree > 32 -> multi-chunk online softmax
    (64, 20, 150, 300, 4, True, False),  # BASELINE config 1 size
    (4, 200, 4000, 300, 4, True, False),  # BASELINE config 5's high-degree end: 200 objects / 4000 edges per graph
    (3, 100, 1500, 300, 4, False, False),  # config 5, 100 objects / 1500 edges
])
def test_edge_fwd_bwd_matches_oracle(B, mn, me, C, H, masked, gen):
    d = _edge_case(B, mn, me, C, H, masked, seed=B * 100 + C, general_mask=gen)
    want = _run_edge_oracle(d, H, C)
    got = _run_edge_cuda(d, H, C)
    for k in ("out", "alpha", "g_x_l", "g_x_r", "g_e_proj", "g_att", "g_bias"):
        assert util.rel_err(got[k], want[k]) <= RTOL, (k, util.rel_err(got[k], want[k]))
    if masked:
        assert util.rel_err(got["g_mask"], want["g_mask"]) <= RTOL


@pytest.mark.parametrize("B,mn,me,masked", [(64, 20, 150, True), (256, 20, 150, False), (5, 60, 1200, True),
                                            (1, 9, 40, False)])
def test_edge_bwd_single_launch_is_bit_identical_to_two_pass(B, mn, me, masked):
    """The single-launch backward (dst / src block roles ordered by tickets, g_eproj consumed from L2) does the
    same arithmetic in the same order as the two-launch form: every output bit-identical, run to run as well."""
    from isg_b200 import ops

    d = _edge_case(B, mn, me, 300, 4, masked, seed=31 + B)
    keys = ("out", "alpha", "g_x_l", "g_x_r", "g_e_proj", "g_att", "g_bias") + (("g_mask",) if masked else ())
    prev = ops._EDGE_BWD_FUSED
    try:
        ops._EDGE_BWD_FUSED = False
        two = _run_edge_cuda(d, 4, 300)
        ops._EDGE_BWD_FUSED = True
        one = _run_edge_cuda(d, 4, 300)
        again = _run_edge_cuda(d, 4, 300)
    finally:
        ops._EDGE_BWD_FUSED = prev
    for k in keys:
        assert torch.equal(one[k], two[k]), k
        assert torch.equal(one[k], again[k]), k


@pytest.mark.parametrize("B,mn,me,masked", [(64, 20, 150, True), (256, 20, 150, False), (3, 300, 20000, False)])
def test_edge_task_order_longest_first_changes_no_bit(B, mn, me, masked):
    """isg_degree_order: the nodes with at least twice the mean degree first in node order, then the rest (ascending for
    the dst ordering, descending for the src ordering); the edge kernels scheduled in that order give bit-identical results to the natural order."""
    from isg_b200 import graph, ops

    d = _edge_case(B, mn, me, 300, 4, masked, seed=11 + B)
    gi = _gi(d["edge_index"], d["batch"], B)
    N = gi.N
    for side, order, ptr in (("dst", gi.dst_order, gi.dst_ptr), ("src", gi.src_order, gi.src_ptr)):
        assert torch.equal(order.cpu().long(), util.heavy_first_order(ptr.cpu(), N, gi.E, side))
    keys = ("out", "alpha", "g_x_l", "g_x_r", "g_e_proj", "g_att", "g_bias") + (("g_mask",) if masked else ())
    prev = ops._EDGE_BWD_FUSED
    try:
        ops._EDGE_BWD_FUSED = False  # the two-launch backward is the one that takes the order
        with_order = _run_edge_cuda(d, 4, 300)
        graph._EDGE_ORDER = False  # GraphIndex then carries no order: natural node order
        natural = _run_edge_cuda(d, 4, 300)
    finally:
        ops._EDGE_BWD_FUSED = prev
        graph._EDGE_ORDER = True
    for k in keys:
        assert torch.equal(with_order[k], natural[k]), k


def test_edge_bwd_falls_back_when_edges_leave_their_graph():
    """An edge between two graphs breaks the precondition of the single-launch backward; GraphIndex.closed
    reports it and the two-launch form runs (results still match the oracle)."""
    d = _edge_case(6, 10, 50, 300, 4, True, seed=5)
    ei = d["edge_index"].clone()
    ei[0, 3] = d["x_l"].shape[0] - 1  # source in the last graph, destination in the first
    d["edge_index"] = ei
    gi = _gi(ei, d["batch"], int(d["batch"].max()) + 1)
    assert gi.closed is False
    want = _run_edge_oracle(d, 4, 300)
    got = _run_edge_cuda(d, 4, 300)
    for k in ("out", "g_x_l", "g_x_r", "g_e_proj", "g_att", "g_mask"):
        assert util.rel_err(got[k], want[k]) <= RTOL, k
    assert _gi(_edge_case(6, 10, 50, 300, 4, True, seed=5)["edge_index"], d["batch"], 6).closed is True


def test_edge_fused_pitch_and_isolated_nodes():
    d = _edge_case(3, 9, 50, 300, 4, True, seed=77)
    # append two nodes without any edge: out must equal the bias there, grads zero
    N = d["x_l"].shape[0]
    g = torch.Generator().manual_seed(5)
    for k in ("x_l", "x_r", "g_out"):
        d[k] = torch.cat([d[k], torch.randn(2, d[k].shape[1], generator=g)])
    d["batch"] = torch.cat([d["batch"], torch.full((2,), int(d["batch"].max()))])
    want = _run_edge_oracle(d, 4, 300)
    got = _run_edge_cuda(d, 4, 300, fused_pitch=True)
    for k in ("out", "alpha", "g_x_l", "g_x_r", "g_e_proj", "g_att", "g_mask"):
        assert util.rel_err(got[k], want[k]) <= RTOL, k
    assert torch.allclose(got["out"][N:].cpu(), d["bias"].expand(2, -1))


@pytest.mark.parametrize("masked", [False, True])
@pytest.mark.parametrize("B,mn,me,H,C", [(12, 14, 90, 4, 300), (3, 40, 900, 4, 300), (5, 9, 40, 2, 64), (4, 10, 50, 3, 300)],
                         ids=["gqa", "deg>32", "H2C64", "H3-8byte-kernels"])
def test_edge_bf16_storage(masked, B, mn, me, H, C):
    """bf16 STORAGE of the edge-kernel tensors (x_l|x_r, e_proj, out and their gradients), fp32 accumulation
    inside the kernel — the configuration BASELINE states separately.  Reference = the oracle in fp64 on the
    bf16-rounded inputs; bound = bf16 output rounding (2^-8 relative) with margin.  Even head counts run the ring
    kernels (one warp per head pair, 16-byte accesses); H = 3 exercises the older 8-byte register-load kernels."""
    import isg_oracle as O
    from isg_b200 import ops

    d = _edge_case(B, mn, me, C, H, masked, seed=21)
    for k in ("x_l", "x_r", "e_proj", "g_out"):
        d[k] = d[k].to(torch.bfloat16).float()  # inputs exactly representable in bf16
    want = _run_edge_oracle(d, H, C)
    gi = _gi(d["edge_index"], d["batch"], int(d["batch"].max()) + 1)
    xlr = torch.cat([d["x_l"], d["x_r"]], dim=1).to(DEV, torch.bfloat16).requires_grad_(True)
    ep = d["e_proj"].to(DEV, torch.bfloat16).requires_grad_(True)
    att = d["att"].to(DEV).requires_grad_(True)
    bias = d["bias"].to(DEV).requires_grad_(True)
    m = d["mask"].to(DEV).requires_grad_(True) if d["mask"] is not None else None
    out, alpha = ops.GatEdge.apply(xlr, ep, att, bias, m, gi, H, 0.2)
    assert out.dtype == torch.bfloat16
    out.backward(d["g_out"].to(DEV, torch.bfloat16))
    tol = 1.2e-2
    assert util.rel_err(out.float(), want["out"]) <= tol
    assert util.rel_err(alpha, want["alpha"]) <= 1e-4  # alpha is fp32 and the inputs are bf16-exact
    assert util.rel_err(xlr.grad[:, : H * C].float(), want["g_x_l"]) <= tol
    assert util.rel_err(xlr.grad[:, H * C:].float(), want["g_x_r"]) <= tol
    assert util.rel_err(ep.grad.float(), want["g_e_proj"]) <= tol
    assert util.rel_err(att.grad, want["g_att"]) <= tol
    if masked:
        assert util.rel_err(m.grad, want["g_mask"]) <= tol


def test_edge_softmax_rows_sum_to_one_at_full_size():
    """BASELINE config 2 size (B=1024): size-independent property of the segment softmax."""
    d = _edge_case(1024, 20, 150, 300, 4, False, seed=1)
    got = _run_edge_cuda(d, 4, 300)
    alpha = got["alpha"].cpu()
    sums = torch.zeros(d["x_l"].shape[0], 4).index_add_(0, d["edge_index"][1], alpha)
    deg = torch.bincount(d["edge_index"][1], minlength=d["x_l"].shape[0])
    assert float((sums[deg > 0] - 1).abs().max()) < 1e-5
    assert bool(torch.isfinite(got["out"]).all())


def test_node_edge_mask_custom_backward():
    import isg_oracle as O
    from isg_b200 import ops

    t = synth.make_topology(5, 9, 50, seed=2)
    N, E = t["batch"].numel(), t["edge_index"].size(1)
    g = torch.Generator().manual_seed(2)
    m = torch.rand(N, 1, generator=g)
    ge = torch.randn(E, 1, generator=g)
    mo = m.clone().requires_grad_(True)
    eo = O.node_mask_to_edge_mask(mo, t["edge_index"])
    eo.backward(ge)
    gi = _gi(t["edge_index"], t["batch"], 5)
    mc = m.to(DEV).requires_grad_(True)
    ec = ops.NodeMaskToEdgeMaskFn.apply(mc, gi)
    ec.backward(ge.to(DEV))
    assert torch.equal(ec.detach().cpu(), eo.detach())  # one fp32 multiply: bit-exact
    assert util.rel_err(mc.grad, mo.grad) <= 1e-6


# ------------------------------------------------------------------------------------------ (c) samplers
def _ragged(B, mn, seed, ties=False):
    g = torch.Generator().manual_seed(seed)
    counts = torch.randint(1, 2 * mn, (B,), generator=g)
    batch = torch.repeat_interleave(torch.arange(B), counts)
    theta = torch.randn(batch.numel(), 1, generator=g)
    if ties:
        theta = torch.round(theta * 2) / 2  # many duplicates -> exercises the >= threshold semantics
    return batch, theta, int(counts.max())


@pytest.mark.parametrize("k", [1, 2, 3, 5, 40])
@pytest.mark.parametrize("ties", [False, True])
def test_topk_mask_bit_exact(k, ties):
    import isg_oracle as O
    from isg_b200 import ops

    B = 37
    batch, theta, nmax = _ragged(B, 9, seed=k + 10 * ties, ties=ties)
    noise = synth.gumbel_noise(B, nmax, 0.3, seed=k)
    if ties:
        noise = torch.zeros_like(noise)
    dense, valid = O.to_dense_batch(theta, batch, B)
    want = O.ImleFn.apply(dense, noise, k, 1.0, 10.0, 1.0, 1.0).squeeze(0)[valid]
    gi = _gi(torch.zeros(2, 0, dtype=torch.int64), batch, B)
    got = ops.TopkImle.apply(theta.to(DEV), noise.to(DEV), gi, k, 1.0, 10.0, 1.0, 1.0)
    assert torch.equal(got.cpu(), want)


def test_topk_pads_compete_and_negative_scores():
    """Quirk Q2: zero pads take part in top-k, so a graph of negative logits can get 0 real nodes."""
    import isg_oracle as O
    from isg_b200 import ops

    batch = torch.tensor([0, 0, 0, 0, 0, 1, 1])
    theta = torch.tensor([[-1.0], [-2.0], [-0.5], [-3.0], [-4.0], [-1.0], [-2.0]])
    dense, valid = O.to_dense_batch(theta, batch, 2)
    noise = torch.zeros(2, 1, 5, 1)
    want = O.ImleFn.apply(dense, noise, 2, 1.0, 10.0, 1.0, 1.0).squeeze(0)[valid]
    gi = _gi(torch.zeros(2, 0, dtype=torch.int64), batch, 2)
    got = ops.TopkImle.apply(theta.to(DEV), noise.to(DEV), gi, 2, 1.0, 10.0, 1.0, 1.0)
    assert torch.equal(got.cpu(), want)
    assert got[5:].sum().item() == 0.0  # graph 1: the three zero pads win (ties at 0 select all pads)


@pytest.mark.parametrize("k", [2, 4])
def test_imle_backward_bit_exact(k):
    import isg_oracle as O
    from isg_b200 import ops

    B = 29
    batch, theta, nmax = _ragged(B, 10, seed=3 + k)
    noise = synth.gumbel_noise(B, nmax, 0.3, seed=4)
    g = torch.Generator().manual_seed(8)
    dy = torch.randn(theta.shape, generator=g) * 0.3
    to = theta.clone().requires_grad_(True)
    dense, valid = O.to_dense_batch(to, batch, B)
    out = O.ImleFn.apply(dense, noise, k, 1.0, 10.0, 1.0, 1.0).squeeze(0)[valid]
    out.backward(dy)
    gi = _gi(torch.zeros(2, 0, dtype=torch.int64), batch, B)
    tc = theta.to(DEV).requires_grad_(True)
    oc = ops.TopkImle.apply(tc, noise.to(DEV), gi, k, 1.0, 10.0, 1.0, 1.0)
    oc.backward(dy.to(DEV))
    assert torch.equal(oc.detach().cpu(), out.detach())
    assert torch.equal(tc.grad.cpu(), to.grad)
    assert float(to.grad.abs().sum()) > 0  # the perturbation actually moves the MAP state


def test_aimle_backward_and_adaptive_state():
    import isg_oracle as O
    from isg_b200 import ops

    B, k = 31, 2
    batch, theta, nmax = _ragged(B, 10, seed=21)
    gi = _gi(torch.zeros(2, 0, dtype=torch.int64), batch, B)
    st_o = O.AimleState(alpha=1.0, beta=3.0)
    st_c = torch.tensor([3.0, 1.0, 0.0, 1.0, 1e-4, 0.9, 1.0, 0.0], dtype=torch.float64, device=DEV)
    g = torch.Generator().manual_seed(5)
    nonzero = 0
    for step in range(4):
        noise = synth.gumbel_noise(B, nmax, 0.3, seed=50 + step)
        dy = torch.randn(theta.shape, generator=g)
        to = theta.clone().requires_grad_(True)
        dense, valid = O.to_dense_batch(to, batch, B)
        out = O.AimleFn.apply(dense, noise, k, st_o, 1.0, 1.0)[valid]
        out.backward(dy)
        tc = theta.to(DEV).requires_grad_(True)
        oc = ops.TopkAimle.apply(tc, noise.to(DEV), gi, k, st_c, True, 1.0, 1.0)
        oc.backward(dy.to(DEV))
        assert torch.equal(oc.detach().cpu(), out.detach())
        assert util.rel_err(tc.grad, to.grad) <= 1e-5, step
        nonzero += int(to.grad.abs().sum() > 0)
        s = st_c.cpu().tolist()
        assert abs(s[0] - st_o.beta) < 1e-12, (s[0], st_o.beta)
        assert abs(s[1] - float(st_o.grad_norm)) < 1e-6
    assert nonzero >= 3


def test_gumbel_topk_fwd_bwd():
    import isg_oracle as O
    from isg_b200 import ops

    B, k = 23, 3
    batch, theta, nmax = _ragged(B, 9, seed=33)
    gum = synth.gumbel_noise(B, nmax, 1.0, seed=6)[:, 0, :, 0].contiguous()
    g = torch.Generator().manual_seed(9)
    dy = torch.randn(theta.shape, generator=g)
    to = theta.clone().requires_grad_(True)
    dense, valid = O.to_dense_batch(to, batch, B)
    out = O.gumbel_topk(dense, gum, k).squeeze(0)[valid]
    out.backward(dy)
    gi = _gi(torch.zeros(2, 0, dtype=torch.int64), batch, B)
    tc = theta.to(DEV).requires_grad_(True)
    oc = ops.GumbelTopk.apply(tc, gum.to(DEV), gi, k, 0.1)
    oc.backward(dy.to(DEV))
    assert float((oc.detach().cpu() - out.detach()).abs().max()) <= 1e-5
    assert util.rel_err(tc.grad, to.grad) <= 2e-4


# ------------------------------------------------------------------------------------------ node-side ops
def _node_case(B, mn, D, seed, with_empty=False):
    g = torch.Generator().manual_seed(seed)
    counts = torch.randint(1, 2 * mn, (B,), generator=g)
    if with_empty:
        counts[1] = 0
        counts[B - 1] = 1
    batch = torch.repeat_interleave(torch.arange(B), counts)
    N = batch.numel()
    return batch, N, g


@pytest.mark.parametrize("D", [8, 300])
def test_instr_gate(D):
    import isg_oracle as O
    from isg_b200 import ops

    B = 11
    batch, N, g = _node_case(B, 7, D, 1)
    x, ins, gy = torch.randn(N, D, generator=g), torch.randn(B, D, generator=g), torch.randn(N, D, generator=g)
    xo, io = x.clone().requires_grad_(True), ins.clone().requires_grad_(True)
    yo = O.instr_gate(xo, io, batch)
    yo.backward(gy)
    gi = _gi(torch.zeros(2, 0, dtype=torch.int64), batch, B)
    xc, ic = x.to(DEV).requires_grad_(True), ins.to(DEV).requires_grad_(True)
    yc = ops.InstrGate.apply(xc, ic, gi)
    yc.backward(gy.to(DEV))
    assert util.rel_err(yc, yo) <= 1e-5
    assert util.rel_err(xc.grad, xo.grad) <= RTOL
    assert util.rel_err(ic.grad, io.grad) <= RTOL


def test_colsum_multi_is_bit_identical_to_colsum():
    """isg_colsum_multi (the layer backward's batched bias-gradient sums) == one isg_colsum per tensor, bit for bit,
    including a zero-row job and a pitched input."""
    import ctypes

    from isg_b200 import lib as L

    lib = L.load()
    g = torch.Generator().manual_seed(5)
    shapes = [(256, 300), (4910, 300), (4910, 600), (4910, 1200), (4910, 2400), (0, 300), (1, 300), (65, 304)]
    wide = torch.randn(4910, 2400, generator=g).to(DEV)
    ins = [torch.randn(r, c, generator=g).to(DEV) for r, c in shapes]
    ins[3] = wide[:, 1200:]  # pitched view: ld = 2400
    n = len(ins)
    rows = np.array([t.shape[0] for t in ins], dtype=np.int64)
    cols = np.array([t.shape[1] for t in ins], dtype=np.int32)
    ld = np.array([t.stride(0) if t.shape[0] else t.shape[1] for t in ins], dtype=np.int64)
    outs = [torch.full((c,), float("nan"), device=DEV) for c in cols]
    pin = (ctypes.c_void_p * n)(*[t.data_ptr() for t in ins])
    pout = (ctypes.c_void_p * n)(*[t.data_ptr() for t in outs])
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    nbytes = lib.isg_colsum_multi_workspace_bytes(n, vp(rows), vp(cols))
    ws = L.workspace(nbytes, DEV)
    L.call("isg_colsum_multi", n, pin, None, vp(ld), vp(rows), vp(cols), pout, L.ptr(ws), nbytes, L.stream())
    for t, o in zip(ins, outs):
        r, c = t.shape
        want = torch.empty(c, device=DEV)
        nb = lib.isg_colsum_workspace_bytes(r, c)
        w1 = L.workspace(nb, DEV)
        L.call("isg_colsum", L.ptr(t) if r else None, t.stride(0) if r else c, r, c, L.ptr(want), L.ptr(w1), nb, L.stream())
        assert torch.equal(o, want), (r, c)
        if r:
            assert util.rel_err(o.cpu(), t.double().sum(0).float().cpu()) <= 1e-5
    with pytest.raises(RuntimeError):  # too small a workspace is reported, not overrun
        L.call("isg_colsum_multi", n, pin, None, vp(ld), vp(rows), vp(cols), pout, L.ptr(ws), 16, L.stream())
    # bf16 storage (the bf16 configuration's g_out / g_xlr / g_z1): same sums of the rounded values
    tb = [t.to(torch.bfloat16) for t in (ins[1], ins[4])]
    ob = [torch.empty(t.shape[1], device=DEV) for t in tb]
    dts = np.array([L.BF16, L.BF16], dtype=np.int32)
    r2 = np.array([t.shape[0] for t in tb], dtype=np.int64)
    c2 = np.array([t.shape[1] for t in tb], dtype=np.int32)
    l2 = np.array([t.stride(0) for t in tb], dtype=np.int64)
    L.call("isg_colsum_multi", 2, (ctypes.c_void_p * 2)(*[t.data_ptr() for t in tb]), vp(dts), vp(l2), vp(r2), vp(c2),
           (ctypes.c_void_p * 2)(*[t.data_ptr() for t in ob]), L.ptr(ws), nbytes, L.stream())
    for t, o in zip(tb, ob):
        assert util.rel_err(o.cpu(), t.double().sum(0).float().cpu()) <= 1e-5


def test_instr_gate_bwd_residual_and_accumulate_at_gqa_size():
    """isg_instr_gate_bwd's vectorised kernel (float4 columns x 4 row lanes) with the residual gradient and the
    accumulate flag the layer executor uses, against fp64 autograd."""
    from isg_b200 import lib as L

    B, D = 40, 300
    batch, N, g = _node_case(B, 25, D, 3)
    x, ins = torch.randn(N, D, generator=g), torch.randn(B, D, generator=g)
    gy, gres, gins0 = torch.randn(N, D, generator=g), torch.randn(N, D, generator=g), torch.randn(B, D, generator=g)
    xd, idd = x.double().requires_grad_(True), ins.double().requires_grad_(True)
    y = torch.nn.functional.gelu(xd * idd[batch])
    y.backward(gy.double())
    gi = _gi(torch.zeros(2, 0, dtype=torch.int64), batch, B)
    gx = torch.empty(N, D, device=DEV)
    gins = gins0.to(DEV).clone()
    gy_d, x_d, ins_d, gres_d = (t.to(DEV) for t in (gy, x, ins, gres))  # keep the device copies alive over the call
    L.call("isg_instr_gate_bwd", L.ptr(gy_d), L.ptr(x_d), L.ptr(ins_d), L.ptr(gi.graph_ptr), B, D,
           L.ptr(gres_d), 1, L.ptr(gx), L.ptr(gins), L.stream())
    torch.cuda.synchronize()
    assert util.rel_err(gx.cpu(), (xd.grad + gres.double()).float()) <= 1e-5
    assert util.rel_err(gins.cpu(), (idd.grad + gins0.double()).float()) <= 1e-5


@pytest.mark.parametrize("D", [16, 300])
def test_gate_theta_double_gather(D):
    """Quirk Q1: node n reads the question row of graph batch[batch[n]]."""
    import isg_oracle as O
    from isg_b200 import lib as L
    from isg_b200 import ops

    B = 9
    batch, N, g = _node_case(B, 6, D, 2)
    x, u, gth = torch.randn(N, D, generator=g), torch.randn(B, D, generator=g), torch.randn(N, 1, generator=g)
    wn, bn = torch.randn(D, D, generator=g) / math.sqrt(D), torch.randn(D, generator=g) * 0.1
    wq, bq = torch.randn(D, D, generator=g) / math.sqrt(D), torch.randn(D, generator=g) * 0.1
    xo, uo = x.clone().requires_grad_(True), u.clone().requires_grad_(True)
    th_o = O.masking_theta(xo, uo[batch], batch, wn, bn, wq, bq)
    th_o.backward(gth)
    gi = _gi(torch.zeros(2, 0, dtype=torch.int64), batch, B)
    xc, uc = x.to(DEV).requires_grad_(True), u.to(DEV).requires_grad_(True)
    xn = ops.linear(xc, wn.to(DEV), bn.to(DEV), L.ACT_GELU)
    q = ops.linear(uc, wq.to(DEV), bq.to(DEV), L.ACT_GELU)
    th_c = ops.GateTheta.apply(xn, q, gi, True)
    th_c.backward(gth.to(DEV))
    assert util.rel_err(th_c, th_o) <= RTOL
    assert util.rel_err(xc.grad, xo.grad) <= RTOL
    assert util.rel_err(uc.grad, uo.grad) <= RTOL
    # single-gather mode (MaskingModel.forward called with a per-node u)
    un = u[batch]
    uo2 = un.clone().requires_grad_(True)
    th_o2 = O.masking_theta(x, uo2, batch, wn, bn, wq, bq)
    th_o2.backward(gth)
    uc2 = un.to(DEV).requires_grad_(True)
    q2 = ops.linear(uc2, wq.to(DEV), bq.to(DEV), L.ACT_GELU)
    th_c2 = ops.GateTheta.apply(xn.detach(), q2, gi, False)
    th_c2.backward(gth.to(DEV))
    assert util.rel_err(th_c2, th_o2) <= RTOL
    assert util.rel_err(uc2.grad, uo2.grad) <= RTOL


@pytest.mark.parametrize("D,with_empty", [(300, False), (300, True), (12, False)])
def test_sdpa_graphnorm_residual(D, with_empty):
    import isg_oracle as O
    from isg_b200 import ops

    B = 10
    batch, N, g = _node_case(B, 8, D, 3, with_empty)
    v, ins, h = torch.randn(N, D, generator=g), torch.randn(B, D, generator=g), torch.randn(N, D, generator=g)
    w, b, ms = 1 + 0.1 * torch.randn(D, generator=g), 0.1 * torch.randn(D, generator=g), 1 + 0.1 * torch.randn(
        D, generator=g)
    go = torch.randn(N, D, generator=g)
    leaves_o = [t.clone().double().requires_grad_(True) for t in (v, ins, h, w, b, ms)]
    y = O.scatter_sdpa(leaves_o[1], leaves_o[0], leaves_o[0], batch)
    out_o = O.graph_norm(y, batch, leaves_o[3], leaves_o[4], leaves_o[5], B) + leaves_o[2]
    out_o.backward(go.double())
    gi = _gi(torch.zeros(2, 0, dtype=torch.int64), batch, B)
    leaves_c = [t.to(DEV).requires_grad_(True) for t in (v, ins, h, w, b, ms)]
    out_c = ops.SdpaGraphNormResidual.apply(*leaves_c, gi, 1e-5)
    out_c.backward(go.to(DEV))
    assert util.rel_err(out_c, out_o) <= RTOL
    for name, lc, lo in zip(("v", "ins", "h", "w", "b", "ms"), leaves_c, leaves_o):
        assert util.rel_err(lc.grad, lo.grad) <= RTOL, name


@pytest.mark.parametrize("B,mn,k,zeros", [(9, 6, 2, True), (16, 20, 2, True), (5, 40, 3, True), (7, 12, 5, True),
                                           (33, 20, 2, False), (3, 100, 4, True)])
def test_simple_marginals_fwd_bwd(B, mn, k, zeros):
    """SIMPLE sampler kernel vs the oracle's restatement of simple.py's circuit (pinned to the live reference,
    incl. the -1000 dummy-pad regime that exact-zero logits — dense pads, dropout — put it in)."""
    import isg_oracle as O
    from isg_b200 import lib as L
    from isg_b200 import ops

    batch, theta, nmax = _ragged(B, mn, seed=300 + B + k)
    g = torch.Generator().manual_seed(B * 7 + k)
    if zeros:
        theta[torch.rand(theta.shape, generator=g) < 0.2] = 0.0  # dropout-style exact zeros
    npad = L.load().isg_simple_npad(nmax)
    gum = synth.gumbel_noise(B, npad, 1.0, seed=77 + B)[:, 0, :, 0].contiguous()
    w = torch.randn(theta.shape[0], 1, generator=g)
    wm = torch.randn(B, nmax, generator=g)
    # oracle on the dense layout (to_dense_batch pads are exact zeros, too)
    to = theta.clone().requires_grad_(True)
    dense, valid = O.to_dense_batch(to, batch, B)
    mo, margo = O.simple_sample(dense, gum, k)
    mo_r = mo.squeeze(0)[valid]
    ((mo_r * w).sum() + (margo[..., 0] * wm).sum()).backward()
    # CUDA, ragged layout
    tc = theta.clone().to(DEV).requires_grad_(True)
    gi = _gi(torch.zeros(2, 0, dtype=torch.int64), batch, B)
    mc, margc = ops.SimpleTopk.apply(tc, gum.to(DEV), gi, k)
    ((mc * w.to(DEV)).sum() + (margc * wm.to(DEV)).sum()).backward()
    assert util.rel_err(margc, margo[..., 0]) <= 1e-5
    assert util.rel_err(mc, mo_r) <= 1e-5
    assert torch.equal(mc.detach().cpu().round(), mo_r.detach().round())  # same hard sample given the same noise
    assert util.rel_err(tc.grad, to.grad) <= 1e-4


@pytest.mark.parametrize("D,masked,with_empty", [(300, True, False), (300, False, True), (16, True, True)])
def test_global_attention_pooling(D, masked, with_empty):
    """the drop-in GlobalAttention module (projections + fused masked softmax-pool kernel)
    against the oracle restatement of models/att_pooling.py:57-77, forward and all gradients (incl. the gate)."""
    import isg_oracle as O
    from isg_b200.isubgvqa import GlobalAttention

    B = 11
    g = torch.Generator().manual_seed(D + 5)
    counts = torch.randint(1, 30, (B,), generator=g)
    if with_empty:
        counts[3] = 0
    batch = torch.repeat_interleave(torch.arange(B), counts)
    N = int(counts.sum())
    mod = GlobalAttention(num_node_features=D, num_out_features=D)
    params = {k: v.detach().clone() for k, v in mod.state_dict().items()}
    x = torch.randn(N, D, generator=g)
    u = torch.randn(B, D, generator=g)
    mask = (torch.rand(N, 1, generator=g) > 0.3).float() if masked else None
    wo, wg = torch.randn(B, D, generator=g), torch.randn(N, 1, generator=g)
    xo, uo = x.clone().requires_grad_(True), u.clone().requires_grad_(True)
    mo = mask.clone().requires_grad_(True) if masked else None
    po = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    oo, go = O.global_attention_pool(xo, uo, batch, B, po, node_mask=mo)
    ((oo * wo).sum() + (go * wg).sum()).backward()
    mod = mod.to(DEV)
    xc, uc = x.to(DEV).requires_grad_(True), u.to(DEV).requires_grad_(True)
    mc = mask.to(DEV).requires_grad_(True) if masked else None
    oc, gc = mod(xc, uc, batch.to(DEV), size=B, return_mask=True, node_mask=mc)
    ((oc * wo.to(DEV)).sum() + (gc * wg.to(DEV)).sum()).backward()
    assert util.rel_err(oc, oo) <= RTOL and util.rel_err(gc, go) <= RTOL
    assert util.rel_err(xc.grad, xo.grad) <= RTOL and util.rel_err(uc.grad, uo.grad) <= RTOL
    if masked:
        assert util.rel_err(mc.grad, mo.grad) <= RTOL
    for name, p in mod.named_parameters():
        if name.startswith("gate_nn"):
            assert p.grad is None  # built, never used (as in the reference)
        else:
            assert util.rel_err(p.grad, po[name].grad) <= RTOL, name


# ------------------------------------------------------------------------------------------ (d) projections
@pytest.mark.parametrize("M,K,Nout,act", [(1, 300, 1200, 0), (37, 300, 1200, 0), (515, 1200, 600, 1),
                                          (130, 600, 300, 1), (9600, 300, 1200, 0), (64, 8, 32, 1),
                                          (257, 300, 300, 1)])
@pytest.mark.parametrize("mode,tol", [(0, 1e-5), (1, 1e-5), (2, 3e-3)], ids=["ffma", "tc3xtf32", "tc1xtf32"])
def test_linear_fwd_dgrad_wgrad(M, K, Nout, act, mode, tol):
    """mode 0: fp32 FFMA; mode 1: tcgen05 3xTF32 split (fp32-grade, same bound); mode 2: single-pass TF32."""
    from isg_b200 import ops

    g = torch.Generator().manual_seed(M + K)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(Nout, K, generator=g) / math.sqrt(K)
    b = torch.randn(Nout, generator=g) * 0.1
    gy = torch.randn(M, Nout, generator=g)
    xo, wo, bo = (t.clone().double().requires_grad_(True) for t in (x, w, b))
    yo = torch.nn.functional.linear(xo, wo, bo)
    if act:
        yo = torch.nn.functional.gelu(yo)
    yo.backward(gy.double())
    xc, wc, bc = (t.to(DEV).requires_grad_(True) for t in (x, w, b))
    yc = ops.linear(xc, wc, bc, act, mode)
    yc.backward(gy.to(DEV))
    assert util.rel_err(yc, yo) <= tol
    assert util.rel_err(xc.grad, xo.grad) <= tol
    assert util.rel_err(wc.grad, wo.grad) <= tol
    assert util.rel_err(bc.grad, bo.grad) <= max(tol, 1e-5)


def test_linear_tc_3xtf32_error_bound_vs_fp64():
    """The default projection mode is documented as fp32-grade: 1.4-2.0e-6 against fp64 for the three products of a
    [9600,300]x[300,1200] layer (include/isg.h); the FFMA kernel sits at 0.8-1.7e-6."""
    from isg_b200 import ops

    g = torch.Generator().manual_seed(5)
    M, K, Nout = 9600, 300, 1200
    x = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(Nout, K, generator=g) / math.sqrt(K)).to(DEV)
    gy = torch.randn(M, Nout, generator=g).to(DEV)
    x64, w64, gy64 = x.double(), w.double(), gy.double()
    y, _ = ops.linear_fwd_raw(x, w, None, 0, False, mode=1)
    gx = ops.linear_dgrad_raw(gy, w, mode=1)
    gw = ops.linear_wgrad_raw(gy, x, mode=1)
    for got, ref in ((y, x64 @ w64.t()), (gx, gy64 @ w64), (gw, gy64.t() @ x64)):
        assert util.rel_err(got, ref) <= 3e-6


@pytest.mark.parametrize("M,K,Nout", [(1500, 300, 1200), (4910, 1200, 600), (257, 300, 300)])
def test_linear_tc_presplit_weight_lo_is_bit_identical(M, K, Nout):
    """isg_split_lo / isg_transpose_split + the pre-split weight arguments of isg_linear_dgrad / _fwd (lo plane fetched
    by TMA; forward on the transposed, MN-major weight) must give exactly the bits of the in-kernel split: same
    products, same accumulation order."""
    from isg_b200 import lib as L
    from isg_b200 import ops

    g = torch.Generator().manual_seed(M + Nout)
    x = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(Nout, K, generator=g) / K ** 0.5).to(DEV)
    b = torch.randn(Nout, generator=g).to(DEV)
    gy = torch.randn(M, Nout, generator=g).to(DEV)
    w_lo = ops.split_lo(w)
    hi = (w.view(torch.int32) & -8192).view(torch.float32)
    assert torch.equal(w_lo, w - hi) and torch.equal(hi + w_lo, w)
    y0, z0 = ops.linear_fwd_raw(x, w, b, L.ACT_GELU, True, mode=1)
    w_t, w_t_lo = ops.transpose_split(w)
    assert torch.equal(w_t, w.t().contiguous()) and torch.equal(w_t_lo, w_lo.t().contiguous())
    y1, z1 = ops.linear_fwd_raw(x, w, b, L.ACT_GELU, True, mode=1, w_t=w_t, w_t_lo=w_t_lo)
    assert torch.equal(y0, y1) and torch.equal(z0, z1)
    gx0 = ops.linear_dgrad_raw(gy, w, mode=1)
    gx1 = ops.linear_dgrad_raw(gy, w, mode=1, w_lo=w_lo)
    assert torch.equal(gx0, gx1)


def test_linear_tc_pitched_views_and_long_reduction():
    """tcgen05 path on column views of a wider buffer (x_l | x_r share one pitch) and on a reduction long
    enough to need many wgrad splits (E = 40k rows, BASELINE config 3 size)."""
    from isg_b200 import ops

    g = torch.Generator().manual_seed(11)
    M, K, Nout = 40000, 300, 1200
    xw = torch.randn(M, 2 * K, generator=g).to(DEV)
    x = xw[:, K:]  # pitched view, 16-byte aligned offset
    w = (torch.randn(Nout, K, generator=g) / math.sqrt(K)).to(DEV)
    gy = torch.randn(M, Nout, generator=g).to(DEV)
    y, _ = ops.linear_fwd_raw(x, w, None, 0, False, mode=1)
    assert util.rel_err(y, x.double() @ w.double().t()) <= 1e-5
    gw = ops.linear_wgrad_raw(gy, x, mode=1)
    assert util.rel_err(gw, gy.double().t() @ x.double()) <= 1e-5
    gx = ops.linear_dgrad_raw(gy, w, mode=1)
    assert util.rel_err(gx, gy.double() @ w.double()) <= 1e-5
    # deterministic: the split reduction has a fixed order
    assert torch.equal(gw, ops.linear_wgrad_raw(gy, x, mode=1))


def test_ops_reject_cpu_tensors():
    from isg_b200 import ops

    with pytest.raises(RuntimeError):
        ops.linear(torch.randn(4, 8), torch.randn(8, 8).to(DEV), None, 0)
