"""torch.autograd.Function wrappers over the libisg.so C ABI (include/isg.h).  PyTorch is plumbing
here — it owns device memory, streams and the autograd tape; all arithmetic happens in the CUDA
kernels.  Each Function cites the reference code it replaces (paths relative to the reference)."""
import os

import torch

from . import lib as L

# Edge backward as one launch (dst / src block roles, g_eproj consumed from L2) or as two launches.  Measured on
# B200 (profiles/r2b_edge_fused_vs_two_pass.txt): the single launch wins where segments are long — 12 % / 19 % /
# 25 % at 12 / 15 / 20 edges per node (batch 4096) — and loses 9-28 % at GQA's 5-8 edges per node, where both forms
# are bound by per-task start-up latency rather than DRAM traffic.  None = pick by mean degree; True / False force.
_EDGE_BWD_FUSED = {"1": True, "0": False}.get(os.environ.get("ISG_EDGE_BWD_FUSED", ""), None)
EDGE_BWD_FUSED_MIN_DEGREE = 10.0


def edge_bwd_fused(gi):
    """Whether GatEdge.backward runs the single-launch form for this batch (needs every edge inside one graph)."""
    want = _EDGE_BWD_FUSED if _EDGE_BWD_FUSED is not None else gi.E >= EDGE_BWD_FUSED_MIN_DEGREE * max(gi.N, 1)
    return bool(want) and gi.closed
_DEBUG_EDGE_BWD = None  # diagnostics: set to a list to capture GatEdge.backward inputs/outputs
# projection arithmetic: 1 = tcgen05 3xTF32 with bounded accumulation chains (default; measured 6.5e-7 relative
# against fp64, i.e. at or below the FFMA kernel's error), 0 = fp32 FFMA, 2 = tcgen05 single-pass TF32 (~8e-4)
_GEMM_MODE = int(os.environ.get("ISG_GEMM_MODE", "1"))


def set_gemm_mode(mode):
    global _GEMM_MODE
    _GEMM_MODE = int(mode)


def gemm_mode():
    return _GEMM_MODE


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------------------------------
# dense projections (kernel d)
# ------------------------------------------------------------------------------------------
# Pre-split weight planes (isg_split_lo / isg_transpose_split, one ~3 us kernel per weight and step): the GEMM fetches
# the weight's lo plane by TMA instead of splitting the weight tile inside its k-loop.  With the 16-wide k-block
# kernel (-DISG_TC_NARROW) this paid off from ~1000 rows (dgrad 216 -> 202 us; forward on the transposed, MN-major
# weight 202 -> 179 us on [39809,300]x[300,1200]); with the default 32-wide kernel the operand rows are 128 bytes
# either way and the extra L2 traffic and launches cost slightly more than the split saves (5.17-5.21 vs 5.24-5.29
# ms per c3 step), so it is off unless ISG_PRESPLIT_MIN_ROWS is set.
PRESPLIT_MIN_ROWS = int(os.environ.get("ISG_PRESPLIT_MIN_ROWS", str(1 << 62)))


def split_lo(w):
    """w - tf32_trunc(w) (exact), the lo plane of the 3xTF32 split of a dense fp32 weight."""
    w_lo = torch.empty_like(w)
    L.call("isg_split_lo", L.ptr(w), w.numel(), L.ptr(w_lo), L.stream())
    return w_lo


def transpose_split(w):
    """(w^T, w^T - tf32_trunc(w^T)) as dense [K, Nout] tensors: the forward product's pre-split MN-major weight."""
    Nout, K = w.shape
    w_t = torch.empty(K, Nout, dtype=torch.float32, device=w.device)
    w_t_lo = torch.empty(K, Nout, dtype=torch.float32, device=w.device)
    L.call("isg_transpose_split", L.ptr(w), Nout, K, L.ptr(w_t), L.ptr(w_t_lo), L.stream())
    return w_t, w_t_lo


def linear_fwd_raw(x, w, b, act, want_pre, mode=None, w_t=None, w_t_lo=None):
    lib = L.load()
    mode = _GEMM_MODE if mode is None else mode
    M, K = x.shape
    Nout = w.shape[0]
    y = torch.empty(M, Nout, dtype=x.dtype, device=x.device)
    z = torch.empty(M, Nout, dtype=x.dtype, device=x.device) if want_pre else None
    L.call("isg_linear_fwd", L.ptr(x), x.stride(0), L.ptr(w), L.ptr(w_t), L.ptr(w_t_lo), L.ptr(b), L.ptr(y), Nout,
                               L.ptr(z), Nout, M, Nout, K, act, mode, L.dtype_code(x), L.stream())
    return y, z


def linear_dgrad_raw(gy, w, z_prev=None, out=None, accumulate=False, mode=None, w_lo=None):
    lib = L.load()
    mode = _GEMM_MODE if mode is None else mode
    M, Nout = gy.shape
    K = w.shape[1]
    gx = out if out is not None else torch.empty(M, K, dtype=gy.dtype, device=gy.device)
    L.call("isg_linear_dgrad", L.ptr(gy), gy.stride(0), L.ptr(w), L.ptr(w_lo), L.ptr(z_prev), K, L.ptr(gx), gx.stride(0),
                                 1 if accumulate else 0, M, Nout, K, mode, L.dtype_code(gy), L.stream())
    return gx


def linear_wgrad_raw(gy, x, mode=None):
    lib = L.load()
    mode = _GEMM_MODE if mode is None else mode
    M, Nout = gy.shape
    K = x.shape[1]
    gw = torch.empty(Nout, K, dtype=torch.float32, device=gy.device)
    nbytes = lib.isg_linear_wgrad_workspace_bytes(M, Nout, K)
    ws = L.workspace(nbytes, gy.device)
    L.call("isg_linear_wgrad", L.ptr(gy), gy.stride(0), L.ptr(x), x.stride(0), L.ptr(gw), None, M, Nout, K,
                                 mode, L.dtype_code(gy), L.ptr(ws), nbytes, L.stream())
    return gw


# Bias / affine-parameter gradients are column sums of tensors the backward pass has already produced; nothing
# on the critical path waits for them.  During a backward pass they are issued on a side stream (they are tiny,
# latency-bound launches that co-reside with the persistent GEMM CTAs) and joined back when autograd finishes.
_SIDE_STREAM_ENABLED = os.environ.get("ISG_SIDE_STREAM", "1") != "0"
_side_streams = {}
_join_pending = False
_side_ok = False  # set per step by MGAT.forward: only when no parameter has a .grad to accumulate into


_side_forbidden = 0  # >0 while some consumer reads gradients DURING backward (see forbid_side_stream)
_side_dist_ok = False


def forbid_side_stream(on):
    """A consumer that reads parameter gradients from hooks DURING the backward pass (DistributedDataParallel's
    reducer, isg_b200.dp.OverlappedGradAllReduce) would see a side-stream gradient before it is joined.  Such a
    consumer calls forbid_side_stream(True) for its lifetime and forbid_side_stream(False) when it goes away;
    the ban is a counter, independent of whether a process group exists."""
    global _side_forbidden
    _side_forbidden = max(0, _side_forbidden + (1 if on else -1))


def set_side_stream_with_dist(ok):
    """With a torch.distributed process group present and nobody vouching for it, the side stream stays off (a
    stock DistributedDataParallel wrapper may be reading gradients during backward).  isg_b200.dp.GradAllReduce
    reads gradients only after backward() has returned and joined, and says so here."""
    global _side_dist_ok
    _side_dist_ok = bool(ok)


def allow_side_stream(ok):
    """The side stream hands autograd a gradient produced off the main stream.  That is safe when the
    AccumulateGrad node merely stores it (param.grad is None, the state after zero_grad()); an in-place
    accumulation into an existing .grad would run on the main stream without waiting — so the caller
    vouches for the former once per step."""
    global _side_ok
    if _side_forbidden > 0:
        ok = False
    elif ok and not _side_dist_ok and torch.distributed.is_available() and torch.distributed.is_initialized():
        ok = False
    _side_ok = bool(ok)



def _side_stream(device):
    st = _side_streams.get(device)
    if st is None:
        st = _side_streams[device] = torch.cuda.Stream(device=device)
    return st


def join_side_stream():
    """Make the current stream wait for every column sum issued on the side stream.  A no-op unless a backward
    pass has issued side-stream work that was not joined yet (and never during CUDA-graph capture, where the
    side stream is not used and waiting on an uncaptured stream would invalidate the capture)."""
    global _join_pending
    if not _join_pending or torch.cuda.is_current_stream_capturing():
        return
    _join_pending = False
    for st in _side_streams.values():
        torch.cuda.current_stream(st.device).wait_stream(st)


def _defer_join():
    """True if we are inside a backward pass and a join has been queued for its end."""
    global _join_pending
    if _join_pending:
        return True
    try:
        torch.autograd.Variable._execution_engine.queue_callback(join_side_stream)
    except RuntimeError:  # not inside a backward pass
        return False
    _join_pending = True
    return True


def _colsum_on(t, stream_handle):
    if t.dtype != torch.float32:
        raise TypeError("isg_colsum is fp32-only")
    lib = L.load()
    rows, cols = t.shape
    out = torch.empty(cols, dtype=torch.float32, device=t.device)
    nbytes = lib.isg_colsum_workspace_bytes(rows, cols)
    ws = L.workspace(nbytes, t.device)
    L.call("isg_colsum", L.ptr(t), t.stride(0), rows, cols, L.ptr(out), L.ptr(ws), nbytes, stream_handle)
    return out


def colsum(t):
    if not (_SIDE_STREAM_ENABLED and _side_ok and t.is_cuda and not torch.cuda.is_current_stream_capturing()
            and _defer_join()):
        return _colsum_on(t, L.stream())
    side = _side_stream(t.device)
    side.wait_stream(torch.cuda.current_stream(t.device))
    t.record_stream(side)  # autograd may free `t` on the main stream while the side stream still reads it
    with torch.cuda.stream(side):
        return _colsum_on(t, side.cuda_stream)


def gelu_bwd(gy, z):
    lib = L.load()
    gz = torch.empty_like(gy)
    L.call("isg_gelu_bwd", L.ptr(gy), L.ptr(z), L.ptr(gz), gy.numel(), L.stream())
    return gz


class LinearAct(torch.autograd.Function):
    """y = act(x W^T + b) — replaces PyG `Linear` / torch.nn.Linear (+ torch.nn.GELU) call sites:
    lin_l/lin_r/lin_edge (models/mgat_v2_conv.py:177,181,259), x_proj (models/mgat.py:156),
    node_nn / ques_nn (models/masking.py:137,152)."""

    @staticmethod
    def forward(ctx, x, w, b, act, mode=None, bwd_mode=None):
        L.require_cuda(x, w, b)
        x, w = _c(x), _c(w)
        b = _c(b) if b is not None else None
        mode = _GEMM_MODE if mode is None else mode
        w_t = w_t_lo = None
        if mode == 1 and x.shape[0] >= PRESPLIT_MIN_ROWS and w.shape[0] % 4 == 0:
            w_t, w_t_lo = transpose_split(w)
        y, z = linear_fwd_raw(x, w, b, act, want_pre=(act != L.ACT_NONE), mode=mode, w_t=w_t, w_t_lo=w_t_lo)
        ctx.act = act
        ctx.mode = mode
        # gradients never feed a discrete decision: a projection forced to strict-fp32 FFMA in the forward pass
        # (the sampler's gate) still runs its backward products on the tensor cores in the global fp32-grade mode
        ctx.bwd_mode = _GEMM_MODE if (mode == 0 and bwd_mode is None) else (mode if bwd_mode is None else bwd_mode)
        ctx.has_bias = b is not None
        ctx.save_for_backward(x, w, z)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, z = ctx.saved_tensors
        if ctx.act == L.ACT_GELU:
            gy = gelu_bwd(_c(gy), z)
        elif gy.stride(1) != 1 or gy.stride(0) % 4 != 0:
            gy = gy.contiguous()  # column views of a wider buffer are consumed through their pitch
        gx = None
        if ctx.needs_input_grad[0]:
            w_lo = split_lo(w) if (ctx.bwd_mode == 1 and gy.shape[0] >= PRESPLIT_MIN_ROWS and w.numel() % 4 == 0) else None
            gx = linear_dgrad_raw(gy, w, mode=ctx.bwd_mode, w_lo=w_lo)
        gw = linear_wgrad_raw(gy, x, mode=ctx.bwd_mode) if ctx.needs_input_grad[1] else None
        gb = colsum(gy) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return gx, gw, gb, None, None, None


def linear(x, w, b=None, act=L.ACT_NONE, mode=None, bwd_mode=None):
    """mode=None -> the global projection mode (set_gemm_mode / ISG_GEMM_MODE); bwd_mode=None -> the backward
    products use `mode`, except that a forward forced to mode 0 falls back to the global mode for its gradients."""
    return LinearAct.apply(x, w, b, act, mode, bwd_mode)


# ------------------------------------------------------------------------------------------
# node-side fused ops
# ------------------------------------------------------------------------------------------
class InstrGate(torch.autograd.Function):
    """x = gelu(x * instruction[batch])  (models/mgat_v2_conv.py:156-157)."""

    @staticmethod
    def forward(ctx, x, ins, gi):
        L.require_cuda(x, ins)
        x, ins = _c(x), _c(ins)
        y = torch.empty_like(x)
        L.call("isg_instr_gate_fwd", L.ptr(x), L.ptr(ins), L.ptr(gi.batch32), x.shape[0], x.shape[1],
                                            L.ptr(y), L.stream())
        ctx.gi = gi
        ctx.save_for_backward(x, ins)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, ins = ctx.saved_tensors
        gi = ctx.gi
        gy = _c(gy)
        gx = torch.empty_like(x)
        gins = torch.empty_like(ins)
        L.call("isg_instr_gate_bwd", L.ptr(gy), L.ptr(x), L.ptr(ins), L.ptr(gi.graph_ptr), gi.B, x.shape[1],
                                            None, 0, L.ptr(gx), L.ptr(gins), L.stream())
        return gx, gins, None


class ConcatInstr(torch.autograd.Function):
    """x = cat(x, instruction[batch])  (models/mgat_v2_conv.py:153-154, the `concat_instr` variant)."""

    @staticmethod
    def forward(ctx, x, ins, gi):
        L.require_cuda(x, ins)
        x, ins = _c(x), _c(ins)
        N, D = x.shape
        y = torch.empty(N, 2 * D, dtype=x.dtype, device=x.device)
        L.call("isg_concat_instr_fwd", L.ptr(x), L.ptr(ins), L.ptr(gi.batch32), N, D, L.ptr(y), L.stream())
        ctx.gi, ctx.D, ctx.ins_shape = gi, D, ins.shape
        return y

    @staticmethod
    def backward(ctx, gy):
        gi, D = ctx.gi, ctx.D
        gy = _c(gy)
        gx = torch.empty(gy.shape[0], D, dtype=gy.dtype, device=gy.device)
        gins = torch.empty(ctx.ins_shape, dtype=gy.dtype, device=gy.device)
        L.call("isg_concat_instr_bwd", L.ptr(gy), L.ptr(gi.graph_ptr), gi.B, D, None, 0, L.ptr(gx), L.ptr(gins), L.stream())
        return gx, gins, None


class GateTheta(torch.autograd.Function):
    """theta = gelu(<xn, q[batch[batch]]> / sqrt(D)) [* keep]  (models/masking.py:151-155, double gather via
    models/mgat_v2_conv.py:166-168; keep [N,1] = the dropout keep-mask of masking.py:159, 0 or 1/(1-p), folded
    into the kernel).  q is [B, D] (one row per graph)."""

    @staticmethod
    def forward(ctx, xn, q, gi, double_gather, keep=None):
        xn, q = _c(xn), _c(q)
        keep = _c(keep.to(torch.float32)) if keep is not None else None
        dbl = 1 if double_gather else 0
        if q.shape[0] != (gi.B if dbl else xn.shape[0]):
            raise ValueError("GateTheta: q must be [B,D] with double_gather, [N,D] without")
        if dbl and gi.B > xn.shape[0]:
            # batch[batch[n]] uses a graph id as a node index: the reference raises here too (masking.py:151-153)
            raise IndexError(f"double gather needs num_graphs <= num_nodes, got B={gi.B}, N={xn.shape[0]}")
        theta = torch.empty(xn.shape[0], 1, dtype=torch.float32, device=xn.device)
        L.call("isg_gate_theta_fwd", L.ptr(xn), L.ptr(q), L.ptr(gi.batch32), xn.shape[0], xn.shape[1], dbl,
                                            L.ptr(keep), L.ptr(theta), L.stream())
        ctx.gi, ctx.dbl = gi, dbl
        ctx.save_for_backward(xn, q, keep)
        return theta

    @staticmethod
    def backward(ctx, gth):
        xn, q, keep = ctx.saved_tensors
        gi = ctx.gi
        gth = _c(gth)
        gxn = torch.empty_like(xn)
        gq = torch.empty_like(q)
        scratch = torch.empty(xn.shape[0], dtype=torch.float32, device=xn.device)
        L.call("isg_gate_theta_bwd", L.ptr(gth), L.ptr(xn), L.ptr(q), L.ptr(gi.batch32), L.ptr(gi.graph_ptr),
                                            xn.shape[0], gi.B, xn.shape[1], ctx.dbl, L.ptr(keep), L.ptr(gxn),
                                            L.ptr(gq), L.ptr(scratch), L.stream())
        return gxn, gq, None, None, None


class SdpaGraphNormResidual(torch.autograd.Function):
    """h_out = GraphNorm(scatter_sdpa(ins, v, v, batch), batch) + h_in  (models/mgat.py:168-172;
    utils/scatter_scaled_dot_product.py:6-15; torch_geometric GraphNorm, eps 1e-5)."""

    @staticmethod
    def forward(ctx, v, ins, h_in, weight, bias, mean_scale, gi, eps):
        v, ins, h_in = _c(v), _c(ins), _c(h_in)
        N, D = v.shape
        h_out = torch.empty_like(v)
        a = torch.empty(N, dtype=torch.float32, device=v.device)
        mean = torch.empty(gi.B, D, dtype=torch.float32, device=v.device)
        rstd = torch.empty(gi.B, D, dtype=torch.float32, device=v.device)
        L.call("isg_sdpa_graphnorm_fwd", L.ptr(v), L.ptr(ins), L.ptr(h_in), L.ptr(weight), L.ptr(bias),
                                                L.ptr(mean_scale), L.ptr(gi.graph_ptr), gi.B, D, gi.nmax, eps,
                                                L.ptr(h_out), L.ptr(a), L.ptr(mean), L.ptr(rstd), L.stream())
        ctx.gi = gi
        ctx.save_for_backward(v, ins, weight, mean_scale, a, mean, rstd)
        return h_out

    @staticmethod
    def backward(ctx, g):
        v, ins, weight, mean_scale, a, mean, rstd = ctx.saved_tensors
        gi = ctx.gi
        g = _c(g)
        N, D = v.shape
        gv = torch.empty_like(v)
        gins = torch.empty_like(ins)
        parts = torch.empty(3, gi.B, D, dtype=torch.float32, device=v.device)
        L.call("isg_sdpa_graphnorm_bwd", L.ptr(g), L.ptr(v), L.ptr(ins), L.ptr(weight), L.ptr(mean_scale),
                                                L.ptr(a), L.ptr(mean), L.ptr(rstd), L.ptr(gi.graph_ptr), gi.B, D,
                                                gi.nmax, L.ptr(gv), L.ptr(gins), L.ptr(parts[0]), L.ptr(parts[1]),
                                                L.ptr(parts[2]), None, L.stream())
        gw, gb, gms = colsum(parts[0]), colsum(parts[1]), colsum(parts[2])
        return gv, gins, g, gw, gb, gms, None, None


# ------------------------------------------------------------------------------------------
# edge kernel (b)
# ------------------------------------------------------------------------------------------
class GatEdge(torch.autograd.Function):
    """MaskingGATv2Conv.message + propagate + softmax + aggregate + bias
    (models/mgat_v2_conv.py:215,226-232,243-279).  xlr [N, 2*H*C] = [x_l | x_r], the output of ONE fused
    projection with the stacked weights [W_l; W_r] (both halves share a row pitch, so the kernels read
    them in place and the backward returns one [g_xl | g_xr] tensor that feeds one dgrad / wgrad);
    e_proj [E,H*C], att [1,H,C], bias [H*C], edge_mask [E,1] or None.  Returns (out [N,H*C], alpha [E,H])."""

    @staticmethod
    def forward(ctx, xlr, e_proj, att, bias, edge_mask, gi, heads, slope):
        L.require_cuda(xlr, e_proj, att)
        xlr, e_proj, att = _c(xlr), _c(e_proj), _c(att)
        em = _c(edge_mask) if edge_mask is not None else None
        N, HC2 = xlr.shape
        HC = HC2 // 2
        C = HC // heads
        x_l, x_r = xlr[:, :HC], xlr[:, HC:]
        out = torch.empty(N, HC, dtype=xlr.dtype, device=xlr.device)
        alpha = torch.empty(gi.E, heads, dtype=torch.float32, device=xlr.device)
        L.call("isg_gat_edge_fwd", L.ptr(x_l), L.ptr(x_r), HC2, L.ptr(e_proj), L.ptr(att),
                                          L.ptr(bias), L.ptr(em), L.ptr(gi.dst_ptr), L.ptr(gi.dst_nbr),
                                          L.ptr(gi.dst_eid), L.ptr(gi.dst_order), L.ptr(out), HC, L.ptr(alpha), N, gi.E,
                                          heads, C,
                                          slope, L.dtype_code(xlr), L.stream())
        ctx.gi, ctx.heads, ctx.slope = gi, heads, slope
        ctx.has_bias = bias is not None
        ctx.save_for_backward(xlr, e_proj, att, bias, em, alpha, out)
        ctx.mark_non_differentiable(alpha)
        return out, alpha

    @staticmethod
    def backward(ctx, g_out, _g_alpha):
        xlr, e_proj, att, bias, em, alpha, out = ctx.saved_tensors
        gi, H = ctx.gi, ctx.heads
        lib = L.load()
        g_out = _c(g_out)
        N, HC = out.shape
        C = HC // H
        dev = out.device
        x_l, x_r = xlr[:, :HC], xlr[:, HC:]
        g_xlr = torch.empty(N, 2 * HC, dtype=xlr.dtype, device=dev)  # [g_xl | g_xr], one pitch
        g_xl, g_xr = g_xlr[:, :HC], g_xlr[:, HC:]
        g_ep = torch.empty(gi.E, HC, dtype=e_proj.dtype, device=dev)
        g_att = torch.empty(att.shape, dtype=torch.float32, device=dev)
        g_em = torch.empty(gi.E, 1, dtype=torch.float32, device=dev) if em is not None else None
        nbytes = lib.isg_gat_edge_bwd_workspace_bytes(N, gi.E, gi.B, H, C)
        ws = L.workspace(nbytes, dev)
        fused = edge_bwd_fused(gi)
        L.call("isg_gat_edge_bwd", L.ptr(g_out), g_out.stride(0), L.ptr(x_l), L.ptr(x_r), 2 * HC,
                                     L.ptr(e_proj), L.ptr(att), L.ptr(bias), L.ptr(em), L.ptr(alpha), L.ptr(out), HC,
                                     L.ptr(gi.dst_ptr), L.ptr(gi.dst_nbr), L.ptr(gi.dst_eid), L.ptr(gi.dst_order),
                                     L.ptr(gi.src_ptr), L.ptr(gi.src_nbr), L.ptr(gi.src_eid), L.ptr(gi.src_order),
                                     L.ptr(g_xl), L.ptr(g_xr), 2 * HC,
                                     L.ptr(g_ep), L.ptr(g_att), L.ptr(g_em), N, gi.E, H, C, ctx.slope,
                                     L.dtype_code(xlr), L.ptr(gi.batch32) if fused else None,
                                     L.ptr(gi.graph_ptr) if fused else None, gi.B, gi.nmax if fused else 0,
                                     L.ptr(ws), nbytes, L.stream())
        # the column-sum kernel is fp32-only: bf16 storage converts g_out once (N x HC, small next to the edge pass)
        g_bias = colsum(g_out if g_out.dtype == torch.float32 else g_out.float()) if ctx.has_bias else None
        if _DEBUG_EDGE_BWD is not None:
            _DEBUG_EDGE_BWD.append(dict(g_out=g_out.clone(), x_l=x_l.clone(), x_r=x_r.clone(), e_proj=e_proj.clone(),
                                        att=att.clone(), bias=bias.clone() if bias is not None else None,
                                        em=em.clone() if em is not None else None, alpha=alpha.clone(),
                                        out=out.clone(), g_xl=g_xl.clone(), g_xr=g_xr.clone(), g_ep=g_ep.clone(),
                                        g_att=g_att.clone()))
        return g_xlr, g_ep, g_att, g_bias, g_em, None, None, None


class NodeMaskToEdgeMaskFn(torch.autograd.Function):
    """sampling/node_edge_masks.py:5-19 — including its custom (dst-only) backward."""

    @staticmethod
    def forward(ctx, mask, gi):
        mask = _c(mask.to(torch.float32))
        em = torch.empty(gi.E, 1, dtype=torch.float32, device=mask.device)
        L.call("isg_node_edge_mask_fwd", L.ptr(mask), L.ptr(gi.edge_index), gi.E, L.ptr(em), L.stream())
        ctx.gi = gi
        ctx.shape = mask.shape
        return em

    @staticmethod
    def backward(ctx, g_em):
        gi = ctx.gi
        g_em = _c(g_em)
        g_m = torch.empty(ctx.shape, dtype=torch.float32, device=g_em.device)
        L.call("isg_node_edge_mask_bwd", L.ptr(g_em), L.ptr(gi.dst_ptr), L.ptr(gi.dst_eid), gi.N, L.ptr(g_m),
                                                L.stream())
        return g_m, None


# ------------------------------------------------------------------------------------------
# samplers (kernel c).  theta is ragged [N,1]; noise is dense [B,Nmax] (or [B,1,Nmax,1]).
# ------------------------------------------------------------------------------------------
def _noise2d(noise, gi):
    if noise is None:
        return None
    n = noise.reshape(gi.B, -1)
    if n.shape[1] != gi.nmax:
        raise ValueError(f"noise has {n.shape[1]} slots per graph, expected Nmax = {gi.nmax}")
    return _c(n.to(torch.float32))


class TopkImle(torch.autograd.Function):
    """IMLE perturb-and-MAP top-k mask (sampling/methods/wrapper.py:75-172, target.py:44-48,
    imle_scheme.py:16-29, deterministic_scheme.py:36-43), nb_samples = 1.  Returns the ragged
    mask [N,1] (== `output[0].squeeze(0)[valid]`, models/masking.py:169-173)."""

    @staticmethod
    def forward(ctx, theta, noise, gi, k, alpha, beta, tau_in, tau_tgt):
        theta = _c(theta)
        noise = _noise2d(noise, gi)
        N = theta.shape[0]
        mask = torch.empty(N, 1, dtype=torch.float32, device=theta.device)
        zd = torch.empty(gi.B, gi.nmax, dtype=torch.float32, device=theta.device)
        L.call("isg_topk_mask_fwd", L.ptr(theta), L.ptr(noise), L.ptr(gi.graph_ptr), gi.B, gi.nmax, k,
                                           tau_in, L.ptr(mask), L.ptr(zd), L.stream())
        ctx.gi, ctx.cfg = gi, (k, alpha, beta, tau_tgt)
        ctx.save_for_backward(theta, noise, zd)
        return mask

    @staticmethod
    def backward(ctx, dy):
        theta, noise, zd = ctx.saved_tensors
        gi = ctx.gi
        k, alpha, beta, tau_tgt = ctx.cfg
        dy = _c(dy)
        g = torch.empty_like(theta)
        L.call("isg_imle_bwd", L.ptr(dy), L.ptr(theta), L.ptr(noise), L.ptr(zd), L.ptr(gi.graph_ptr), gi.B,
                                      gi.nmax, k, alpha, beta, tau_tgt, L.ptr(g), L.stream())
        return g, None, None, None, None, None, None, None


class TopkAimle(torch.autograd.Function):
    """AIMLE (sampling/methods/aimle.py:83-243 + target_aimle.py:87-162), symmetric perturbation,
    adaptive beta kept in a device-resident state vector (no host sync)."""

    @staticmethod
    def forward(ctx, theta, noise, gi, k, state, adaptive, tau_in, tau_tgt):
        theta = _c(theta)
        noise = _noise2d(noise, gi)
        N = theta.shape[0]
        mask = torch.empty(N, 1, dtype=torch.float32, device=theta.device)
        zd = torch.empty(gi.B, gi.nmax, dtype=torch.float32, device=theta.device)
        L.call("isg_topk_mask_fwd", L.ptr(theta), L.ptr(noise), L.ptr(gi.graph_ptr), gi.B, gi.nmax, k,
                                           tau_in, L.ptr(mask), L.ptr(zd), L.stream())
        ctx.gi, ctx.cfg, ctx.state = gi, (k, adaptive, tau_tgt), state
        ctx.save_for_backward(theta, noise)
        return mask

    @staticmethod
    def backward(ctx, dy):
        theta, noise = ctx.saved_tensors
        gi, state = ctx.gi, ctx.state
        k, adaptive, tau_tgt = ctx.cfg
        lib = L.load()
        dy = _c(dy)
        g = torch.empty_like(theta)
        nbytes = lib.isg_aimle_workspace_bytes()
        ws = L.workspace(nbytes, theta.device)
        L.call("isg_aimle_bwd", L.ptr(dy), L.ptr(theta), L.ptr(noise), L.ptr(gi.graph_ptr), theta.shape[0], gi.B,
                                  gi.nmax, k, tau_tgt, 1 if adaptive else 0, L.ptr(state), L.ptr(g), L.ptr(ws),
                                  nbytes, L.stream())
        return g, None, None, None, None, None, None, None


class GumbelTopk(torch.autograd.Function):
    """Relaxed Gumbel top-k with straight-through (sampling/methods/gumbel_scheme.py:26-107)."""

    @staticmethod
    def forward(ctx, theta, gumbel, gi, k, tau):
        theta = _c(theta)
        gumbel = _noise2d(gumbel, gi)
        N = theta.shape[0]
        k = max(1, min(int(k), gi.nmax))  # local_k = min(k, Nmax), gumbel_scheme.py:54
        mask = torch.empty(N, 1, dtype=torch.float32, device=theta.device)
        saved = torch.empty(gi.B, k, gi.nmax, dtype=torch.float32, device=theta.device)
        L.call("isg_gumbel_topk_fwd", L.ptr(theta), L.ptr(gumbel), L.ptr(gi.graph_ptr), gi.B, gi.nmax, k, tau,
                                             L.ptr(mask), L.ptr(saved), L.stream())
        ctx.gi, ctx.cfg = gi, (k, tau)
        ctx.save_for_backward(saved)
        return mask

    @staticmethod
    def backward(ctx, dy):
        (saved,) = ctx.saved_tensors
        gi = ctx.gi
        k, tau = ctx.cfg
        dy = _c(dy)
        g = torch.empty(dy.shape, dtype=torch.float32, device=dy.device)
        L.call("isg_gumbel_topk_bwd", L.ptr(dy), L.ptr(saved), L.ptr(gi.graph_ptr), gi.B, gi.nmax, k, tau,
                                             L.ptr(g), L.stream())
        return g, None, None, None, None


class SimpleTopk(torch.autograd.Function):
    """SIMPLE exact k-subset marginals + Gumbel top-k sample + straight-through
    (sampling/methods/simple_scheme.py:44-162, simple.py:113-252).  theta [N,1] ragged, gumbel
    [B, n_pad] Gumbel(0,1).  Returns (mask [N,1], marginals [B,Nmax])."""

    @staticmethod
    def forward(ctx, theta, gumbel, gi, k):
        theta = _c(theta)
        lib = L.load()
        npad = lib.isg_simple_npad(gi.nmax)
        g = gumbel.reshape(gi.B, -1)
        if g.shape[1] != npad:
            raise ValueError(f"SIMPLE noise has {g.shape[1]} slots per graph, expected n_pad = {npad}")
        g = _c(g.to(torch.float32))
        k = max(1, min(int(k), gi.nmax))  # local_k = min(k, Nmax), simple_scheme.py:81
        mask = torch.empty(theta.shape[0], 1, dtype=torch.float32, device=theta.device)
        marg = torch.empty(gi.B, gi.nmax, dtype=torch.float32, device=theta.device)
        ctx.gi, ctx.k = gi, k
        ctx.save_for_backward(theta)
        if npad == 1:
            # one slot per graph and k = 1: the only 1-subset is certain (marginal 1, sample 1, zero gradient)
            mask.fill_(1.0)
            marg.fill_(1.0)
            return mask, marg
        L.call("isg_simple_marginals_fwd", L.ptr(theta), L.ptr(g), L.ptr(gi.graph_ptr), gi.B, gi.nmax, k,
               L.ptr(mask), L.ptr(marg), L.stream())
        return mask, marg

    @staticmethod
    def backward(ctx, dy, dmarg):
        (theta,) = ctx.saved_tensors
        gi = ctx.gi
        dy = _c(dy)
        dmarg = _c(dmarg) if dmarg is not None else None
        if gi.nmax <= 1:
            return torch.zeros_like(theta), None, None, None
        g = torch.empty_like(theta)
        L.call("isg_simple_marginals_bwd", L.ptr(dy), L.ptr(dmarg), L.ptr(theta), L.ptr(gi.graph_ptr), gi.B,
               gi.nmax, ctx.k, L.ptr(g), L.stream())
        return g, None, None, None


class AttnPool(torch.autograd.Function):
    """Masked per-graph attention pooling — GlobalAttention.forward after its MLPs (models/att_pooling.py:64-77):
    gate = pyg_softmax(<x*mask, q[batch]>/sqrt(D), batch); out = scatter_add(gate * x*mask).  x [N,D],
    node_mask [N,1] or None, q [B,D].  Returns (out [B,D], gate [N,1])."""

    @staticmethod
    def forward(ctx, x, node_mask, q, gi):
        L.require_cuda(x, q)
        x, q = _c(x), _c(q)
        m = _c(node_mask.to(torch.float32)) if node_mask is not None else None
        N, D = x.shape
        out = torch.empty(gi.B, D, dtype=torch.float32, device=x.device)
        gate = torch.empty(N, 1, dtype=torch.float32, device=x.device)
        L.call("isg_attn_pool_fwd", L.ptr(x), L.ptr(m), L.ptr(q), L.ptr(gi.graph_ptr), gi.B, D, gi.nmax, L.ptr(out),
               L.ptr(gate), L.stream())
        ctx.gi = gi
        ctx.has_mask = m is not None
        ctx.save_for_backward(x, m, q, gate)
        return out, gate

    @staticmethod
    def backward(ctx, g_out, g_gate):
        x, m, q, gate = ctx.saved_tensors
        gi = ctx.gi
        g_out = _c(g_out)
        g_gate = _c(g_gate) if g_gate is not None else None
        g_x = torch.empty_like(x)
        g_m = torch.empty(x.shape[0], 1, dtype=torch.float32, device=x.device) if ctx.has_mask else None
        g_q = torch.empty_like(q)
        L.call("isg_attn_pool_bwd", L.ptr(g_out), L.ptr(g_gate), L.ptr(x), L.ptr(m), L.ptr(q), L.ptr(gate),
               L.ptr(gi.graph_ptr), gi.B, x.shape[1], gi.nmax, L.ptr(g_x), L.ptr(g_m), L.ptr(g_q), L.stream())
        return g_x, g_m, g_q, None


# ------------------------------------------------------------------------------------------
# scene-graph encoding layer (csrc/sgenc.cu)
# ------------------------------------------------------------------------------------------
class GatherAddAct(torch.autograd.Function):
    """y = act(a[src] + b[dst] + q): the first Linear of EdgeModel.edge_mlp / NodeModel.node_mlp_1 applied to the
    concatenation [x[src], x[dst], e] / [x[src], e'] (models/scene_graph_encoder.py:118-120,138-140) with its
    weight split by column block — a, b [N,D] are the per-NODE projections, q [E,D] the per-edge one (+ bias)."""

    @staticmethod
    def forward(ctx, a, b, q, gi, act):
        L.require_cuda(q)
        a = _c(a) if a is not None else None
        b = _c(b) if b is not None else None
        q = _c(q)
        E, D = q.shape
        y = torch.empty_like(q)
        z = torch.empty_like(q) if act != L.ACT_NONE else None
        L.call("isg_gather_add_act_fwd", L.ptr(a), L.ptr(b), L.ptr(q), L.ptr(gi.edge_index), E, D, act, L.ptr(z),
               L.ptr(y), L.stream())
        ctx.gi, ctx.act, ctx.has = gi, act, (a is not None, b is not None)
        ctx.save_for_backward(z)
        return y

    @staticmethod
    def backward(ctx, gy):
        (z,) = ctx.saved_tensors
        gi = ctx.gi
        gz = gelu_bwd(_c(gy), z) if ctx.act == L.ACT_GELU else _c(gy)
        D = gz.shape[1]
        ga = gb = None
        if ctx.has[0]:  # a was gathered by source: transpose = segment sum over the src-sorted CSR
            ga = torch.empty(gi.N, D, dtype=torch.float32, device=gz.device)
            L.call("isg_segment_sum", L.ptr(gz), L.ptr(gi.src_ptr), L.ptr(gi.src_eid), gi.N, D, 0, L.ptr(ga), L.stream())
        if ctx.has[1]:
            gb = torch.empty(gi.N, D, dtype=torch.float32, device=gz.device)
            L.call("isg_segment_sum", L.ptr(gz), L.ptr(gi.dst_ptr), L.ptr(gi.dst_eid), gi.N, D, 0, L.ptr(gb), L.stream())
        return ga, gb, gz, None, None


class SegmentMeanByDst(torch.autograd.Function):
    """torch_scatter.scatter_mean(m, col, dim=0, dim_size=N) (models/scene_graph_encoder.py:141) over the
    dst-sorted CSR: deterministic, no atomics."""

    @staticmethod
    def forward(ctx, m, gi):
        m = _c(m)
        out = torch.empty(gi.N, m.shape[1], dtype=torch.float32, device=m.device)
        L.call("isg_segment_sum", L.ptr(m), L.ptr(gi.dst_ptr), L.ptr(gi.dst_eid), gi.N, m.shape[1], 1, L.ptr(out),
               L.stream())
        ctx.gi = gi
        return out

    @staticmethod
    def backward(ctx, g):
        gi = ctx.gi
        g = _c(g)
        gm = torch.empty(gi.E, g.shape[1], dtype=torch.float32, device=g.device)
        L.call("isg_gather_rows", L.ptr(g), gi.edge_index.data_ptr() + 8 * gi.E, L.ptr(gi.dst_ptr), gi.E, g.shape[1],
               L.ptr(gm), L.stream())
        return gm, None


class GraphNorm64(torch.autograd.Function):
    """GraphNorm evaluated in float64 on the device (models/scene_graph_encoder.py:99-102 does it through a CPU
    DoubleTensor round trip every forward)."""

    @staticmethod
    def forward(ctx, x, weight, bias, mean_scale, gi, eps):
        x = _c(x)
        N, D = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(gi.B, D, dtype=torch.float64, device=x.device)
        rstd = torch.empty(gi.B, D, dtype=torch.float64, device=x.device)
        L.call("isg_graphnorm64_fwd", L.ptr(x), L.ptr(weight), L.ptr(bias), L.ptr(mean_scale), L.ptr(gi.graph_ptr),
               gi.B, D, float(eps), L.ptr(y), L.ptr(mean), L.ptr(rstd), L.stream())
        ctx.gi = gi
        ctx.save_for_backward(x, weight, mean_scale, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight, mean_scale, mean, rstd = ctx.saved_tensors
        gi = ctx.gi
        g = _c(g)
        D = x.shape[1]
        gx = torch.empty_like(x)
        parts = torch.empty(3, gi.B, D, dtype=torch.float32, device=x.device)
        L.call("isg_graphnorm64_bwd", L.ptr(g), L.ptr(x), L.ptr(weight), L.ptr(mean_scale), L.ptr(mean), L.ptr(rstd),
               L.ptr(gi.graph_ptr), gi.B, D, L.ptr(gx), L.ptr(parts[0]), L.ptr(parts[1]), L.ptr(parts[2]), L.stream())
        return gx, colsum(parts[0]), colsum(parts[1]), colsum(parts[2]), None, None
