"""Host side of the layer executor (csrc/executor.cu): MGAT.forward as ONE autograd Function whose forward
issues one C call per layer and whose backward issues one per layer in reverse — instead of ~155 per-operator
C-ABI calls, ~220 tensor allocations and 33 autograd Functions per training step (the per-operator path in
ops.py stays as the API of the individual modules and as the cross-check of this one).

Reference being mirrored: the loop of MGAT.forward (models/mgat.py:131-177).

Memory: the activations of all layers live in one arena allocated per forward; the backward's temporaries in one
workspace; all parameter gradients in one flat fp32 buffer laid out in backward order (layer L-1 first), whose
slices are returned as the parameters' gradients (autograd adopts them without a copy when .grad is None).  A
data-parallel reducer can (a) supply that buffer (`model._isg_grad_bucket`), so the all-reduce runs in place on
what .grad points at, and (b) be notified after each layer's backward (`model._isg_after_layer_backward`) to start
that layer's all-reduce while the earlier layers are still computing (isg_b200.dp.LayerOverlappedAllReduce)."""
import ctypes
import os

import numpy as np
import torch

from .. import lib as L
from .. import ops

SAMPLER_CODE = {"imle": 1, "aimle": 2, "gumbel": 3, "simple": 4}
_slots = None


class _Slots:
    def __init__(self):
        lib = L.load()
        self.n_dims, self.n_scalars, self.n_ptrs = (lib.isg_layer_slot_count(i) for i in range(3))

    def __getattr__(self, name):
        idx = L.load().isg_layer_slot(name.encode())
        if idx < 0:
            raise AttributeError(f"libisg.so has no layer slot {name!r}")
        setattr(self, name, idx)
        return idx


def slots():
    global _slots
    if _slots is None:
        _slots = _Slots()
    return _slots


# Weight gradients on a side stream (csrc/executor.cu: fork / join inside isg_mgat_layer_bwd): one stream and two
# events per device, created on first use.  ISG_SIDE_WGRAD=0 keeps everything on the caller's stream.
_SIDE_WGRAD = os.environ.get("ISG_SIDE_WGRAD", "1") != "0"
_side = {}


def _side_handles(device):
    h = _side.get(device)
    if h is None:
        st = torch.cuda.Stream(device=device)
        evs = [torch.cuda.Event() for _ in range(4 + _MAX_EPROJ_EVENTS)]
        if torch.cuda.is_current_stream_capturing():
            return None  # the events must exist before a capture starts (first use happens in the warm-up steps)
        for ev in evs:
            ev.record()  # materialises the cudaEvent_t
        h = _side[device] = (st, evs[:2], evs[4:], evs[2:4])
    return h


# Forward: the lin_edge products of ALL layers depend on edge_attr only, so they are issued on the side stream at the
# start of the forward pass and run beside the main stream's chain — gating, the gate projections and the sampler of
# a masked layer, lin_l|lin_r, the x_proj pair, scatter-SDPA + GraphNorm — most of which are launches that fill a
# fraction of the 148 SMs (one CTA per graph, or a last wave of a few tiles).  Layer i's edge kernel waits for event i.
# ISG_SIDE_EPROJ=0 keeps the products inside the layer call on the caller's stream.
_SIDE_EPROJ = os.environ.get("ISG_SIDE_EPROJ", "1") != "0"
_MAX_EPROJ_EVENTS = 8


def set_side_eproj(on):
    global _SIDE_EPROJ
    _SIDE_EPROJ = bool(on)


# Backward: the join with the side stream at the end of every layer made the main stream wait for the two products
# forked last (the E-sized lin_edge weight gradient among them), which then ran beside nothing.  With two alternating
# workspaces layer i's weight gradients may still run while layer i-1's chain is under way; layer i-2 waits for them
# (it re-uses the workspace) and the main stream joins once after the last layer.  Off when a per-layer hook reads the
# gradients during the backward pass (dp.LayerGradAllReduce).  ISG_DEFER_JOIN=0 restores the per-layer join.
_DEFER_JOIN = os.environ.get("ISG_DEFER_JOIN", "1") != "0"


def set_defer_join(on):
    global _DEFER_JOIN
    _DEFER_JOIN = bool(on)


def _al(nbytes):
    return (int(nbytes) + 255) & ~255


# parameters of one layer in the order their gradients are laid out; (attribute path, stacked-with-previous)
def layer_params(model, i):
    """(name, Parameter) pairs of layer i in gradient-layout order.  Cached per model: walking nn.Module attributes
    costs ~1 us each and this list is needed a dozen times per step."""
    cache = model.__dict__.get("_isg_layer_params")
    if cache is None:
        cache = model.__dict__["_isg_layer_params"] = {}
    hit = cache.get(i)
    if hit is not None and hit[0] == model.convs[i].mask.masking_threshold:
        return hit[1]
    ps = _layer_params(model, i)
    cache[i] = (model.convs[i].mask.masking_threshold, ps)
    return ps


def _layer_params(model, i):
    conv, xp, bn = model.convs[i], model.x_proj[i], model.bns[i]
    ps = [("W_L", conv.lin_l.weight), ("W_R", conv.lin_r.weight), ("B_L", conv.lin_l.bias), ("B_R", conv.lin_r.bias),
          ("W_E", conv.lin_edge.weight), ("ATT", conv.att), ("BIAS", conv.bias),
          ("WP0", xp[0].weight), ("BP0", xp[0].bias), ("WP2", xp[2].weight), ("BP2", xp[2].bias),
          ("BN_W", bn.weight), ("BN_B", bn.bias), ("BN_MS", bn.mean_scale)]
    if conv.mask.masking_threshold != 1.0:
        m = conv.mask
        ps += [("WN", m.node_nn[0].weight), ("BNN", m.node_nn[0].bias), ("WQ", m.ques_nn[0].weight),
               ("BQ", m.ques_nn[0].bias)]
    return ps


def ensure_stacked(conv):
    """lin_l / lin_r weights (and biases) adjacent in memory, so [W_l; W_r] is ONE [2HC, D] operand without a
    per-step torch.cat: the parameters' storage is re-pointed into a shared buffer once (and again whenever
    .to() / load-time re-allocation separated them).  state_dict keys, shapes and optimizer state are unaffected."""
    wl, wr, bl, br = conv.lin_l.weight, conv.lin_r.weight, conv.lin_l.bias, conv.lin_r.bias
    if wr.data_ptr() != wl.data_ptr() + wl.numel() * 4 or not (wl.is_contiguous() and wr.is_contiguous()):
        buf = torch.empty(2 * wl.shape[0], wl.shape[1], dtype=wl.dtype, device=wl.device)
        buf[: wl.shape[0]].copy_(wl.data)
        buf[wl.shape[0]:].copy_(wr.data)
        wl.data, wr.data = buf[: wl.shape[0]], buf[wl.shape[0]:]
    if br.data_ptr() != bl.data_ptr() + bl.numel() * 4:
        buf = torch.empty(2 * bl.shape[0], dtype=bl.dtype, device=bl.device)
        buf[: bl.shape[0]].copy_(bl.data)
        buf[bl.shape[0]:].copy_(br.data)
        bl.data, br.data = buf[: bl.shape[0]], buf[bl.shape[0]:]


def supported(model, explainer, params=None):
    """The executor covers the configuration ISubGVQA builds (models/isubgvqa.py:159-176); anything else runs the
    per-operator path.  `params`: the caller's cached parameter list (nn.Module.parameters() walks the module tree,
    ~0.4 ms per call — a fifth of the host time of a step)."""
    if explainer or model.use_global_mask or getattr(model, "debug_tensors", None) is not None:
        return False
    for conv in model.convs:
        m = conv.mask
        masked = m.masking_threshold != 1.0
        if getattr(conv, "debug_tensors", None) is not None or conv.share_weights or not conv.use_instr:
            return False
        if conv.bias is None or conv.lin_l.bias is None or conv.lin_r.bias is None:
            return False
        if masked and (model.interpretable_mode or not m.use_topk or m.sampler_type not in SAMPLER_CODE):
            return False
        if conv.heads * conv.out_channels != conv.lin_l.weight.shape[0] or conv.in_channels != conv.out_channels:
            return False
    return all(p.dtype == torch.float32 for p in (params if params is not None else model.parameters()))


BF16_MODE = 3  # ops.set_gemm_mode(3): the bf16 configuration (projections on kind::f16, bf16 activation storage)


def pad8(n):
    return (int(n) + 7) & ~7


# the four big projection weights of a layer, as (slot stem, parameter getter)
def _bf16_weights(model, i):
    conv, xp = model.convs[i], model.x_proj[i]
    wl, wr = conv.lin_l.weight, conv.lin_r.weight
    w_lr = wl.data.new_empty(0).set_(wl.data.untyped_storage(), wl.data.storage_offset(),
                                     (wl.shape[0] + wr.shape[0], wl.shape[1]))  # the stacked [W_l; W_r] (ensure_stacked)
    return [("P_W_LR", w_lr), ("P_W_E", conv.lin_edge.weight.data), ("P_WP0", xp[0].weight.data),
            ("P_WP2", xp[2].weight.data)]


def convert_weights_bf16(model):
    """bf16 copies W [Nout, pad8(K)] and W^T [K, pad8(Nout)] of the four projection weights of every layer, made by
    ONE isg_weights_to_bf16 launch per forward (weights change every optimizer step).  Returns (buffer, {(layer,
    slot): data_ptr})."""
    jobs = [(i, stem, w) for i in range(len(model.convs)) for stem, w in _bf16_weights(model, i)]
    total = sum(w.shape[0] * pad8(w.shape[1]) + w.shape[1] * pad8(w.shape[0]) + 16 for _i, _s, w in jobs)
    buf = torch.empty(total, dtype=torch.bfloat16, device=jobs[0][2].device)
    n = len(jobs)
    base = buf.data_ptr()
    ptrs, off = {}, 0
    w_in, w_out, t_out = (ctypes.c_void_p * n)(), (ctypes.c_void_p * n)(), (ctypes.c_void_p * n)()
    rows, cols, ldw, ldt = ((ctypes.c_int * n)() for _ in range(4))
    for j, (i, stem, w) in enumerate(jobs):
        nout, k = w.shape
        w_in[j], rows[j], cols[j], ldw[j], ldt[j] = w.data_ptr(), nout, k, pad8(k), pad8(nout)
        w_out[j] = base + 2 * off
        off += (nout * pad8(k) + 7) & ~7
        t_out[j] = base + 2 * off
        off += (k * pad8(nout) + 7) & ~7
        ptrs[(i, stem + "_BF")], ptrs[(i, stem + "_T_BF")] = w_out[j], t_out[j]
    L.call("isg_weights_to_bf16", n, w_in, rows, cols, w_out, ldw, t_out, ldt, L.stream())
    return buf, ptrs


class _Plan:
    """Per-(model, sizes) constants: arena offsets and the flat layout of the parameter gradients."""

    def __init__(self, model, N, E, B, nmax, bf16=False):
        conv0 = model.convs[0]
        self.L = len(model.convs)
        self.D, self.H = conv0.out_channels, conv0.heads
        self.HC = self.D * self.H
        self.HID = model.x_proj[0][0].weight.shape[0]
        D, H, HC, HID = self.D, self.H, self.HC, self.HID
        self.masked = [c.mask.masking_threshold != 1.0 for c in model.convs]
        off = 0
        self.bf16 = bool(bf16)
        half = {"P_XLR", "P_EPROJ", "P_OUT", "P_Z1", "P_Y1", "P_XG_BF"} if bf16 else set()  # stored as bf16
        self.act = []  # per layer: slot name -> byte offset in the arena
        for i in range(self.L):
            sizes = [("P_XG", N * D), ("P_XLR", N * 2 * HC), ("P_EPROJ", E * HC), ("P_OUT", N * HC), ("P_ALPHA", E * H),
                     ("P_Z1", N * HID), ("P_Y1", N * HID), ("P_Z2", N * D), ("P_Y2", N * D), ("P_SA", N),
                     ("P_MEAN", B * D), ("P_RSTD", B * D), ("P_H_OUT", N * D)]
            if bf16:
                sizes.append(("P_XG_BF", N * pad8(D)))
            if self.masked[i]:
                k = int(model.convs[i].mask.sample_k)
                sizes += [("P_XN_PRE", N * D), ("P_XN", N * D), ("P_Q_PRE", B * D), ("P_Q", B * D), ("P_THETA", N),
                          ("P_MASK", N), ("P_ZD", B * max(k, 1) * max(nmax, 1)), ("P_MARG", B * max(nmax, 1)),
                          ("P_EMASK", E)]
            o = {}
            for name, n in sizes:
                o[name] = off
                off += _al((2 if name in half else 4) * max(n, 1))
            self.act.append(o)
        self.arena_bytes = off
        # parameter gradients: backward order (last layer first) so each layer is one contiguous slice
        self.grad_off, self.layer_span, goff = {}, {}, 0
        for i in reversed(range(self.L)):
            start = goff
            for name, p in layer_params(model, i):
                self.grad_off[(i, name)] = goff
                goff += p.numel()
            self.layer_span[i] = (start, goff)
        self.grad_numel = goff


def _plan(model, N, E, B, nmax, bf16=False):
    cache = model.__dict__.setdefault("_isg_plans", {})
    key = (N, E, B, nmax, bool(bf16))
    pl = cache.get(key)
    if pl is None:
        if len(cache) > 64:
            cache.clear()
        pl = cache[key] = _Plan(model, N, E, B, nmax, bf16)
    return pl


def _arrays():
    s = slots()
    return (np.zeros(s.n_dims, dtype=np.int64), np.zeros(s.n_scalars, dtype=np.float64),
            np.zeros(s.n_ptrs, dtype=np.uint64))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _fill_common(model, pl, gi, i, d, f, p, x_in_ptr, ins_ptr, glf, edge_attr, spec, arena_ptr, gemm_mode, bf=None):
    s = slots()
    conv = model.convs[i]
    if pl.bf16:
        ea_bf, wptrs = bf
        d[s.D_BF16] = 1
        p[s.P_EDGE_ATTR_BF] = ea_bf.data_ptr()
        for (layer, slot), ptr in wptrs.items():
            if layer == i:
                p[getattr(s, slot)] = ptr
        gemm_mode = 1  # the small fp32 products that remain (gate projections' backward) run fp32-grade
    d[s.D_N], d[s.D_E], d[s.D_B], d[s.D_D], d[s.D_H], d[s.D_HID] = gi.N, gi.E, gi.B, pl.D, pl.H, pl.HID
    d[s.D_NMAX], d[s.D_MASKED] = gi.nmax, 1 if pl.masked[i] else 0
    d[s.D_GEMM_MODE], d[s.D_GATE_MODE] = gemm_mode, conv.mask.GATE_GEMM_MODE
    d[s.D_CLOSED] = 1 if gi.closed else 0
    d[s.D_EDGE_FUSED] = 1 if ops.edge_bwd_fused(gi) else 0
    f[s.F_SLOPE], f[s.F_EPS] = float(conv.negative_slope), float(model.bns[i].eps)
    for name, t in (("P_DST_PTR", gi.dst_ptr), ("P_DST_NBR", gi.dst_nbr), ("P_DST_EID", gi.dst_eid),
                    ("P_SRC_PTR", gi.src_ptr), ("P_SRC_NBR", gi.src_nbr), ("P_SRC_EID", gi.src_eid),
                    ("P_GRAPH_PTR", gi.graph_ptr), ("P_BATCH32", gi.batch32), ("P_EDGE_INDEX", gi.edge_index),
                    ("P_DST_ORDER", gi.dst_order), ("P_SRC_ORDER", gi.src_order),
                    ("P_GLF", glf), ("P_EDGE_ATTR", edge_attr)):
        p[getattr(s, name)] = t.data_ptr() if t is not None else 0
    p[s.P_X_IN], p[s.P_INS] = x_in_ptr, ins_ptr
    for name, t in layer_params(model, i):
        if name in ("W_R", "B_R"):
            continue
        slot = {"W_L": "P_W_LR", "B_L": "P_B_LR"}.get(name, "P_" + name)
        p[getattr(s, slot)] = t.data_ptr()
    for name, off in pl.act[i].items():
        p[getattr(s, name)] = arena_ptr + off
    if pl.masked[i]:
        d[s.D_SAMPLER], d[s.D_K] = spec["code"], spec["k"]
        d[s.D_AIMLE_ADAPTIVE] = 1 if spec.get("adaptive") else 0
        f[s.F_ALPHA], f[s.F_BETA] = spec.get("alpha", 1.0), spec.get("beta", 1.0)
        f[s.F_TAU_IN], f[s.F_TAU_TGT], f[s.F_GUMBEL_TAU] = spec.get("tau_in", 1.0), spec.get("tau_tgt", 1.0), \
            spec.get("gumbel_tau", 0.1)
        p[s.P_NOISE] = spec["noise"].data_ptr() if spec["noise"] is not None else 0
        p[s.P_KEEP] = spec["keep"].data_ptr() if spec["keep"] is not None else 0
        p[s.P_AIMLE_STATE] = spec["state"].data_ptr() if spec.get("state") is not None else 0


def kernel_launches(spec, gi, backward, bf16=False):
    """Kernels one isg_mgat_layer_fwd / _bwd call launches (memsets excluded) — mirrors csrc/executor.cu; used
    for bench.py's gpu_launches claim."""
    K = L.KERNELS_PER_CALL
    masked = spec is not None
    if not backward:
        n = K["isg_instr_gate_fwd"] + 4 * K["isg_linear_fwd"] + K["isg_gat_edge_fwd"] + K["isg_sdpa_graphnorm_fwd"]
        if masked:
            n += 2 * K["isg_linear_fwd"]
            if spec["code"] in (1, 2) and gi.closed:
                n += K["isg_sampler_fused_fwd"]
            else:
                n += K["isg_gate_theta_fwd"] + 1 + K["isg_node_edge_mask_fwd"]
        return n + (1 if bf16 else 0)  # + isg_to_bf16(xg)
    n = (K["isg_sdpa_graphnorm_bwd"] + K["isg_colsum_multi"] + 4 * K["isg_linear_dgrad"] +
         4 * K["isg_linear_wgrad"] + K["isg_gat_edge_bwd"] + K["isg_instr_gate_bwd"])
    if masked:
        n += K["isg_node_edge_mask_bwd"] + (3 if spec["code"] == 2 else 1) + K["isg_gate_theta_bwd"] + \
            2 * (K["isg_gelu_bwd"] + K["isg_linear_dgrad"] + K["isg_linear_wgrad"])
    return n + (1 if bf16 else 0)  # + isg_to_bf16(g_y2)


class MgatFunction(torch.autograd.Function):
    """(x, edge_attr, instr_vectors, global_language_feats, *parameters) -> (h, mask of the last layer or None)."""

    @staticmethod
    def forward(ctx, model, gi, specs, gemm_mode, x, edge_attr, iv, glf, *params):
        ctx.set_materialize_grads(False)
        N, E, B = gi.N, gi.E, gi.B
        bf16 = gemm_mode == BF16_MODE
        pl = _plan(model, N, E, B, gi.nmax, bf16)
        x, edge_attr, iv, glf = (t.contiguous() for t in (x, edge_attr, iv, glf))
        arena = torch.empty(pl.arena_bytes, dtype=torch.uint8, device=x.device)
        ap = arena.data_ptr()
        st = L.stream()
        s = slots()
        bf = None
        if bf16:  # per step: edge_attr once (shared by the 4 layers, forward and wgrad), all projection weights once
            Dp = pad8(pl.D)
            ea_bf = torch.empty(max(E, 1), Dp, dtype=torch.bfloat16, device=x.device)
            L.call("isg_to_bf16", edge_attr.data_ptr(), pl.D, E, pl.D, ea_bf.data_ptr(), Dp, st)
            wbuf, wptrs = convert_weights_bf16(model)
            bf = (ea_bf, wptrs)
            ctx.bf_keep = (ea_bf, wbuf)
        ctx.bf = bf
        x_in_ptr = x.data_ptr()
        side = _side_handles(x.device) if (_SIDE_EPROJ and E > 0 and pl.L <= _MAX_EPROJ_EVENTS) else None
        if side is not None:
            side_st, (ev_fork, _ev_join), ev_eproj = side[:3]
            cur = torch.cuda.current_stream(x.device)
            ev_fork.record(cur)  # edge_attr, the weights and the arena are ready in the caller's stream order
            side_st.wait_event(ev_fork)
            sst = side_st.cuda_stream
            for i in range(pl.L):
                if bf16:  # (the conversions of edge_attr and of the weights above are ordered before the fork)
                    L.call("isg_linear_bf16_fwd", bf[0].data_ptr(), Dp, bf[1][(i, "P_W_E_BF")], Dp, None,
                           ap + pl.act[i]["P_EPROJ"], pl.HC, None, 0, E, pl.HC, pl.D, L.ACT_NONE, L.BF16, sst)
                else:
                    w_e = model.convs[i].lin_edge.weight
                    L.call("isg_linear_fwd", edge_attr.data_ptr(), pl.D, w_e.data_ptr(), None, None, None,
                           ap + pl.act[i]["P_EPROJ"], pl.HC, None, 0, E, pl.HC, pl.D, L.ACT_NONE, gemm_mode, L.F32, sst)
                ev_eproj[i].record(side_st)
        try:
            for i in range(pl.L):
                d, f, p = _arrays()
                _fill_common(model, pl, gi, i, d, f, p, x_in_ptr, iv.data_ptr() + i * B * pl.D * 4, glf, edge_attr,
                             specs[i], ap, gemm_mode, bf)
                n_launch = kernel_launches(specs[i], gi, False, bf16)
                if side is not None:
                    d[s.D_EPROJ_READY], p[s.P_EV_EPROJ] = 1, ev_eproj[i].cuda_event
                    n_launch -= L.KERNELS_PER_CALL["isg_linear_bf16_fwd" if bf16 else "isg_linear_fwd"]  # counted above
                L.call("isg_mgat_layer_fwd", _p(d), _p(f), _p(p), st, launches=n_launch)
                x_in_ptr = ap + pl.act[i]["P_H_OUT"]
        except BaseException:
            if side is not None:  # the arena goes back to the allocator: order its re-use after the side stream's writes
                torch.cuda.current_stream(x.device).wait_stream(side[0])
            raise
        ctx.model, ctx.gi, ctx.specs, ctx.pl, ctx.gemm_mode, ctx.arena = model, gi, specs, pl, gemm_mode, arena
        ctx.save_for_backward(x, edge_attr, iv, glf, *params)
        D = pl.D
        last = pl.L - 1

        def view(name, shape):
            off = pl.act[last][name]
            n = 1
            for v in shape:
                n *= v
            return arena[off: off + 4 * n].view(torch.float32).view(shape)

        cap = getattr(model, "capture_activations", None) if not bf16 else None
        if cap is not None:  # test hook: the edge kernel's forward inputs of every layer (copies)
            HC = pl.HC
            for i in range(pl.L):
                o = pl.act[i]
                xlr = arena[o["P_XLR"]: o["P_XLR"] + 4 * N * 2 * HC].view(torch.float32).view(N, 2 * HC)
                ep = arena[o["P_EPROJ"]: o["P_EPROJ"] + 4 * E * HC].view(torch.float32).view(E, HC)
                cap[f"x_l.{i}"], cap[f"x_r.{i}"] = xlr[:, :HC].clone(), xlr[:, HC:].clone()
                cap[f"e_proj.{i}"] = ep.clone()
        h = view("P_H_OUT", (N, D))
        if pl.masked[last]:
            return h, view("P_MASK", (N, 1))
        return h, None

    @staticmethod
    def backward(ctx, g_h, g_mask):
        model, gi, specs, pl, arena = ctx.model, ctx.gi, ctx.specs, ctx.pl, ctx.arena
        x, edge_attr, iv, glf = ctx.saved_tensors[:4]
        N, E, B, D = gi.N, gi.E, gi.B, pl.D
        dev = x.device
        s = slots()
        lib = L.load()
        st = L.stream()
        if g_h is None:
            g_h = torch.zeros(N, D, dtype=torch.float32, device=dev)
        g_h = g_h.contiguous()
        g_mask = g_mask.contiguous().to(torch.float32) if g_mask is not None else None
        bucket = getattr(model, "_isg_grad_bucket", None)
        gflat = bucket(pl.grad_numel, dev) if bucket is not None else None
        if gflat is None:
            gflat = torch.empty(pl.grad_numel, dtype=torch.float32, device=dev)
        need_gea = bool(ctx.needs_input_grad[5])
        g_ea = torch.empty_like(edge_attr) if need_gea else None
        g_iv = torch.empty_like(iv)
        any_masked = any(pl.masked)
        g_glf = torch.empty_like(glf) if any_masked else None
        g_x = torch.empty_like(x)
        pong = [torch.empty(N, D, dtype=torch.float32, device=dev) for _ in range(2 if pl.L > 1 else 0)]
        ap = arena.data_ptr()
        ws, ws_bytes = None, 0
        hook = getattr(model, "_isg_after_layer_backward", None)
        side = _side_handles(dev) if _SIDE_WGRAD else None
        defer = side is not None and hook is None and _DEFER_JOIN and pl.L > 1
        ws_pair = [None, None]
        g_in_ptr = g_h.data_ptr()
        first_ea, first_glf = True, True
        for idx, i in enumerate(reversed(range(pl.L))):
            d, f, p = _arrays()
            x_in_ptr = x.data_ptr() if i == 0 else ap + pl.act[i - 1]["P_H_OUT"]
            _fill_common(model, pl, gi, i, d, f, p, x_in_ptr, iv.data_ptr() + i * B * D * 4, glf, edge_attr, specs[i],
                         ap, ctx.gemm_mode, ctx.bf)
            if ws is None:
                ws_bytes = int(lib.isg_mgat_layer_bwd_workspace_bytes(_p(d)))
                ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
                ws_pair = [ws, torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev) if defer else ws]
            d[s.D_WS_BYTES] = ws_bytes
            d[s.D_NEED_GEA] = 1 if need_gea else 0
            d[s.D_ACC_EDGE_ATTR] = 0 if first_ea else 1
            first_ea = False
            if pl.masked[i]:
                d[s.D_ACC_GLF] = 0 if first_glf else 1
                first_glf = False
                p[s.P_G_GLF] = g_glf.data_ptr()
            p[s.P_WS] = ws_pair[idx % 2].data_ptr()
            if side is not None:
                d[s.D_SIDE_WGRAD] = 1
                p[s.P_SIDE_STREAM], p[s.P_EV_FORK], p[s.P_EV_JOIN] = side[0].cuda_stream, side[1][0].cuda_event, \
                    side[1][1].cuda_event
                if defer:
                    d[s.D_DEFER_JOIN] = 1
                    p[s.P_EV_JOIN] = side[3][idx % 2].cuda_event
                    p[s.P_EV_WS_FREE] = side[3][idx % 2].cuda_event if idx >= 2 else 0
            p[s.P_G_H_OUT] = g_in_ptr
            p[s.P_G_MASK_EXT] = g_mask.data_ptr() if (g_mask is not None and i == pl.L - 1 and pl.masked[i]) else 0
            out = g_x if i == 0 else pong[i % 2]
            p[s.P_G_X_IN] = out.data_ptr()
            p[s.P_G_INS] = g_iv.data_ptr() + i * B * D * 4
            p[s.P_G_EDGE_ATTR] = g_ea.data_ptr() if need_gea else 0
            gp = gflat.data_ptr()
            for name, _t in layer_params(model, i):
                if name in ("W_R", "B_R"):
                    continue
                slot = {"W_L": "P_G_W_LR", "B_L": "P_G_B_LR"}.get(name, "P_G_" + name)
                p[getattr(s, slot)] = gp + 4 * pl.grad_off[(i, name)]
            try:
                L.call("isg_mgat_layer_bwd", _p(d), _p(f), _p(p), st,
                       launches=kernel_launches(specs[i], gi, True, pl.bf16))
            except BaseException:
                if defer:  # the workspaces go back to the allocator: order their re-use after the pending products
                    torch.cuda.current_stream(dev).wait_stream(side[0])
                raise
            g_in_ptr = out.data_ptr()
            if hook is not None:
                lo, hi = pl.layer_span[i]
                hook(i, gflat, lo, hi)
        if defer:  # the side stream is serial: the event of the last layer covers every weight gradient
            cur = torch.cuda.current_stream(dev)
            for k in range(min(2, pl.L)):
                cur.wait_event(side[3][(pl.L - 1 - k) % 2])
        grads = []
        for i in range(pl.L):
            for name, t in layer_params(model, i):
                off = pl.grad_off[(i, name)]
                grads.append(gflat[off: off + t.numel()].view(t.shape))
        ctx.arena = None
        return (None, None, None, None, g_x, g_ea, g_iv, g_glf, *grads)


def flat_params(model):
    """The executor's parameters in gradient-layout order (cached: the Parameter objects of a module never change,
    only their .data may be re-pointed)."""
    cached = model.__dict__.get("_isg_flat_params")
    if cached is None or cached[0] != len(model.convs):
        cached = model.__dict__["_isg_flat_params"] = (
            len(model.convs), [t for i in range(len(model.convs)) for _n, t in layer_params(model, i)])
    return cached[1]
