"""TEST INFRASTRUCTURE — CPU restatement (pure torch / numpy) of the ISubGVQA hot path.

This file is the parity oracle for the CUDA kernels in
`intrinsic-subgraph-generation-for-vqa_b200/csrc/`.  It is NOT shipped and NOT on the product
path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import it.  Every function cites the reference file:line it restates (paths relative to
/root/reference).  The third-party arithmetic that is not in the reference tree
(torch_geometric==2.6.1, torch_scatter==2.1.2 — requirements.txt:14-15) is restated from its
published semantics.

PINNING STATUS.  The reference ships no tests / golden vectors, so the oracle is
pinned against OUTPUTS OF THE REFERENCE ITSELF: oracle/make_golden.py imports the unmodified
reference modules from /root/reference (on the shim in oracle/shim), runs them on seeded inputs
with injected noise and stores the results in tests/golden/*.pt; tests/test_oracle_golden.py
checks this restatement against those fixtures everywhere, and
tests/test_oracle_vs_reference.py re-runs the live comparison wherever /root/reference exists.

Gradients are obtained with torch autograd over the restated forward, exactly as the reference
obtains them, except where the reference defines a custom backward (NodeMaskToEdgeMask, IMLE,
AIMLE), which is restated explicitly.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# (a) CSR rebuild of the COO edge_index — no reference counterpart;
# nearest in-tree analogue is the stable sort in sampling/methods/tensor_utils.py:91-114.
# Definition: stable sort of edge ids by destination (resp. source).
# --------------------------------------------------------------------------------------


def csr_build(edge_index, num_nodes):
    ei = edge_index.cpu().numpy().astype(np.int64)
    out = {}
    for name, key, other in (("dst", ei[1], ei[0]), ("src", ei[0], ei[1])):
        order = np.argsort(key, kind="stable")
        ptr = np.zeros(num_nodes + 1, dtype=np.int64)
        np.cumsum(np.bincount(key, minlength=num_nodes), out=ptr[1:])
        out[name + "_ptr"] = torch.from_numpy(ptr.astype(np.int32))
        out[name + "_eid"] = torch.from_numpy(order.astype(np.int32))
        out[name + "_nbr"] = torch.from_numpy(other[order].astype(np.int32))
    return out


# --------------------------------------------------------------------------------------
# third-party segment ops (restated)
# --------------------------------------------------------------------------------------


def _seg_sum(src, index, n):
    shape = (n,) + tuple(src.shape[1:])
    return torch.zeros(shape, dtype=src.dtype, device=src.device).index_add_(0, index, src)


def _seg_max(src, index, n):
    shape = (n,) + tuple(src.shape[1:])
    idx = index.view([-1] + [1] * (src.dim() - 1)).expand_as(src)
    out = torch.full(shape, torch.finfo(src.dtype).min, dtype=src.dtype, device=src.device)
    return out.scatter_reduce(0, idx, src, reduce="amax", include_self=True)


def pyg_softmax(src, index, n):
    """torch_geometric.utils.softmax (call site models/mgat_v2_conv.py:272):
    max over src.detach(); exp(src-max) / (segment_sum + 1e-16)."""
    m = _seg_max(src.detach(), index, n)
    e = (src - m.index_select(0, index)).exp()
    return e / (_seg_sum(e, index, n) + 1e-16).index_select(0, index)


def scatter_softmax(src, index, n):
    """torch_scatter.scatter_softmax (call site utils/scatter_scaled_dot_product.py:7):
    no epsilon, max NOT detached (quirk Q6)."""
    m = _seg_max(src, index, n)
    e = (src - m.index_select(0, index)).exp()
    return e / _seg_sum(e, index, n).index_select(0, index)


def to_dense_batch(x, batch, num_graphs=None):
    """torch_geometric.utils.to_dense_batch (call site models/masking.py:162): ragged [N,*] ->
    dense [B,Nmax,*] filled with 0.0 + bool [B,Nmax]; Nmax = max nodes of THIS batch."""
    B = int(batch.max()) + 1 if num_graphs is None else num_graphs
    counts = torch.bincount(batch, minlength=B)
    nmax = int(counts.max())
    cum = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
    pos = torch.arange(batch.numel(), device=x.device) - cum[batch]
    idx = pos + batch * nmax
    dense = x.new_zeros((B * nmax,) + tuple(x.shape[1:]))
    dense = dense.index_put((idx,), x)  # out-of-place so autograd flows to x
    valid = torch.zeros(B * nmax, dtype=torch.bool, device=x.device)
    valid[idx] = True
    return dense.view((B, nmax) + tuple(x.shape[1:])), valid.view(B, nmax)


# --------------------------------------------------------------------------------------
# conv-side pieces
# --------------------------------------------------------------------------------------


def instr_gate(x, ins, batch):
    """models/mgat_v2_conv.py:156-157: x = gelu(x * instruction[batch]) (exact erf GELU)."""
    return F.gelu(x * ins[batch])


class NodeMaskToEdgeMaskFn(torch.autograd.Function):
    """sampling/node_edge_masks.py:5-19.  fwd: (m[src]*m[dst]).float() [E,1].
    bwd (quirk Q4, NOT the true gradient): grad_m = scatter_sum(grad_out, dst)."""

    @staticmethod
    def forward(ctx, mask, edge_index):
        ctx.save_for_backward(edge_index)
        ctx.n = mask.shape[0]
        return (mask[edge_index[0]] * mask[edge_index[1]]).to(torch.float)

    @staticmethod
    def backward(ctx, grad_output):
        (edge_index,) = ctx.saved_tensors
        return _seg_sum(grad_output, edge_index[1], ctx.n), None


def node_mask_to_edge_mask(mask, edge_index):
    return NodeMaskToEdgeMaskFn.apply(mask, edge_index)


def gat_edge(x_l, x_r, e_proj, att, edge_index, edge_mask=None, negative_slope=0.2):
    """The edge kernel: MaskingGATv2Conv.message (models/mgat_v2_conv.py:243-279) + PyG
    propagate/aggregate (sum over incoming edges, flow source->target, node_dim=0 :47).
    x_l, x_r [N,H,C]; e_proj [E,H,C] (= lin_edge(edge_attr).view(E,H,C), :259-260);
    att [1,H,C]; edge_mask [E,1] fp32 or None.  Returns out [N,H,C] (no bias), alpha [E,H]
    in ORIGINAL COO order.  Masked edges keep logit 0, not -inf (quirk Q3)."""
    src, dst = edge_index[0], edge_index[1]
    x_j = x_l.index_select(0, src)
    x_i = x_r.index_select(0, dst)
    s = x_i + x_j
    s = s + e_proj  # :261
    if edge_mask is not None:
        s = s * edge_mask.unsqueeze(-1)  # :263-264
    s = F.leaky_relu(s, negative_slope)  # :266
    if edge_mask is not None:
        s = s * edge_mask.unsqueeze(-1)  # :268-269
    logit = (s * att).sum(dim=-1)  # :271
    alpha = pyg_softmax(logit, dst, x_r.size(0))  # :272
    if edge_mask is None:
        msg = x_j * alpha.unsqueeze(-1)  # :277-278
    else:
        msg = x_j * (alpha * edge_mask).unsqueeze(-1)  # :279
    out = _seg_sum(msg, dst, x_r.size(0))
    return out, alpha


def scatter_sdpa(query, key, value, batch):
    """utils/scatter_scaled_dot_product.py:6-15: a_n = softmax over the nodes of graph b of
    <query[b], key_n>/sqrt(D); returns a_n * value_n."""
    B = query.size(0)
    logit = (query[batch] * key).sum(-1) / math.sqrt(query.size(-1))
    a = scatter_softmax(logit, batch, B)
    return a.unsqueeze(1) * value


def global_attention_pool(x, u, batch, num_graphs, params, node_mask=None):
    """GlobalAttention.forward (models/att_pooling.py:57-77): node_nn / ques_nn are Linear-GELU-Linear
    (:36-46); gate = <x_n, q[batch_n]>/sqrt(D) (:67-70); torch_geometric.utils.softmax over the nodes of a graph
    (:72); out = scatter_add(gate * x) (:74).  params: dict with node_nn.{0,2}.{weight,bias}, ques_nn.{0,2}.*"""
    x = F.linear(F.gelu(F.linear(x, params["node_nn.0.weight"], params["node_nn.0.bias"])),
                 params["node_nn.2.weight"], params["node_nn.2.bias"])
    if node_mask is not None:
        x = x * node_mask
    q = F.linear(F.gelu(F.linear(u, params["ques_nn.0.weight"], params["ques_nn.0.bias"])),
                 params["ques_nn.2.weight"], params["ques_nn.2.bias"])
    gate = torch.bmm(x.unsqueeze(1), q[batch].unsqueeze(2)).squeeze(-1) / torch.sqrt(torch.tensor(x.size(1)))
    gate = pyg_softmax(gate, batch, num_graphs)
    out = _seg_sum(gate * x, batch, num_graphs)
    return out, gate


def graph_norm(x, batch, weight, bias, mean_scale, num_graphs, eps=1e-5):
    """torch_geometric.nn.norm.GraphNorm(eps=1e-5) (call site models/mgat.py:93-95,171)."""
    cnt = torch.bincount(batch, minlength=num_graphs).clamp(min=1).to(x.dtype).unsqueeze(-1)
    mean = _seg_sum(x, batch, num_graphs) / cnt
    out = x - mean.index_select(0, batch) * mean_scale
    var = _seg_sum(out.pow(2), batch, num_graphs) / cnt
    std = (var + eps).sqrt().index_select(0, batch)
    return weight * out / std + bias


# --------------------------------------------------------------------------------------
# the scene-graph encoding layer in front of MGAT
# --------------------------------------------------------------------------------------


def scene_graph_meta_layer(x, edge_index, edge_attr, p):
    """models/scene_graph_encoder.py:107-146 — torch_geometric.nn.MetaLayer(EdgeModel, NodeModel) with u=None:
      e'  = edge_mlp(cat[x[src], x[dst], e])                       (:118-120; Linear(900,300) GELU Linear(300,300))
      m   = node_mlp_1(cat[x[src], e'])                             (:138-140; Linear(600,300) GELU Linear(300,300))
      agg = scatter_mean(m, dst, dim_size=N)                        (:141)
      x'  = node_mlp_2(cat[x, agg])                                 (:142-143)
    `p` maps the reference's state_dict keys (edge_model.edge_mlp.0.weight, ...) to tensors.  -> (x', e')."""
    src, dst = edge_index[0], edge_index[1]
    N = x.size(0)

    def mlp(prefix, t):
        t = F.linear(t, p[prefix + ".0.weight"], p[prefix + ".0.bias"])
        return F.linear(F.gelu(t), p[prefix + ".2.weight"], p[prefix + ".2.bias"])

    e2 = mlp("edge_model.edge_mlp", torch.cat([x[src], x[dst], edge_attr], 1))
    m = mlp("node_model.node_mlp_1", torch.cat([x[src], e2], 1))
    cnt = torch.bincount(dst, minlength=N).clamp(min=1).to(x.dtype).unsqueeze(-1)
    agg = _seg_sum(m, dst, N) / cnt
    x2 = mlp("node_model.node_mlp_2", torch.cat([x, agg], 1))
    return x2, e2


def scene_graph_encode(x, edge_index, edge_attr, batch, p, gn_weight, gn_bias, gn_mean_scale, num_graphs):
    """models/scene_graph_encoder.py:91-104 — MetaLayer, then GraphNorm evaluated in float64 (the reference moves
    x to the CPU as a DoubleTensor, normalises there and casts back, :99-102).  -> (x_encoded, edge_attr_encoded)."""
    x2, e2 = scene_graph_meta_layer(x, edge_index, edge_attr, p)
    y = graph_norm(x2.double(), batch, gn_weight, gn_bias, gn_mean_scale, num_graphs)  # fp32 params promote
    return y.to(x2.dtype), e2


# --------------------------------------------------------------------------------------
# gate logits theta (MaskingModel.forward up to the sampler)
# --------------------------------------------------------------------------------------


def masking_theta(x, u_nodes, batch, node_w, node_b, ques_w, ques_b):
    """models/masking.py:137,151-155 with the caller's gather (models/mgat_v2_conv.py:166-168):
    xn = GELU(Linear(x)); q = GELU(Linear(u_nodes)) where u_nodes = imle_att[batch] is ALREADY
    [N,D]; the code indexes q[batch] AGAIN (quirk Q1: node n uses row batch[n] of an N-row
    tensor); theta = gelu(<xn, q[batch]> / sqrt(D)) -> [N,1]."""
    xn = F.gelu(F.linear(x, node_w, node_b))
    q = F.gelu(F.linear(u_nodes, ques_w, ques_b))
    gate = torch.bmm(xn.unsqueeze(1), q[batch].unsqueeze(2)).squeeze(-1) / torch.sqrt(
        torch.tensor(float(xn.size(1))))
    return F.gelu(gate)


# --------------------------------------------------------------------------------------
# top-k MAP + IMLE / AIMLE (perturb-and-MAP with perturbation-based gradients)
# --------------------------------------------------------------------------------------


def topk_mask(scores, k):
    """sampling/methods/deterministic_scheme.py:36-43: scores [B,Nmax,E]; k >= Nmax -> ones;
    else thresh = k-th largest along dim 1 and mask = (scores >= thresh).float() (ties -> more
    than k ones, quirk Q5; zero pads compete, quirk Q2)."""
    _, nmax, _ = scores.shape
    if k >= nmax:
        return scores.new_ones(scores.shape)
    thresh = torch.topk(scores, k, dim=1, largest=True, sorted=True).values[:, -1, :][:, None, :]
    return (scores >= thresh).to(torch.float)


class ImleFn(torch.autograd.Function):
    """sampling/methods/wrapper.py:75-172 + target.py:44-48 + imle_scheme.py:16-29, nb_samples=S
    with the noise tensor [B,S,Nmax,1] passed in (reference draws it at noise.py:86-89).
    fwd: z = MAP(theta[:,None] + noise*tau_in) -> [S,B,Nmax,1]
    bwd: z' = MAP(alpha*theta - beta*dy + noise*tau_tgt); grad = mean_S(z - z')."""

    @staticmethod
    def forward(ctx, theta, noise, k, alpha, beta, tau_in, tau_tgt):
        B, S = noise.shape[0], noise.shape[1]
        pert = theta[:, None, ...].repeat(1, S, 1, 1) + noise * tau_in
        z = topk_mask(pert.view((-1,) + tuple(theta.shape[1:])), k).view(noise.shape)
        ctx.save_for_backward(theta, noise, z)
        ctx.cfg = (k, alpha, beta, tau_tgt)
        return z.permute(1, 0, 2, 3)

    @staticmethod
    def backward(ctx, dy):
        theta, noise, z = ctx.saved_tensors
        k, alpha, beta, tau_tgt = ctx.cfg
        S = noise.shape[1]
        dy = dy.permute(1, 0, 2, 3)
        theta_2d = theta[:, None, ...].repeat(1, S, 1, 1).view(dy.shape)
        target = alpha * theta_2d - beta * dy  # target.py:47
        pert = target.view(noise.shape) + noise * tau_tgt
        z2 = topk_mask(pert.view((-1,) + tuple(theta.shape[1:])), k).view(noise.shape)
        return (z - z2).mean(dim=1), None, None, None, None, None, None


class AimleState:
    """sampling/methods/target_aimle.py:87-109 (AdaptiveTargetDistribution state; Python floats,
    grad_norm becomes an fp32 tensor after the first update).  Built at models/masking.py:257-259
    with alpha, beta0 = 0."""

    def __init__(self, alpha=1.0, beta=0.0):
        self.alpha = alpha
        self.beta = beta
        self.grad_norm = 1.0
        self.beta_update_step = 0.0001
        self.beta_update_momentum = 0.0
        self.previous_beta_update = 0.0
        self.grad_norm_decay_rate = 0.9
        self.target_norm = 1.0

    adaptive = True

    def pm(self, theta, dy):  # target_aimle.py:111-115
        norm_dy = torch.linalg.norm(dy).item()
        return 0.0 if norm_dy <= 0.0 else self.beta * (torch.linalg.norm(theta) / norm_dy)


class AimleFixedTarget:
    """sampling/methods/target_aimle.py:30-84 (non-adaptive TargetDistribution(alpha=1, beta=1),
    do_gradient_scaling=False) — what the AIMLE *validation* scheme gets because
    models/masking.py:269 passes target_distribution=None (aimle.py:63-64)."""

    adaptive = False

    def __init__(self, alpha=1.0, beta=1.0):
        self.alpha, self.beta = alpha, beta

    def pm(self, theta, dy):
        return self.beta


class AimleFn(torch.autograd.Function):
    """sampling/methods/aimle.py:83-243 with symmetric_perturbation=True, nb_marginal_samples=1
    (models/masking.py:256-266) and injected noise.  Returns z [B*S,Nmax,1]."""

    @staticmethod
    def forward(ctx, theta, noise, k, state, tau_in, tau_tgt):
        B, S = noise.shape[0], noise.shape[1]
        pert = theta.view(B, 1, -1).repeat(1, S, 1).view(noise.shape) + noise * tau_in
        z2d = topk_mask(pert.view((-1,) + tuple(noise.shape[2:])), k)
        ctx.save_for_backward(theta, noise, z2d.view(noise.shape))
        ctx.cfg = (k, state, tau_tgt)
        return z2d

    @staticmethod
    def backward(ctx, dy):
        theta, noise, z3d = ctx.saved_tensors
        k, st, tau_tgt = ctx.cfg
        B, S = noise.shape[0], noise.shape[1]
        theta_2d = theta.view(B, 1, -1).repeat(1, S, 1).view(dy.shape)
        pm = st.pm(theta_2d, dy)
        t_r = st.alpha * theta_2d - pm * dy  # target_aimle.py:125-128
        t_l = st.alpha * theta_2d - pm * (-dy)
        eps = noise * tau_tgt
        z_r = topk_mask((t_r.view(noise.shape) + eps).view(dy.shape), k).view(noise.shape)
        z_l = topk_mask((t_l.view(noise.shape) + eps).view(dy.shape), k).view(noise.shape)
        g = (z_l - z_r) / 2.0
        if not st.adaptive:  # target_aimle.py:78-84, do_gradient_scaling False
            return g.mean(dim=1), None, None, None, None, None
        # process(): target_aimle.py:130-162
        pm2 = st.pm(theta, dy)
        nnz = torch.count_nonzero(g).float()
        st.grad_norm = st.grad_norm_decay_rate * st.grad_norm + (1.0 - st.grad_norm_decay_rate) * (
            nnz / (g.shape[0] * g.shape[1]))
        upd = (1.0 if float(st.grad_norm) < st.target_norm else -1.0) * st.beta_update_step
        upd = st.beta_update_momentum * st.previous_beta_update + upd
        st.beta = max(st.beta + upd, 0.0)
        st.previous_beta_update = upd
        g = g / (pm2 if pm2 > 0.0 else 1.0)
        return g.mean(dim=1), None, None, None, None, None


# --------------------------------------------------------------------------------------
# Gumbel relaxed top-k and SIMPLE
# --------------------------------------------------------------------------------------

_EPS_TINY = float(np.finfo(np.float32).tiny)


def gumbel_topk(scores, g, k, tau=0.1, hard=True):
    """sampling/methods/gumbel_scheme.py:26-107, policy 'edge_candid', ensemble E=1, with the
    Gumbel(0,1) noise `g` [B,Nmax] passed in (reference draws it on-device at :65-70).
    scores [B,Nmax,1] -> ([1,B,Nmax,1])."""
    B, nmax, _ = scores.shape
    flat = scores.permute(0, 2, 1).reshape(B, nmax)
    local_k = min(k, nmax)
    flat = flat + g
    khot = flat.new_zeros(flat.shape)
    onehot = flat.new_zeros(flat.shape)
    for _ in range(local_k):
        khot_mask = torch.max(1.0 - onehot, torch.tensor([_EPS_TINY], dtype=flat.dtype))
        flat = flat + torch.log(khot_mask)
        onehot = torch.softmax(flat / tau, dim=1)
        khot = khot + onehot
    if hard:
        khot_hard = khot.new_zeros(khot.shape)
        _, ind = torch.topk(khot, local_k, dim=1)
        khot_hard = khot_hard.scatter_(1, ind, 1)
        res = khot_hard - khot.detach() + khot  # quirk Q7: evaluate in this order
    else:
        res = khot
    return res.reshape(1, B, 1, nmax).permute(0, 1, 3, 2)


def _log1mexp(x):
    """sampling/methods/simple.py:44-56: log(1 - exp(-|x|))."""
    x = -x.abs()
    return torch.where(x > -0.6931471805599453094, torch.log(-torch.expm1(x)), torch.log1p(-torch.exp(x)))


def simple_circuit_plan(n, k):
    """Structure of the SDD that sampling/methods/create_simple_constraint.py:34-73 builds for
    "exactly k of n" (n = 2^p): a balanced binary tree; the node (level L, position i, count j) is a
    decomposition whose elements are the pairs (left child count jj, right child count j - jj) for
    which both children exist (:52-60).  A child exists iff its count fits its subtree
    (jj <= min(k, 2^(L-1)); leaves: jj <= 1, :44-47).  Only nodes reachable from the root
    (level p, count k) are evaluated (simple.py:124-139 iterate `beta.positive_iter()`).
    simple.py pads every node's element list to `max_elements` with a dummy node of log-value -1000
    (:192-203, :222) and every node's parent list to `max_parents` with (dummy, 0) (:153-161); those
    pads are NOT neutral once a literal weight is -inf (theta == 0 gives log(1 - e^0) = -inf), so the
    pad counts are part of the function and are reproduced here.
    Returns dict(p, cap[L], reach[L], elems[L][j] = [jj...], max_elements, max_parents)."""
    p = int(round(math.log2(n)))
    assert 2 ** p == n and p >= 1, "SIMPLE needs n = 2^p >= 2 (the reference fails for n = 1 too)"
    cap = [1] + [min(k, 2 ** L) for L in range(1, p + 1)]
    elems = [None] + [{j: [jj for jj in range(j + 1) if jj <= cap[L - 1] and j - jj <= cap[L - 1]]
                       for j in range(cap[L] + 1)} for L in range(1, p + 1)]
    reach = [set() for _ in range(p + 1)]
    reach[p] = {k} if k <= cap[p] else set()
    for L in range(p, 0, -1):
        for j in reach[L]:
            for jj in elems[L][j]:
                reach[L - 1].add(jj)
                reach[L - 1].add(j - jj)
    max_elements = max(len(elems[L][j]) for L in range(1, p + 1) for j in reach[L])
    max_parents = 0
    for L in range(0, p):
        for c in reach[L]:
            # parent entries of a child with count c: one per reachable parent count j that has c as one side
            cnt = sum(1 for j in reach[L + 1] if 0 <= j - c <= cap[L] and (j - c) in reach[L])
            max_parents = max(max_parents, cnt)
    return dict(p=p, cap=cap, reach=[sorted(r) for r in reach], elems=elems, max_elements=max_elements,
                max_parents=max_parents)


_SIMPLE_DUMMY = -1000.0  # simple.py:222  data[self.id] = -float(1000)


def simple_marginals(theta, k):
    """Layer.log_pr(theta).exp() of sampling/methods/simple.py:214-244 restated as a balanced-tree DP
    (no pickle, no node objects): bottom-up pass = levelwiseSL (:16-27), top-down pass = levelwiseMars
    (:30-41), including the -1000 dummy pads.  theta [R, n] (n = Nmax padded to 2^p with -1e10 by
    simple_scheme.py:87-106 — the caller pads).  The negative-literal weight uses theta.detach()
    (simple.py:215-217), so autograd flows only through the positive literals, like the reference.
    Returns marginals [R, n]; differentiable (the reference differentiates through both passes)."""
    R, n = theta.shape
    plan = simple_circuit_plan(n, k)
    p, reach, elems = plan["p"], plan["reach"], plan["elems"]
    me, mp = plan["max_elements"], plan["max_parents"]
    # ---- bottom-up (levelwiseSL): D[L][j] [R, n/2^L] log-values, C[L][j][e] log-conditionals
    D = [{0: _log1mexp(-theta.detach()), 1: theta}]
    C = [None]
    for L in range(1, p + 1):
        prev = D[L - 1]
        dl, cl = {}, {}
        for j in reach[L]:
            terms = [prev[jj][:, 0::2] + prev[j - jj][:, 1::2] for jj in elems[L][j]]
            pads = me - len(terms)
            stack = torch.stack(terms + [torch.full_like(terms[0], 2 * _SIMPLE_DUMMY)] * pads, dim=-1)
            val = torch.logsumexp(stack, dim=-1)
            dl[j] = val
            cl[j] = [t - val for t in terms]
        D.append(dl)
        C.append(cl)
    # ---- top-down (levelwiseMars): M[L][j] log-marginal of node (L, ., j); root: data - data (:229)
    root = D[p][k]
    M = [None] * (p + 1)
    M[p] = {k: root - root}
    for L in range(p - 1, -1, -1):
        ml = {}
        width = n >> L
        for c in reach[L]:
            left, right = [], []  # contributions to even (left-child) and odd (right-child) positions
            for j in reach[L + 1]:
                js = elems[L + 1][j]
                if c in js:  # this node is the LEFT child (prime) of element jj = c
                    left.append(C[L + 1][j][js.index(c)] + M[L + 1][j])
                if (j - c) in js:  # this node is the RIGHT child (sub) of element jj = j - c
                    right.append(C[L + 1][j][js.index(j - c)] + M[L + 1][j])
            out = theta.new_empty(R, width)
            for side, lst in ((0, left), (1, right)):
                pads = mp - len(lst)
                ref_shape = M[L + 1][reach[L + 1][0]]
                stack = torch.stack(lst + [torch.full_like(ref_shape, _SIMPLE_DUMMY)] * pads, dim=-1)
                out[:, side::2] = torch.logsumexp(stack, dim=-1)
            ml[c] = out
        M[L] = ml
    return M[0][1].exp()


def simple_sample(theta_dense, gumbel, k):
    """sampling/methods/simple_scheme.py:44-162 ('edge_candid', E=1, logits_activation None) with
    the Gumbel(0,1) sampling noise `gumbel` [B, n_pad] passed in (reference: simple.py:91-110
    draws uniform noise on-device).  theta_dense [B,Nmax,1] -> (mask [1,B,Nmax,1], marginals [B,Nmax,1])."""
    B, nmax, _ = theta_dense.shape
    flat = theta_dense.permute(0, 2, 1).reshape(B, nmax)
    local_k = min(k, nmax)
    n_pad = 2 ** math.ceil(math.log2(nmax)) if nmax > 1 else 1
    flat = torch.cat([flat, flat.new_full((B, n_pad - nmax), -1.0e10)], dim=1)
    marg = simple_marginals(flat, local_k)
    with torch.no_grad():
        idx = (flat + gumbel).topk(local_k, dim=-1).indices
        hot = torch.zeros_like(flat).scatter_(1, idx, 1.0)
    samples = (hot - marg).detach() + marg
    samples = samples[:, :nmax]
    marg = marg[:, :nmax]
    return samples.reshape(1, B, 1, nmax).permute(0, 1, 3, 2), marg.reshape(B, 1, nmax).permute(0, 2, 1)


# --------------------------------------------------------------------------------------
# module-level restatement with the reference's state_dict layout
# --------------------------------------------------------------------------------------


class _ForcedMask(torch.autograd.Function):
    """Replays a recorded discrete sampler decision (mask) and its perturbation-based gradient.
    Lets the fp64 arbiter evaluate everything AROUND the sampler at high precision while the
    discrete part — compared bit-exactly elsewhere — is held fixed."""

    @staticmethod
    def forward(ctx, theta, mask, g_theta):
        ctx.save_for_backward(g_theta)
        return mask.clone()

    @staticmethod
    def backward(ctx, dy):
        (g_theta,) = ctx.saved_tensors
        return g_theta.clone(), None, None


class OracleMGAT(torch.nn.Module):
    """models/mgat.py:9-184 + mgat_v2_conv.py:21-241 + masking.py:53-199 restated functionally.
    Parameters are held in a flat dict whose keys equal the reference `state_dict()` keys, so
    `load_state_dict(reference.state_dict())` works in both directions.

    forward(..., noise=None, theta_dropout_mask=None): `noise` is the injected sampler noise
    (IMLE/AIMLE: Gumbel(0,0.3) [B,S,Nmax,1]; gumbel: Gumbel(0,1) [B,Nmax]; simple: [B,n_pad]);
    `theta_dropout_mask` [N,1] replaces F.dropout(p=0.2) at masking.py:159 (values 0 or 1/0.8)."""

    def __init__(self, channels=300, num_ins=4, heads=4, masking_thresholds=(1.0, 1.0, 1.0, 0.1),
                 sampler_type="imle", sample_k=2, nb_samples=1, alpha=1.0, beta=10.0, tau=1.0,
                 interpretable_mode=False, use_global_mask=False, concat_instr=False):
        super().__init__()
        self.concat_instr = bool(concat_instr)  # mgat.py:41-44, mgat_v2_conv.py:153-154
        self.C, self.H, self.L = channels, heads, num_ins
        self.thr = [int(t) if t > 1 else t for t in masking_thresholds]
        self.sampler_type, self.k, self.S = sampler_type, sample_k, nb_samples
        self.alpha, self.beta, self.tau = alpha, beta, tau
        self.interpretable_mode, self.use_global_mask = interpretable_mode, use_global_mask
        self.noise_scale = 0.3  # masking.py:215
        self.teacher = None  # dict of forced forward values {"x_l.i", "x_r.i", "e_proj.i"} (see conv())
        self.kink_margin = None  # set to [] to collect min |pre-activation| per layer
        self.record = None  # set to {} to capture the sampler's mask and theta-gradient of a run
        self.replay = None  # set to a recorded dict to force those discrete decisions (fp64 arbiter)
        self.aimle_state = [AimleState(alpha, 0.0) for _ in range(num_ins)]
        D, H = channels, heads
        Din = 2 * D if concat_instr else D
        shapes = {}
        for i in range(num_ins):
            p = f"convs.{i}."
            shapes.update({
                p + "att": (1, H, D), p + "bias": (H * D,),
                p + "lin_l.weight": (H * D, Din), p + "lin_l.bias": (H * D,),
                p + "lin_r.weight": (H * D, Din), p + "lin_r.bias": (H * D,),
                p + "lin_edge.weight": (H * D, D),
                p + "mask.gate_nn.0.weight": (D, D), p + "mask.gate_nn.0.bias": (D,),
                p + "mask.gate_nn.2.weight": (1, D), p + "mask.gate_nn.2.bias": (1,),
                p + "mask.node_nn.0.weight": (D, Din), p + "mask.node_nn.0.bias": (D,),
                p + "mask.ques_nn.0.weight": (D, D), p + "mask.ques_nn.0.bias": (D,),
                p + "mask.gate_top.select.weight": (1, D),
            })
        for i in range(num_ins):
            p = f"x_proj.{i}."
            shapes.update({p + "0.weight": (D * (H // 2), H * D), p + "0.bias": (D * (H // 2),),
                           p + "2.weight": (D, D * (H // 2)), p + "2.bias": (D,)})
        for i in range(num_ins):
            p = f"bns.{i}."
            shapes.update({p + "weight": (D,), p + "bias": (D,), p + "mean_scale": (D,)})
        shapes.update({"node_logits.0.weight": (512, D), "node_logits.0.bias": (512,),
                       "node_logits.2.weight": (2577, 512), "node_logits.2.bias": (2577,)})
        self._keys = list(shapes)
        self._params = torch.nn.ParameterDict(
            {k.replace(".", "__"): torch.nn.Parameter(torch.zeros(s)) for k, s in shapes.items()})

    def p(self, key):
        return self._params[key.replace(".", "__")]

    def state_dict(self, *a, **kw):
        return {k: self.p(k).detach().clone() for k in self._keys}

    def load_state_dict(self, sd, strict=True):
        with torch.no_grad():
            for k in self._keys:
                self.p(k).copy_(sd[k])

    def named_ref_parameters(self):
        return [(k, self.p(k)) for k in self._keys]

    # -- MaskingModel.forward (masking.py:132-199), top-k samplers only
    def _mask(self, i, x, u_nodes, batch, noise, drop_mask, num_graphs):
        pre = f"convs.{i}.mask."
        theta = masking_theta(x, u_nodes, batch, self.p(pre + "node_nn.0.weight"),
                              self.p(pre + "node_nn.0.bias"), self.p(pre + "ques_nn.0.weight"),
                              self.p(pre + "ques_nn.0.bias"))
        if self.training:  # masking.py:159 (F.dropout p=0.2) with an injectable Bernoulli mask
            if drop_mask is not None:
                theta = theta * drop_mask
            else:
                theta = F.dropout(theta, p=0.2, training=True)
        # fp64 arbiter runs: the DISCRETE samplers (IMLE / AIMLE: top-k mask and perturbation gradient) are
        # replayed from an fp32 run; Gumbel and SIMPLE are differentiable relaxations (their only discrete step is
        # a top-k over noise-separated keys) and are evaluated natively in the arbiter's precision.
        if self.replay is not None and self.sampler_type in ("imle", "aimle"):
            mask = _ForcedMask.apply(theta, self.replay["mask"].to(theta.dtype), self.replay["g_theta"].to(theta.dtype))
            return mask, theta
        if self.record is not None and theta.requires_grad:
            theta.register_hook(lambda g: self.record.__setitem__("g_theta", g.detach().clone()))
        dense, valid = to_dense_batch(theta, batch, num_graphs)
        st = self.sampler_type
        if st == "imle":  # masking.py:214-245 ; eval: tau_in = 0 when S == 1 (:238)
            tau_in = self.tau if (self.training or self.S > 1) else 0.0
            beta = self.beta if self.training else 1.0
            alpha = self.alpha if self.training else 1.0
            out = ImleFn.apply(dense, noise, self.k, alpha, beta, tau_in, self.tau)
            return out.squeeze(0)[valid], theta
        if st == "aimle":  # masking.py:248-283 ; eval noise temperature = tau when S == 1 (:275)
            tau_in = self.tau if (self.training or self.S == 1) else 1.0
            state = self.aimle_state[i] if self.training else AimleFixedTarget(1.0, 1.0)
            out = AimleFn.apply(dense, noise, self.k, state, tau_in, self.tau)
            return out[valid], theta
        if st == "gumbel":
            out = gumbel_topk(dense, noise, self.k)
            return out.squeeze(0)[valid], theta
        if st == "simple":
            out, _ = simple_sample(dense, noise, self.k)
            return out.squeeze(0)[valid], theta
        raise ValueError(st)

    def conv(self, i, x, edge_index, batch, edge_attr, ins, imle_att, noise, drop_mask, num_graphs):
        """MaskingGATv2Conv.forward (mgat_v2_conv.py:138-241)."""
        H, C = self.H, self.C
        pre = f"convs.{i}."
        x = torch.cat((x, ins[batch]), dim=1) if self.concat_instr else instr_gate(x, ins, batch)  # :153-157
        mask, em, theta = None, None, None
        if self.thr[i] != 1.0:  # :161
            mask, theta = self._mask(i, x, imle_att[batch], batch, noise, drop_mask, num_graphs)
            em = node_mask_to_edge_mask(mask, edge_index)
        x_l = F.linear(x, self.p(pre + "lin_l.weight"), self.p(pre + "lin_l.bias")).view(-1, H, C)
        x_r = F.linear(x, self.p(pre + "lin_r.weight"), self.p(pre + "lin_r.bias")).view(-1, H, C)
        e_proj = F.linear(edge_attr, self.p(pre + "lin_edge.weight")).view(-1, H, C)
        if self.teacher is not None:
            # Teacher forcing of the edge kernel's forward inputs (values only; gradients still flow
            # through the oracle's own projections).  leaky_relu has a discontinuous derivative, so two
            # fp32 implementations whose pre-activations differ by 1e-6 can disagree by O(1) on single
            # gradient elements wherever |s| ~ 1e-6.  Forcing identical forward values removes that
            # artefact of the reference's own non-smoothness from end-to-end gradient comparisons.
            x_l = x_l + (self.teacher[f"x_l.{i}"].to(x_l.dtype).view_as(x_l) - x_l).detach()
            x_r = x_r + (self.teacher[f"x_r.{i}"].to(x_r.dtype).view_as(x_r) - x_r).detach()
            e_proj = e_proj + (self.teacher[f"e_proj.{i}"].to(e_proj.dtype).view_as(e_proj) - e_proj).detach()
        if self.kink_margin is not None:
            with torch.no_grad():
                s_ = x_r.index_select(0, edge_index[1]) + x_l.index_select(0, edge_index[0]) + e_proj
                if em is not None:
                    s_ = s_[em.view(-1) != 0]
                if s_.numel():
                    self.kink_margin.append(float(s_.abs().min()))
        if getattr(self, "debug_tensors", None) is not None:
            for name, t in (("xg", x), ("x_l", x_l), ("x_r", x_r), ("e_proj", e_proj)):
                t.retain_grad()
                self.debug_tensors[f"{name}.{i}"] = t
        out, alpha = gat_edge(x_l, x_r, e_proj, self.p(pre + "att"), edge_index, em)
        out = out.view(-1, H * C) + self.p(pre + "bias")
        return out, mask, alpha, theta

    def forward(self, x, edge_index, instr_vectors, global_language_feats, edge_attr, batch,
                noise=None, theta_dropout_mask=None, return_aux=False):
        """MGAT.forward (mgat.py:110-184)."""
        B = instr_vectors.shape[1]
        h = x
        mask = None
        aux = {"alpha": [], "theta": None, "conv_out": []}
        if self.use_global_mask:
            global_mask = torch.ones((h.size(0), 1), dtype=h.dtype)
        for i in range(self.L):
            ins = instr_vectors[i]
            conv_res, mask, alpha, theta = self.conv(i, h, edge_index, batch, edge_attr, ins,
                                                     global_language_feats, noise, theta_dropout_mask, B)
            if mask is not None and self.record is not None:
                self.record["mask"] = mask.detach().clone()
            aux["alpha"].append(alpha)
            aux["conv_out"].append(conv_res)
            if theta is not None:
                aux["theta"] = theta
            pre = f"x_proj.{i}."
            conv_res = F.gelu(F.linear(conv_res, self.p(pre + "0.weight"), self.p(pre + "0.bias")))
            conv_res = F.gelu(F.linear(conv_res, self.p(pre + "2.weight"), self.p(pre + "2.bias")))
            aux.setdefault("proj", []).append(conv_res)
            if self.use_global_mask:
                global_mask = mask * global_mask
            conv_res = scatter_sdpa(ins, conv_res, conv_res, batch)
            conv_res = graph_norm(conv_res, batch, self.p(f"bns.{i}.weight"), self.p(f"bns.{i}.bias"),
                                  self.p(f"bns.{i}.mean_scale"), B)
            h = conv_res + h
            aux.setdefault("h", []).append(h)
            if self.use_global_mask:
                h = global_mask * h
            elif self.interpretable_mode and mask is not None:
                h = mask * h
        if return_aux:
            return h, mask, aux
        return h, mask, [], []
