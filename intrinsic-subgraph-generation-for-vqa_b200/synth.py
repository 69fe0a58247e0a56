"""Synthetic GQA-shaped scene-graph batches.  Test/bench input generator —
NOT part of the arithmetic path.  Edge order follows the reference's graph construction
(ISubGVQA/datasets/scene_graph.py:263-343): for each node v ascending, the self-loop (v, v)
first, then each of its relations (v, t) immediately followed by the reverse edge (t, v)
iff (t, v) is not itself a relation; graphs are concatenated with node-index offsets the way
PyG's Batch.from_data_list does (datasets/gqa.py:237-272), so `batch` is sorted and the COO
`edge_index` [2, E] int64 is source-major within every graph."""
import numpy as np
import torch


def _one_graph(rng, n, target_edges, relations=None):
    """Returns (src, dst) int64 arrays for one graph with n nodes in reference order.  `relations`: a list that
    receives the drawn (source, target) relation arrays (tests/test_scene_graph_data.py feeds them through the
    reference-pinned record conversion and expects these very edges)."""
    n_rel = 0
    if n >= 2 and target_edges > n:
        # E = n + 2R - R*p with p ~= P[(t,v) is itself a relation] ~= R / (n(n-1)); two fixed-point steps
        r = (target_edges - n) / 2.0
        for _ in range(3):
            p = min(1.0, r / (n * (n - 1.0)))
            r = (target_edges - n) / (2.0 - p)
        n_rel = max(0, int(round(r)))
    if n_rel == 0:
        v = np.arange(n, dtype=np.int64)
        if relations is not None:
            relations.append((v[:0], v[:0]))
        return v, v.copy()
    rs = np.sort(rng.integers(0, n, size=n_rel), kind="stable").astype(np.int64)
    rt = rng.integers(0, n - 1, size=n_rel).astype(np.int64)
    rt = rt + (rt >= rs)  # uniform over targets != source
    if relations is not None:
        relations.append((rs, rt))
    key = rs * n + rt
    has_rev = np.isin(rt * n + rs, key)
    emit = 2 - has_rev.astype(np.int64)  # relation + optional reverse
    rel_start = np.concatenate([[0], np.cumsum(emit)[:-1]]) + rs + 1  # +1 self-loop per node <= rs
    total = n + int(emit.sum())
    src = np.empty(total, dtype=np.int64)
    dst = np.empty(total, dtype=np.int64)
    # self-loop of node v sits before v's relations: v + (edges emitted by relations of nodes < v)
    emitted_before = np.concatenate([[0], np.cumsum(np.bincount(rs, weights=emit, minlength=n))[:-1]]).astype(np.int64)
    loop_pos = np.arange(n, dtype=np.int64) + emitted_before
    src[loop_pos] = np.arange(n)
    dst[loop_pos] = np.arange(n)
    src[rel_start] = rs
    dst[rel_start] = rt
    rev = ~has_rev
    src[rel_start[rev] + 1] = rt[rev]
    dst[rel_start[rev] + 1] = rs[rev]
    return src, dst


def make_topology(num_graphs, mean_nodes=20, mean_edges=150, seed=3407, max_nodes=126):
    """-> dict(edge_index [2,E] int64, batch [N] int64, num_nodes [B] int64, num_edges [B] int64)."""
    rng = np.random.default_rng(seed)
    srcs, dsts, batches, nn, ne = [], [], [], [], []
    off = 0
    for g in range(num_graphs):
        n = int(round(float(np.clip(rng.normal(mean_nodes, 0.3 * mean_nodes), 2, max_nodes if max_nodes else 1e9))))
        s, d = _one_graph(rng, n, mean_edges)
        srcs.append(s + off)
        dsts.append(d + off)
        batches.append(np.full(n, g, dtype=np.int64))
        nn.append(n)
        ne.append(len(s))
        off += n
    ei = np.stack([np.concatenate(srcs), np.concatenate(dsts)])
    return dict(
        edge_index=torch.from_numpy(ei),
        batch=torch.from_numpy(np.concatenate(batches)),
        num_nodes=torch.tensor(nn, dtype=torch.int64),
        num_edges=torch.tensor(ne, dtype=torch.int64),
    )


def make_batch(num_graphs, channels=300, num_ins=4, mean_nodes=20, mean_edges=150, seed=3407,
               max_nodes=126, dtype=torch.float32):
    """Topology + stand-ins for the encoder outputs that feed MGAT.forward
    (models/isubgvqa.py:255-278): x [N,D], edge_attr [E,D], instr_vectors [num_ins,B,D],
    global_language_feats [B,D], all ~ N(0,1) from a seeded CPU generator."""
    topo = make_topology(num_graphs, mean_nodes, mean_edges, seed, max_nodes)
    g = torch.Generator().manual_seed(seed)
    N = topo["batch"].numel()
    E = topo["edge_index"].size(1)
    out = dict(topo)
    out["x"] = torch.randn(N, channels, generator=g).to(dtype)
    out["edge_attr"] = torch.randn(E, channels, generator=g).to(dtype)
    out["instr_vectors"] = torch.randn(num_ins, num_graphs, channels, generator=g).to(dtype)
    out["global_language_feats"] = torch.randn(num_graphs, channels, generator=g).to(dtype)
    out["nmax"] = int(topo["num_nodes"].max())
    return out


def subset_topology(topo, graph_ids):
    """The sub-batch made of the given graphs (ascending ids) of a make_topology() batch, nodes renumbered the way
    Batch.from_data_list would number them."""
    ids = torch.as_tensor(sorted(int(g) for g in graph_ids), dtype=torch.int64)
    B = int(topo["num_nodes"].numel())
    new_gid = torch.full((B,), -1, dtype=torch.int64)
    new_gid[ids] = torch.arange(ids.numel())
    keep_n = new_gid[topo["batch"]] >= 0
    remap = torch.full((topo["batch"].numel(),), -1, dtype=torch.int64)
    remap[keep_n] = torch.arange(int(keep_n.sum()))
    ei = topo["edge_index"]
    keep_e = keep_n[ei[0]]
    return dict(edge_index=remap[ei[:, keep_e]], batch=new_gid[topo["batch"][keep_n]],
                num_nodes=topo["num_nodes"][ids], num_edges=topo["num_edges"][ids])


def make_batch_from_topology(topo, channels=300, num_ins=4, seed=3407, dtype=torch.float32):
    """make_batch() on a given topology (features from the seeded CPU generator)."""
    g = torch.Generator().manual_seed(seed)
    N, E, B = topo["batch"].numel(), topo["edge_index"].size(1), int(topo["num_nodes"].numel())
    out = dict(topo)
    out["x"] = torch.randn(N, channels, generator=g).to(dtype)
    out["edge_attr"] = torch.randn(E, channels, generator=g).to(dtype)
    out["instr_vectors"] = torch.randn(num_ins, B, channels, generator=g).to(dtype)
    out["global_language_feats"] = torch.randn(B, channels, generator=g).to(dtype)
    out["nmax"] = int(topo["num_nodes"].max())
    return out


def gumbel_noise(num_graphs, nmax, scale=0.3, seed=3407, nb_samples=1):
    """Gumbel(0, scale) noise [B, S, Nmax, 1] from the CPU generator — the tensor that is injected
    into both the reference sampler and the CUDA sampler for bit-exact mask parity
    (sampling/methods/noise.py:86-89 draws it with torch.distributions.Gumbel on the CPU)."""
    g = torch.Generator().manual_seed(seed + 7919)
    u = torch.rand(num_graphs, nb_samples, nmax, 1, generator=g).clamp_(1e-10, 1.0 - 1e-7)
    return -scale * torch.log(-torch.log(u))


def mgat_param_shapes(channels=300, heads=4, num_ins=4, concat_instr=False):
    """state_dict layout of the reference MGAT (enumerated from models/mgat.py:55-102,
    mgat_v2_conv.py:63-128, masking.py:77-90)."""
    D, H = channels, heads
    Din = 2 * D if concat_instr else D  # mgat.py:41-44: the conv sees [x, instruction[batch]]
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
    return shapes


def make_state_dict(channels=300, heads=4, num_ins=4, seed=3407, concat_instr=False):
    """Deterministic random-init weights in the reference MGAT layout, independent of module
    construction order (so goldens need not store 42 MB of weights).  Matrices ~ U(+-sqrt(6/(in+out))),
    vectors ~ small noise around their neutral value so every parameter is exercised."""
    g = torch.Generator().manual_seed(seed + 1)
    sd = {}
    for k, s in mgat_param_shapes(channels, heads, num_ins, concat_instr).items():
        if len(s) >= 2:
            a = (6.0 / (s[-1] + s[-2])) ** 0.5
            t = (torch.rand(s, generator=g) * 2 - 1) * a
        else:
            t = 0.1 * torch.randn(s, generator=g)
            if k.endswith("bns.weight") or k.split(".")[-1] in ("weight", "mean_scale") and k.startswith("bns."):
                t = t + 1.0
        sd[k] = t
    return sd


def sgenc_param_shapes(nf=300, ef=300, hd=300):
    """state_dict layout of the reference's scene-graph encoding MetaLayer (models/scene_graph_encoder.py:107-146)."""
    return {
        "edge_model.edge_mlp.0.weight": (hd, 2 * nf + ef), "edge_model.edge_mlp.0.bias": (hd,),
        "edge_model.edge_mlp.2.weight": (hd, hd), "edge_model.edge_mlp.2.bias": (hd,),
        "node_model.node_mlp_1.0.weight": (hd, nf + hd), "node_model.node_mlp_1.0.bias": (hd,),
        "node_model.node_mlp_1.2.weight": (hd, hd), "node_model.node_mlp_1.2.bias": (hd,),
        "node_model.node_mlp_2.0.weight": (hd, nf + hd), "node_model.node_mlp_2.0.bias": (hd,),
        "node_model.node_mlp_2.2.weight": (hd, hd), "node_model.node_mlp_2.2.bias": (hd,),
    }


def make_sgenc_state_dict(nf=300, ef=300, hd=300, seed=3407):
    """Deterministic weights for the MetaLayer (torch.nn.Linear-style scale) + a GraphNorm (weight, bias, mean_scale)."""
    g = torch.Generator().manual_seed(seed + 2)
    sd = {}
    for k, s in sgenc_param_shapes(nf, ef, hd).items():
        a = 1.0 / (s[-1] ** 0.5) if len(s) == 2 else 0.05
        sd[k] = (torch.rand(s, generator=g) * 2 - 1) * a
    gn = {"weight": 1.0 + 0.1 * torch.randn(hd, generator=g), "bias": 0.1 * torch.randn(hd, generator=g),
          "mean_scale": 1.0 + 0.1 * torch.randn(hd, generator=g)}
    return sd, gn
