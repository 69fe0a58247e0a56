"""batch construction with per-image cached CSR (isg_b200.collate).  The CPU tests check
the collated batch against PyG's Batch.from_data_list layout and the concatenated CSR against the oracle's stable
sort of the whole batch; the GPU test checks it against the device build (isg_csr_build) bit for bit and runs MGAT on
the uploaded index."""
import numpy as np
import pytest
import torch

import util
from isg_b200 import collate, synth


def _graphs(B, seed, mn=9, me=40, C=16):
    b = synth.make_batch(B, channels=C, mean_nodes=mn, mean_edges=me, seed=seed)
    gs = []
    for g in range(B):
        nodes = (b["batch"] == g).nonzero().flatten()
        n0 = int(nodes[0]) if nodes.numel() else 0
        keep = (b["batch"][b["edge_index"][0]] == g)
        gs.append(dict(x=b["x"][nodes], edge_index=b["edge_index"][:, keep] - n0, edge_attr=b["edge_attr"][keep],
                       image_id=f"img{seed}_{g}"))
    return b, gs


def test_collate_matches_from_data_list_layout_and_oracle_csr():
    import isg_oracle as O

    b, gs = _graphs(7, 3)
    cache = collate.SceneGraphCsrCache()
    out = collate.collate_scene_graphs(gs, cache)
    for k in ("x", "edge_index", "edge_attr", "batch"):
        assert torch.equal(out[k], b[k]), k
    want = O.csr_build(b["edge_index"], b["x"].shape[0])
    hi = out["host_index"]
    for k in ("dst_ptr", "dst_eid", "dst_nbr", "src_ptr", "src_eid", "src_nbr"):
        assert torch.equal(hi[k], want[k]), k
    assert hi["nmax"] == b["nmax"] and hi["num_graphs"] == 7
    assert torch.equal(hi["graph_ptr"].long(), torch.cat([torch.zeros(1, dtype=torch.long),
                                                          torch.bincount(b["batch"], minlength=7).cumsum(0)]))
    N, E = b["x"].shape[0], b["edge_index"].shape[1]
    for side in ("dst", "src"):
        assert torch.equal(hi[side + "_order"].long(), util.heavy_first_order(hi[side + "_ptr"], N, E, side))
    # second epoch: every image comes out of the cache; a shuffled batch of the same images is consistent
    assert cache.misses == 7 and cache.hits == 0
    perm = [4, 0, 6, 2]
    out2 = collate.collate_scene_graphs([gs[i] for i in perm], cache)
    assert cache.hits == 4 and cache.misses == 7
    want2 = O.csr_build(out2["edge_index"], out2["x"].shape[0])
    for k in ("dst_ptr", "dst_eid", "dst_nbr", "src_ptr", "src_eid", "src_nbr"):
        assert torch.equal(out2["host_index"][k], want2[k]), k


def test_collate_rejects_out_of_range_edges_and_handles_edgeless_graphs():
    g0 = dict(x=torch.zeros(3, 4), edge_index=torch.zeros(2, 0, dtype=torch.int64), edge_attr=torch.zeros(0, 4))
    g1 = dict(x=torch.zeros(2, 4), edge_index=torch.tensor([[0, 1], [1, 0]]), edge_attr=torch.zeros(2, 4))
    out = collate.collate_scene_graphs([g0, g1])
    assert out["host_index"]["dst_ptr"].tolist() == [0, 0, 0, 0, 1, 2]
    assert out["edge_index"].tolist() == [[3, 4], [4, 3]]
    with pytest.raises(IndexError):
        collate.GraphCsr(torch.tensor([[0], [5]]), 3)


@pytest.mark.gpu
def test_host_index_equals_device_build_and_drives_mgat():
    from isg_b200.graph import GraphIndex, clear_cache, register_graph_index
    from isg_b200.isubgvqa import MGAT

    dev = "cuda"
    b, gs = _graphs(24, 11, mn=14, me=90, C=300)
    out = collate.collate_scene_graphs(gs, collate.SceneGraphCsrCache(), pin=True)
    ei, batch = out["edge_index"].to(dev), out["batch"].to(dev)
    built = GraphIndex(ei, batch, 24)
    up = GraphIndex.from_host(ei, batch, out["host_index"])
    for k in ("dst_ptr", "dst_nbr", "dst_eid", "src_ptr", "src_nbr", "src_eid", "graph_ptr"):
        assert torch.equal(getattr(built, k), getattr(up, k)), k
    assert torch.equal(built.batch32[: up.N], up.batch32[: up.N]) and built.nmax == up.nmax and up.closed
    model = MGAT(channels=300, num_ins=4, heads=4, use_instr=True, masking_thresholds=[1.0, 1.0, 1.0, 0.1], use_topk=True,
                 interpretable_mode=False, sampler_type="imle", sample_k=2).to(dev).eval()
    model.load_state_dict(synth.make_state_dict(300, 4, 4, 11))
    args = (out["x"].to(dev), ei, b["instr_vectors"].to(dev), b["global_language_feats"].to(dev), out["edge_attr"].to(dev), batch)
    noise = util.case_noise("imle", 24, b["nmax"], 11).to(dev)
    hs = []
    for use_host in (False, True):
        clear_cache()
        if use_host:
            register_graph_index(up)
        model.convs[3].mask.injected_noise = noise
        with torch.no_grad():
            h, mask, _, _ = model(*args, return_masks=True)
        hs.append((h.clone(), mask.clone()))
    assert torch.equal(hs[0][0], hs[1][0]) and torch.equal(hs[0][1], hs[1][1])
