"""SURVEY.md section 8 row f3, the loader side: isg_b200.loader.DevicePrefetcher (pinned host batch -> HBM + CSR build
on a copy stream, one batch ahead) and HostResults (results read by the host one step later) give the SAME numbers as
the synchronous path the reference runs (training/train_epoch.py:66-73, 120-133: `.to("cuda")` on the compute stream,
`.item()` right after backward) — three different batches back to back, training steps, IMLE sampler."""
import pytest
import torch

import util
from isg_b200 import collate, synth

pytestmark = pytest.mark.gpu

KEYS = ("x", "edge_index", "instr_vectors", "global_language_feats", "edge_attr", "batch")


def _model(C, seed, dev):
    from isg_b200.isubgvqa import MGAT

    model = MGAT(channels=C, num_ins=4, heads=4, use_instr=True, masking_thresholds=[1.0, 1.0, 1.0, 0.1], use_topk=True,
                 interpretable_mode=False, sampler_type="imle", sample_k=2, nb_samples=1, alpha=1.0, beta=10.0, tau=1.0)
    model.load_state_dict(synth.make_state_dict(C, 4, 4, seed))
    return model.to(dev).train()


def _step(model, t, noise, drop):
    model.convs[3].mask.injected_noise = noise
    model.convs[3].mask.injected_dropout_mask = drop
    x = t["x"].detach().requires_grad_(True)
    model.zero_grad()
    h, mask, _, _ = model(x, t["edge_index"], t["instr_vectors"], t["global_language_feats"], t["edge_attr"], t["batch"],
                          return_masks=True)
    loss = util.loss_fn(h)
    loss.backward()
    return loss, mask, x.grad


def _host_batches(B, C, seed, n):
    out = []
    for i in range(n):
        b = synth.make_batch(B, channels=C, mean_nodes=8 + 3 * i, mean_edges=40 + 25 * i, seed=seed + i)
        host = {k: b[k].pin_memory() for k in KEYS}
        extra = {"noise": util.case_noise("imle", B, b["nmax"], seed + i).pin_memory(),
                 "drop": util.case_dropout(b["x"].shape[0], True, seed + i).pin_memory()}
        out.append((b, host, extra))
    return out


@pytest.mark.parametrize("collated", [False, True], ids=["device_csr_build", "collate_time_csr"])
def test_prefetched_steps_equal_synchronous_steps(collated):
    from isg_b200.graph import clear_cache
    from isg_b200.loader import DevicePrefetcher, HostResults

    dev = torch.device("cuda")
    B, C, seed, n = 10, 64, 77, 3
    batches = _host_batches(B, C, seed, n)
    model = _model(C, seed, dev)

    # the synchronous path: copy on the compute stream, CSR built inside the forward pass, .item() after backward
    want = []
    for b, host, extra in batches:
        clear_cache()
        t = {k: host[k].to(dev) for k in KEYS}
        loss, mask, gx = _step(model, t, extra["noise"].to(dev), extra["drop"].to(dev))
        want.append((float(loss.item()), mask.detach().cpu().clone(), gx.detach().cpu().clone()))
    clear_cache()

    if collated:  # the index comes from the collate-time per-image cache instead of the device build
        cache = collate.SceneGraphCsrCache()
        for j, (b, host, _extra) in enumerate(batches):
            gs = []
            for g in range(B):
                nodes = (b["batch"] == g).nonzero().flatten()
                keep = b["batch"][b["edge_index"][0]] == g
                gs.append(dict(x=b["x"][nodes], edge_index=b["edge_index"][:, keep] - int(nodes[0]),
                               edge_attr=b["edge_attr"][keep], image_id=f"img{j}_{g}"))
            out = collate.collate_scene_graphs(gs, cache, pin=True)
            assert torch.equal(out["edge_index"], b["edge_index"]) and torch.equal(out["batch"], b["batch"])
            host["host_index"] = out["host_index"]

    pre, results = DevicePrefetcher(dev), HostResults(lag=1)
    pending = [pre.stage(batches[0][1], extra=batches[0][2], nmax=batches[0][0]["nmax"])]
    got, grads = [], []
    for i in range(n):
        t, ext = pre.get(pending.pop(0))
        if i + 1 < n:  # batch i+1 is copied (and indexed) while batch i computes
            pending.append(pre.stage(batches[i + 1][1], extra=batches[i + 1][2], nmax=batches[i + 1][0]["nmax"]))
        loss, mask, gx = _step(model, t, ext["noise"], ext["drop"])
        grads.append(gx)
        results.push(loss=loss, mask=mask)
        r = results.pop()
        assert (r is None) == (i == 0), "lag 1: nothing to read after the first push, then one result per step"
        if r is not None:
            got.append(r)
    got += results.drain()
    assert len(got) == n and results.pop() is None
    torch.cuda.synchronize()
    for i in range(n):
        w_loss, w_mask, w_gx = want[i]
        assert isinstance(got[i]["loss"], float)
        assert abs(got[i]["loss"] - w_loss) <= 1e-6 * max(1.0, abs(w_loss)), (i, got[i]["loss"], w_loss)
        assert torch.equal(got[i]["mask"].cpu(), w_mask), f"node mask of batch {i}"
        assert util.rel_err(grads[i].cpu(), w_gx) <= 1e-6, (i, util.rel_err(grads[i].cpu(), w_gx))
