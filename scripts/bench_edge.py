"""Edge-kernel roofline study (BASELINE.json config 5 and a quick single-size check): times isg_gat_edge_fwd /
isg_gat_edge_bwd alone on synthetic GQA-shaped graphs, CUDA events on the launching stream, L2 flushed between
repetitions, and reports achieved algorithmic GB/s (SURVEY.md §8d byte counts) against the measured HBM peak.

    python scripts/bench_edge.py                     # B=256, 20 nodes / 150 edges (the c3 training size)
    python scripts/bench_edge.py --sweep             # B=4096, (10,50) (20,150) (50,600) (100,1500) (200,4000)

Large points are processed in chunks of whole graphs (the [E,1200] fp32 edge projections of the largest point are
78 GB); bytes and times are summed over the chunks."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import isg_b200  # noqa: E402,F401
from isg_b200 import lib as L  # noqa: E402
from isg_b200 import synth  # noqa: E402
from isg_b200.graph import GraphIndex  # noqa: E402

H, C = 4, 300
HC = H * C


def edge_bytes(N, E, masked, s=4):
    fwd = s * HC * (E + 3 * N) + 4 * E * H + (4 * E if masked else 0) + 4 * (2 * E + N + 1) + 2 * s * HC
    bwd = s * HC * (2 * E + 5 * N) + 4 * E * H + (8 * E if masked else 0) + 4 * (4 * E + 2 * N + 2) + 2 * s * HC
    return fwd, bwd


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def run_point(B, mn, me, masked, reps, chunk_graphs, seed=7, bf16=False):
    dev = torch.device("cuda")
    lib = L.load()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tot = dict(fwd_ms=0.0, bwd_ms=0.0, fwd_b=0, bwd_b=0, N=0, E=0)
    done = 0
    while done < B:
        nb = min(chunk_graphs, B - done)
        g = synth.make_topology(nb, mean_nodes=mn, mean_edges=me, seed=seed + done, max_nodes=None)
        ei, batch = g["edge_index"].to(dev), g["batch"].to(dev)
        if os.environ.get("BENCH_SORT_DEGREE"):  # experiment: relabel nodes by descending in-degree
            deg = torch.bincount(ei[1], minlength=batch.numel())
            order = torch.argsort(deg, descending=True, stable=True)
            rank = torch.empty_like(order); rank[order] = torch.arange(order.numel(), device=dev)
            ei = rank[ei]
            print("max/mean in-degree", int(deg.max()), float(deg.float().mean()), file=sys.stderr)
        N, E = int(batch.numel()), int(ei.shape[1])
        gi = GraphIndex(ei, batch, nb)
        gen = torch.Generator(device=dev).manual_seed(seed)
        fdt = torch.bfloat16 if bf16 else torch.float32
        es = 2 if bf16 else 4
        code = 1 if bf16 else 0
        xlr = torch.randn(N, 2 * HC, device=dev, generator=gen).to(fdt)
        ep = torch.randn(E, HC, device=dev, generator=gen).to(fdt)
        att = torch.randn(HC, device=dev, generator=gen) * 0.1
        bias = torch.zeros(HC, device=dev)
        em = (torch.rand(E, device=dev, generator=gen) > 0.3).float() if masked else None
        out = torch.empty(N, HC, device=dev, dtype=fdt)
        alpha = torch.empty(E, H, device=dev)
        gout = torch.randn(N, HC, device=dev, generator=gen).to(fdt)
        gxlr = torch.empty(N, 2 * HC, device=dev, dtype=fdt)
        gep = torch.empty(E, HC, device=dev, dtype=fdt)
        gatt = torch.empty(HC, device=dev)
        gem = torch.empty(E, device=dev) if masked else None
        wsb = lib.isg_gat_edge_bwd_workspace_bytes(N, E, nb, H, C)
        from isg_b200 import ops
        fused = ops.edge_bwd_fused(gi) and not bf16
        ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dev)
        st = L.stream()
        order = (None, None) if os.environ.get("ISG_EDGE_ORDER", "1") == "0" else (gi.dst_order, gi.src_order)

        def fwd():
            L.call("isg_gat_edge_fwd", xlr.data_ptr(), xlr.data_ptr() + HC * es, 2 * HC, ep.data_ptr(), att.data_ptr(),
                   bias.data_ptr(), L.ptr(em), L.ptr(gi.dst_ptr), L.ptr(gi.dst_nbr), L.ptr(gi.dst_eid), L.ptr(order[0]),
                   out.data_ptr(), HC, alpha.data_ptr(), N, E, H, C, 0.2, code, st)

        def bwd():
            L.call("isg_gat_edge_bwd", gout.data_ptr(), HC, xlr.data_ptr(), xlr.data_ptr() + HC * es, 2 * HC,
                   ep.data_ptr(), att.data_ptr(), bias.data_ptr(), L.ptr(em), alpha.data_ptr(), out.data_ptr(), HC,
                   L.ptr(gi.dst_ptr), L.ptr(gi.dst_nbr), L.ptr(gi.dst_eid), L.ptr(order[0]), L.ptr(gi.src_ptr),
                   L.ptr(gi.src_nbr), L.ptr(gi.src_eid), L.ptr(order[1]), gxlr.data_ptr(), gxlr.data_ptr() + HC * es, 2 * HC, gep.data_ptr(),
                   gatt.data_ptr(), L.ptr(gem), N, E, H, C, 0.2, code, L.ptr(gi.batch32) if fused else None,
                   L.ptr(gi.graph_ptr) if fused else None, nb, gi.nmax if fused else 0, ws.data_ptr(), wsb, st)

        for fn, key in ((fwd, "fwd_ms"), (bwd, "bwd_ms")):
            for _ in range(3):
                fn()
            times = []
            for _ in range(reps):
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                times.append(a.elapsed_time(b))
            times.sort()
            tot[key] += times[len(times) // 2]
        fb, bb = edge_bytes(N, E, masked, es)
        tot["fwd_b"] += fb
        tot["bwd_b"] += bb
        tot["N"] += N
        tot["E"] += E
        done += nb
        del xlr, ep, out, alpha, gout, gxlr, gep, ws, gi
        torch.cuda.empty_cache()
    return tot


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--masked", action="store_true")
    ap.add_argument("--bf16", action="store_true", help="bf16 storage of the feature tensors (fp32 accumulation)")
    args = ap.parse_args()
    pk, src = peak()
    points = [(4096, 10, 50), (4096, 20, 150), (4096, 50, 600), (4096, 100, 1500), (4096, 200, 4000)] \
        if args.sweep else [(256, 20, 150), (1024, 20, 150)]
    if os.environ.get("BENCH_POINT"):  # e.g. BENCH_POINT=1024,100,1500
        points = [tuple(int(v) for v in os.environ["BENCH_POINT"].split(","))]
    for B, mn, me in points:
        # keep e_proj + g_eproj of one chunk under ~40 GB
        chunk = max(1, min(B, int(40e9 / (2 * 4 * HC * me))))
        t = run_point(B, mn, me, args.masked, args.reps if B * me < 2e6 else max(3, args.reps // 4), chunk,
                      bf16=args.bf16)
        f = t["fwd_b"] / (t["fwd_ms"] * 1e-3) / 1e9
        b = t["bwd_b"] / (t["bwd_ms"] * 1e-3) / 1e9
        print(json.dumps({"graphs": B, "mean_nodes": mn, "mean_edges": me, "N": t["N"], "E": t["E"],
                          "masked": args.masked, "storage": "bf16" if args.bf16 else "fp32", "fwd_ms": round(t["fwd_ms"], 4), "bwd_ms": round(t["bwd_ms"], 4),
                          "fwd_GBps": round(f, 1), "bwd_GBps": round(b, 1), "fwd_frac": round(f / pk, 4),
                          "bwd_frac": round(b / pk, 4), "peak_GBps": pk, "peak_source": src,
                          "chunk_graphs": chunk}), flush=True)


if __name__ == "__main__":
    main()
