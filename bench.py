#!/usr/bin/env python
"""bench.py — scene graphs/s of the ISubGVQA hot path (MGAT fwd+bwd + sampler) on B200.

    python bench.py --gpus N --steps K --warmup W [--workload c3|c1|c2] [--impl isg|reference]

A "step" is one pass of the hot path over one batch of synthetic GQA-shaped scene graphs:
4-layer MGAT forward + backward (+ the gradient all-reduce when N > 1).  Default workload = BASELINE.json
config 3/4 (training step, AIMLE sampler, 256 graphs per GPU, fp32): the metric is quoted as
"scene graphs/sec fwd+bwd at 1/2/4/8 B200", i.e. on the training configuration, which is also the one
that weak-scales over GPUs (config 4).  `value` times the step with inputs resident in HBM; `e2e` times
the same step through the public nn.Module API with pinned HOST buffers (H2D of the batch and D2H of
loss + mask inside the timed region).  `roofline` is for the dominant HBM kernel of the path (the fused
edge-attention backward), measured live with CUDA events inside the timed steps.  `cpu_baseline` /
`--impl reference` time the CPU restatement of the reference (oracle/, kind "port": the reference's PyG
dependencies are not installable offline and /root/reference does not exist on the GPU box).
"""
import argparse
import contextlib
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (description, sampler, train, graphs per GPU)
    "c3": ("BASELINE config 3/4: training step, AIMLE top-k sampler, 256 graphs/GPU, fp32", "aimle", True, 256),
    "c1": ("BASELINE config 1: fwd+bwd, 64 graphs, GAT + IMLE top-k", "imle", True, 64),
    "c2": ("BASELINE config 2: inference only, 1024 graphs, Gumbel top-k", "gumbel", False, 1024),
}
CHANNELS, HEADS, LAYERS, K_SAMPLE = 300, 4, 4, 2
METRIC, UNIT = "scene_graphs_per_sec_fwd_bwd", "graphs/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks/throttle reasons during the timed region (recipe in B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def edge_bytes(N, E, masked, s=4):
    """Algorithmic HBM bytes of the fused edge kernel per layer (unfused lin_edge)."""
    HC, H = HEADS * CHANNELS, HEADS
    fwd = s * HC * (E + 3 * N) + 4 * E * H + (4 * E if masked else 0) + 4 * (2 * E + N + 1) + 2 * s * HC
    bwd = s * HC * (2 * E + 5 * N) + 4 * E * H + (8 * E if masked else 0) + 4 * (4 * E + 2 * N + 2) + 2 * s * HC
    return fwd, bwd


def gemm_flops(N, E, B, train):
    """Algorithmic FLOPs of the dense projections per step: per layer forward
    2*300*1200*(2N + E) + 2N*(1200*600 + 600*300), the masked layer adds 2*(N + B)*300*300; the backward
    (dgrad + wgrad of every projection) is twice the forward."""
    D, HC, HID = CHANNELS, HEADS * CHANNELS, CHANNELS * (HEADS // 2)
    fwd = LAYERS * (2 * D * HC * (2 * N + E) + 2 * N * (HC * HID + HID * D)) + 2 * (N + B) * D * D
    return fwd * (3 if train else 1)


def sampler_bytes(N, E, B, nmax):
    """Algorithmic HBM bytes of the fused sampler forward (gate dot + dropout + noise + top-k + node mask + edge
    mask, one launch): xn [N,D] + q [B,D] + keep/theta/mask [N] each + noise/z_dense [B,Nmax] each + graph_ptr +
    the dst-sorted CSR (ptr, nbr, eid) + edge_mask [E]."""
    return 4 * (N * CHANNELS + B * CHANNELS + 3 * N + 2 * B * nmax + (B + 1) + (N + 1) + 3 * E)


def sampler_study(N, E, B, nmax, gi, dev, reps=50):
    """The fused sampler forward timed alone (L2 flushed between launches): SURVEY.md §8d asks for its achieved
    GB/s 'honestly' — it is latency-bound at every BASELINE size (~20 B per node)."""
    from isg_b200 import lib as L

    g = torch.Generator(device=dev).manual_seed(1)
    xn = torch.randn(N, CHANNELS, device=dev, generator=g)
    q = torch.randn(B, CHANNELS, device=dev, generator=g)
    keep = (torch.rand(N, device=dev, generator=g) > 0.2).float() / 0.8
    noise = torch.randn(B, nmax, device=dev, generator=g) * 0.3
    theta, mask, em = torch.empty(N, device=dev), torch.empty(N, device=dev), torch.empty(E, device=dev)
    zd = torch.empty(B, nmax, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = L.stream()

    def run():
        L.call("isg_sampler_fused_fwd", xn.data_ptr(), q.data_ptr(), keep.data_ptr(), noise.data_ptr(),
               gi.batch32.data_ptr(), gi.graph_ptr.data_ptr(), gi.dst_ptr.data_ptr(), gi.dst_nbr.data_ptr(),
               gi.dst_eid.data_ptr(), B, CHANNELS, 1, nmax, K_SAMPLE, 1.0, theta.data_ptr(), mask.data_ptr(),
               zd.data_ptr(), em.data_ptr(), st)

    for _ in range(5):
        run()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def edge_study(points=((4096, 20, 150),), reps=10):
    """BASELINE config 5 (edge-kernel roofline study), one point by default: the fused edge kernels timed alone
    at batch 4096 (scripts/bench_edge.py runs the whole 10-200 objects / 50-4000 edges sweep)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_edge", os.path.join(ROOT, "scripts", "bench_edge.py"))
    be = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(be)
    pk, _ = be.peak()
    out = []
    for B, mn, me in points:
        t = be.run_point(B, mn, me, False, reps, B)
        f = t["fwd_b"] / (t["fwd_ms"] * 1e-3) / 1e9
        b = t["bwd_b"] / (t["bwd_ms"] * 1e-3) / 1e9
        out.append({"graphs": B, "mean_nodes": mn, "mean_edges": me, "N": t["N"], "E": t["E"],
                    "fwd_ms": round(t["fwd_ms"], 4), "bwd_ms": round(t["bwd_ms"], 4), "fwd_GBps": round(f, 1),
                    "bwd_GBps": round(b, 1), "fwd_frac": round(f / pk, 4), "bwd_frac": round(b / pk, 4)})
    return out


def measured_traffic(kernel_key, workload):
    """DRAM bytes per launch of the roofline kernel from the committed `ncu --set full` capture of THIS workload
    (profiles/edge_traffic.json: {workload: {entry point: bytes}}, written by scripts/ncu_summary.py --traffic);
    None when no capture of this workload is committed."""
    p = os.path.join(ROOT, "profiles", "edge_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        return json.load(open(p)).get(workload, {}).get(kernel_key)
    except Exception:
        return None


def make_inputs(B, seed):
    from isg_b200 import synth

    b = synth.make_batch(B, channels=CHANNELS, num_ins=LAYERS, seed=seed)
    return b


def make_inputs_dp(B, world, rank, seed):
    """Data-parallel runs: ONE global batch of B x world graphs (same seed as the single-GPU batch), whole graphs
    dealt to the ranks by isg_b200.dp.balanced_shards (equal counts, near-equal node + edge totals) — a size-aware
    sampler in place of the reference's random DistributedSampler, so that no rank is the step's straggler."""
    from isg_b200 import synth
    from isg_b200.dp import balanced_shards

    topo = synth.make_topology(B * world, seed=seed)
    # per-graph cost ~ 17 ns per edge + 86 ns per node and layer (E-sized projections and edge kernels; node-sized
    # projections and the per-node kernels): weight a node as five edges
    sizes = (topo["num_edges"] + 5 * topo["num_nodes"]).tolist()
    ids = balanced_shards(sizes, world)[rank]
    return synth.make_batch_from_topology(synth.subset_topology(topo, ids), channels=CHANNELS, num_ins=LAYERS,
                                          seed=seed + rank)


# --------------------------------------------------------------------------------------------- CPU arm
def run_cpu_port(sampler, train, B, steps, warmup, seed=3407):
    """oracle/isg_oracle.py::OracleMGAT (CPU restatement of the reference) on all host threads."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import isg_oracle as O
    from isg_b200 import synth

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    b = make_inputs(B, seed)
    model = O.OracleMGAT(channels=CHANNELS, sampler_type=sampler, sample_k=K_SAMPLE)
    model.load_state_dict(synth.make_state_dict(CHANNELS, HEADS, LAYERS, seed))
    model.train(train)
    for st in model.aimle_state:
        st.beta = 1.0
    if sampler in ("imle", "aimle"):
        noise = synth.gumbel_noise(B, b["nmax"], 0.3, seed)
    else:
        noise = synth.gumbel_noise(B, b["nmax"], 1.0, seed)[:, 0, :, 0].contiguous()
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        if train:
            x = b["x"].clone().requires_grad_(True)
            ea = b["edge_attr"].clone().requires_grad_(True)
            model.zero_grad()
            h, _, _, _ = model(x, b["edge_index"], b["instr_vectors"], b["global_language_feats"], ea, b["batch"],
                               noise=noise)
            (h * h).mean().backward()
        else:
            with torch.no_grad():
                model(b["x"], b["edge_index"], b["instr_vectors"], b["global_language_feats"], b["edge_attr"],
                      b["batch"], noise=noise)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": B * len(times) / total, "ms_per_step": 1e3 * total / len(times), "cores": cores,
            "N": int(b["x"].shape[0]), "E": int(b["edge_index"].shape[1])}


def run_cpu_reference(sampler, train, B, steps, warmup, seed=3407):
    """The reference's OWN MGAT (models/mgat.py + mgat_v2_conv.py + masking.py + sampling/**, unmodified files
    staged into the git-ignored oracle/_ref/ by oracle/stage_reference.py, or /root/reference where it exists),
    executed on the host cores on the pure-torch shim of its missing PyG wheels.  Same synthetic batch, weights
    and injected noise as the GPU arm.  Returns None when no reference tree is available (-> the port)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import reference_loader as rl
    from isg_b200 import synth

    if not rl.available():
        return None
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # the reference prints "using aimle as sample method" from its constructors: keep stdout for the ONE JSON line
    with rl.scratch_cwd(), contextlib.redirect_stdout(sys.stderr):
        rl.load()
        from ISubGVQA.models.mgat import MGAT as RefMGAT

        model = RefMGAT(channels=CHANNELS, num_ins=LAYERS, heads=HEADS, use_instr=True,
                        masking_thresholds=[1.0, 1.0, 1.0, 0.1], use_topk=True, interpretable_mode=False,
                        sampler_type=sampler, sample_k=K_SAMPLE, nb_samples=1, alpha=1.0, beta=10.0, tau=1.0)
    model.load_state_dict(synth.make_state_dict(CHANNELS, HEADS, LAYERS, seed))
    model.train(train)
    if sampler == "aimle":  # same warmed-up beta as the GPU arm (the state object lives in the decorator's closure)
        for cell in model.convs[3].mask.sampler_train.__closure__:
            if type(cell.cell_contents).__name__ == "AdaptiveTargetDistribution":
                cell.cell_contents.beta = 1.0
    b = make_inputs(B, seed)
    if sampler in ("imle", "aimle"):
        noise = synth.gumbel_noise(B, b["nmax"], 0.3, seed)
    else:
        noise = synth.gumbel_noise(B, b["nmax"], 1.0, seed)[:, 0, :, 0].contiguous()
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        ctx = rl.inject_noise(noise) if sampler in ("imle", "aimle") else rl.inject_device_gumbel(noise)
        with rl.scratch_cwd(), ctx:
            if train:
                x = b["x"].clone().requires_grad_(True)
                ea = b["edge_attr"].clone().requires_grad_(True)
                model.zero_grad()
                h, _, _, _ = model(x, b["edge_index"], b["instr_vectors"], b["global_language_feats"], ea,
                                   b["batch"], return_masks=True)
                (h * h).mean().backward()
            else:
                with torch.no_grad():
                    model(b["x"], b["edge_index"], b["instr_vectors"], b["global_language_feats"], b["edge_attr"],
                          b["batch"], return_masks=True)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": B * len(times) / total, "ms_per_step": 1e3 * total / len(times), "cores": cores,
            "N": int(b["x"].shape[0]), "E": int(b["edge_index"].shape[1]), "nmax": int(b["nmax"]),
            "root": rl.REFERENCE_ROOT}


def run_cpu_arm(sampler, train, B, steps, warmup):
    """-> (result, kind, description): the reference's own code when its files are available, else the port."""
    r = run_cpu_reference(sampler, train, B, steps, warmup)
    if r is not None:
        return r, "reference", ("the reference's own MGAT (unmodified models/*.py + sampling/**, staged by "
                                "oracle/stage_reference.py) on the pure-torch shim of its PyG wheels")
    return run_cpu_port(sampler, train, B, steps, warmup), "port", "oracle/isg_oracle.py OracleMGAT (CPU port)"


def main_reference(args, rank, world):
    desc, sampler, train, B = WORKLOADS[args.workload]
    if rank != 0:
        return
    r, kind, what = run_cpu_arm(sampler, train, B, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "graphs_per_gpu": B, "nodes": r["N"], "edges": r["E"],
                   "nmax": r.get("nmax"), "channels": CHANNELS, "heads": HEADS, "layers": LAYERS, "sample_k": K_SAMPLE,
                   "step": "MGAT forward+backward" if train else "MGAT forward (no_grad)",
                   "sample": f"the full {B}-graph batch of the workload per step, CPU only"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": kind,
                         "sample": f"{what}; {B}-graph batches (N={r['N']}, E={r['E']}) of the same workload, "
                                   f"{args.steps} steps after {args.warmup} warm-up, all host threads"},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- GPU arm
def main_isg(args, rank, world, local_rank):
    import torch.distributed as dist

    import isg_b200  # noqa: F401
    from isg_b200 import lib as L
    from isg_b200 import synth
    from isg_b200.dp import GradAllReduce
    from isg_b200.graph import clear_cache
    from isg_b200.isubgvqa import MGAT

    L.load()  # fails loudly if libisg.so is missing — there is no fallback path
    from isg_b200 import ops

    ops.set_gemm_mode(args.gemm_mode)
    gemm_desc = {0: "fp32 FFMA (mode 0)", 1: "tcgen05 3xTF32 split, fp32-grade (mode 1)",
                 2: "tcgen05 single-pass TF32 (mode 2, NOT the parity configuration)",
                 3: "bf16 configuration (mode 3): tcgen05 kind::f16 projections, bf16 storage of x_l|x_r / e_proj / out "
                    "and their gradients, fp32 accumulation, fp32 gate logits / SDPA / GraphNorm; NOT the parity "
                    "configuration (tests/test_bf16_gpu.py states its tolerances)"}[args.gemm_mode]
    bf16 = args.gemm_mode == 3
    desc, sampler, train, B = WORKLOADS[args.workload]
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    out_stream = sys.stdout
    if world > 1:
        # NCCL writes its version banner to fd 1; the contract is ONE JSON line on stdout, so the process's
        # fd 1 is pointed at stderr for the run and the JSON line goes to a private copy of the real stdout
        sys.stdout.flush()
        out_stream = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
    seed = 3407 + rank
    b = make_inputs(B, seed) if world == 1 else make_inputs_dp(B, world, rank, 3407)
    N, E, nmax = int(b["x"].shape[0]), int(b["edge_index"].shape[1]), b["nmax"]
    model = MGAT(channels=CHANNELS, num_ins=LAYERS, heads=HEADS, use_instr=True,
                 masking_thresholds=[1.0, 1.0, 1.0, 0.1], use_topk=True, interpretable_mode=False,
                 sampler_type=sampler, sample_k=K_SAMPLE, nb_samples=1, alpha=1.0, beta=10.0, tau=1.0)
    model.load_state_dict(synth.make_state_dict(CHANNELS, HEADS, LAYERS, 3407))  # same weights on every rank
    model.to(dev).train(train)
    if sampler == "aimle":
        model.convs[3].mask.sampler_train.target.beta = 1.0  # warmed-up beta (beta0 = 0 gives zero grads)
    # Multi-GPU training: one flat all-reduce after each replay of the captured step (default).  --overlap buckets
    # the collective, issues it from gradient hooks during the backward pass and captures it into the step's CUDA
    # graph (isg_b200.dp.OverlappedGradAllReduce; the process group must then exist before the capture).  Measured
    # on 2 GPUs (r1i): 5.47-5.48 vs 5.43 ms/step — the bucket collectives compete with the persistent GEMM CTAs for
    # SMs and nothing is gained at this scale, so it stays opt-in (8 GPUs not measured).
    # Reducers (isg_b200.dp), measured on 2 and 8 B200 in r2 (profiles/r2_bench_8gpu*.json):
    #   flat (default)   GradAllReduce: pack -> one 42 MB NCCL AVG all-reduce -> unpack, issued after each replay
    #                    8 GPUs 5.35 ms/step, 2 GPUs 5.24
    #   --dp-layer       LayerGradAllReduce: .grad tensors ARE views of the executor's flat gradient buffer (no pack /
    #                    unpack), one collective after the backward pass, captured in the step graph: 5.42 / 5.33
    #   --dp-layer --dp-overlap   the same bucket, one slice per layer issued on a communication stream as soon as that
    #                    layer's backward has been issued: 5.36 / 5.29 (eager e2e suffers: the collectives queue behind
    #                    the persistent GEMM CTAs and couple the ranks' timing)
    #   --overlap        r1's hook-based bucketing (OverlappedGradAllReduce)
    # All within 1.5 %: the collective itself is 0.10 ms at 2 GPUs (scripts/allreduce_probe.py) and the rest of the
    # 1 -> 8 GPU gap is the max over ranks of differently sized batches, which no reducer removes.
    dp_mode = "none"
    if world > 1 and train:
        dp_mode = "hooks" if (args.overlap and not args.no_graph) else ("layer" if (args.dp_layer or args.dp_overlap) else "flat")
    overlap = dp_mode in ("hooks", "layer")  # the reduction is part of step_core (and of the captured graph)
    reducer = None
    if overlap:
        from isg_b200.dp import LayerGradAllReduce, OverlappedGradAllReduce

        dist.init_process_group("nccl", device_id=dev)
        reducer = (OverlappedGradAllReduce(model) if dp_mode == "hooks"
                   else LayerGradAllReduce(model, overlap=args.dp_overlap))

    keys = ("x", "edge_index", "instr_vectors", "global_language_feats", "edge_attr", "batch")
    host = {k: b[k].pin_memory() for k in keys}
    if sampler in ("imle", "aimle"):
        noise_h = synth.gumbel_noise(B, nmax, 0.3, seed).pin_memory()
    else:
        noise_h = synth.gumbel_noise(B, nmax, 1.0, seed)[:, 0, :, 0].contiguous().pin_memory()
    resident = {k: host[k].to(dev) for k in keys}
    noise_d = noise_h.to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    param_list = list(model.parameters())

    def step_core(t, noise):
        """MGAT forward (+ backward): the part that is captured into the CUDA graph."""
        model.convs[3].mask.injected_noise = noise
        if train:
            x = t["x"].detach().requires_grad_(True)
            ea = t["edge_attr"].detach().requires_grad_(True)
            for p in param_list:
                p.grad = None
            h, mask, _, _ = model(x, t["edge_index"], t["instr_vectors"], t["global_language_feats"], ea, t["batch"],
                                  return_masks=True)
            loss = (h * h).mean()
            loss.backward()
            if overlap:
                reducer.finish()  # joins the bucket all-reduces issued by the gradient hooks during backward()
            return loss, mask
        with torch.no_grad():
            h, mask, _, _ = model(t["x"], t["edge_index"], t["instr_vectors"], t["global_language_feats"],
                                  t["edge_attr"], t["batch"], return_masks=True)
            return (h * h).mean(), mask

    def step(t, noise):
        out = step_core(t, noise)
        if reducer is not None and not overlap:
            reducer.all_reduce_mean()
        return out

    from isg_b200.loader import DevicePrefetcher

    prefetcher = DevicePrefetcher(dev)
    pending = []

    def stage_next():
        # a new batch arrives as new device tensors: its CSR / graph_ptr are built afresh (on the copy stream)
        pending.append(prefetcher.stage(host, extra={"noise": noise_h}, nmax=nmax))

    e2e_prof = {"stage_ms": 0.0, "enqueue_ms": 0.0, "wait_ms": 0.0, "n": 0}
    from isg_b200.loader import HostResults

    results = HostResults(lag=1)
    e2e_seen = []

    def step_e2e():
        # public-API path, every step: pinned host batch -> H2D + CSR build (copy stream, one batch ahead of the
        # compute) -> MGAT fwd+bwd -> D2H copy of the step's loss and node mask into pinned buffers; the host reads
        # them one step later (HostResults), so it enqueues step i+1 while the GPU still runs step i.
        t0 = time.perf_counter()
        if not pending:
            stage_next()
        t, ext = prefetcher.get(pending.pop(0))
        stage_next()
        t1 = time.perf_counter()
        loss, mask = step(t, ext["noise"])
        results.push(loss=loss, mask=mask)
        t2 = time.perf_counter()
        r = results.pop()
        if r is not None:
            e2e_seen.append(float(r["loss"]))
        t3 = time.perf_counter()
        e2e_prof["stage_ms"] += 1e3 * (t1 - t0)
        e2e_prof["enqueue_ms"] += 1e3 * (t2 - t1)
        e2e_prof["wait_ms"] += 1e3 * (t3 - t2)
        e2e_prof["n"] += 1

    def timed_e2e(steps):
        """K end-to-end steps under ONE event pair (no L2 flush: every step's inputs arrive by H2D into fresh
        buffers and its 3 GB working set exceeds the L2), the last results drained inside the timed region."""
        barrier()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        a.record()
        for _ in range(steps):
            step_e2e()
        for r in results.drain():
            e2e_seen.append(float(r["loss"]))
        e.record()
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - w0)
        ms = max(a.elapsed_time(e), 0.0)
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, wall_ms

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    per_step_log = []  # per-step times of every timed() pass, in call order (ISG_BENCH_PER_STEP=1 prints them)

    def timed(fn, steps, timing_names=None):
        evs = []
        barrier()
        if timing_names:
            L.enable_timing(timing_names)
        launches0 = L.launch_count
        for _ in range(steps):
            flush.fill_(1)  # evict L2 between timed iterations (untimed)
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            e.record()
            evs.append((a, e))
        barrier()
        t = L.disable_timing() if timing_names else None
        ms = sum(a.elapsed_time(e) for a, e in evs)
        per_step_log.append([round(a.elapsed_time(e), 4) for a, e in evs])
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, L.launch_count - launches0, L.timing_summary(t) if t else {}

    # warm-up (also builds the CSR once for the resident batch, as a data loader would at collate time).  With
    # CUDA-graph capture every pre-capture step runs on the capture stream: autograd binds each parameter's
    # AccumulateGrad node to the stream it was first used on, and a mismatch with the capture stream makes the
    # engine synchronise with an uncaptured stream, which invalidates the capture.
    cap_stream = torch.cuda.Stream(device=dev) if not args.no_graph else None
    if cap_stream is not None:
        cap_stream.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(cap_stream) if cap_stream is not None else contextlib.nullcontext():
        for _ in range(max(args.warmup, 3)):
            step_core(resident, noise_d)
    torch.cuda.synchronize()
    # The resident-input step is launch-bound on the host (~180 kernel launches + autograd bookkeeping take
    # about as long as the GPU needs for them), so forward+backward is captured ONCE into a CUDA graph and the
    # timed region replays it (the NCCL all-reduce, if any, is issued eagerly after each replay): identical
    # kernels, no per-launch host cost.  The per-kernel CUDA-event timings for the roofline come from an eager
    # pass of the same K steps (events cannot be recorded inside a replay).  A failed capture leaves the CUDA
    # RNG in capture mode, so the process re-executes itself with --no-graph (before any process group exists).
    graph = None
    if not args.no_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=cap_stream):
                step_core(resident, noise_d)
            for _ in range(2):
                graph.replay()
            torch.cuda.synchronize()
        except Exception as exc:  # noqa: BLE001
            sys.stderr.write(f"[bench] CUDA graph capture failed ({type(exc).__name__}: {str(exc)[:200]}); "
                             "re-running with --no-graph\n")
            sys.stderr.flush()
            if world > 1:
                os.dup2(out_stream.fileno(), 1)
            os.execv(sys.executable, [sys.executable] + sys.argv + ["--no-graph"])

    if world > 1 and not overlap:
        dist.init_process_group("nccl", device_id=dev)
        if train:
            reducer = GradAllReduce(model)

    no_reduce = os.environ.get("ISG_BENCH_NO_REDUCE") == "1"  # diagnostics: what the gradient exchange costs

    def graphed_step():
        graph.replay()
        if reducer is not None and not overlap and not no_reduce:
            reducer.all_reduce_mean()

    from isg_b200.isubgvqa import mgat as mgat_mod

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    # (1) eager step through the layer executor (what a caller of the nn.Module gets with resident inputs).  The
    # warm-up ran on the capture stream, so the FIRST eager step on the current stream pays a one-off ~30 ms
    # (autograd re-binds its AccumulateGrad streams, the allocator grows this stream's pool —
    # profiles/r2c_per_step_power_cap.json): one untimed step takes it out of the quoted eager time.
    step(resident, noise_d)
    ms_eager, launches, _ = timed(lambda: step(resident, noise_d), args.steps)
    # (2) value: the same step replayed from the captured graph
    if graph is not None:
        ms, _, _ = timed(graphed_step, args.steps)
    else:
        ms = ms_eager
    clk = clocks.stop() if rank == 0 else None
    n_buckets = len(reducer._buckets) if dp_mode == "hooks" else 0
    layer_reducer = reducer if dp_mode == "layer" else None
    if overlap:
        # the per-operator timing pass below has no layer executor underneath (and the per-bucket hooks cost host
        # time in eager mode): it uses the flat reducer
        if dp_mode == "hooks":
            reducer.remove_hooks()
        else:
            reducer.detach()
        reducer = GradAllReduce(model)
        overlap = False
    # (3) per-kernel CUDA events for the roofline blocks: the same K steps through the per-operator path, where
    # every kernel family is its own C-ABI call that lib.call can bracket with events on the launching stream
    # (the executor issues them from C).  Same kernels, same order, same step.
    kernel_names = ["isg_gat_edge_fwd", "isg_gat_edge_bwd", "isg_linear_fwd", "isg_linear_dgrad", "isg_linear_wgrad"]
    tsum = {}
    if not bf16:  # (the bf16 configuration exists in the layer executor only: its edge kernels are timed alone below)
        mgat_mod.set_executor(False)
        for _ in range(2):
            step(resident, noise_d)
        _, _, tsum = timed(lambda: step(resident, noise_d), args.steps,
                           kernel_names if not args.breakdown else list(L.KERNELS_PER_CALL))
        mgat_mod.set_executor(True)
    tall = tsum
    if layer_reducer is not None:  # e2e: the layer reducer again (eager, same hooks)
        from isg_b200.dp import LayerGradAllReduce

        reducer = LayerGradAllReduce(model, overlap=args.dp_overlap)
        overlap = True
    for _ in range(3):
        step_e2e()
    results.drain()
    del e2e_seen[:]
    e2e_prof.update(stage_ms=0.0, enqueue_ms=0.0, wait_ms=0.0, n=0)
    ms_e2e, wall_e2e = timed_e2e(args.steps)
    if len(e2e_seen) != args.steps or not all(math.isfinite(v) for v in e2e_seen):
        raise RuntimeError(f"e2e: read {len(e2e_seen)} step results on the host, expected {args.steps}")

    def teardown():
        # a communicator whose collectives were captured into a CUDA graph must outlive that graph
        if world > 1:
            torch.cuda.synchronize()
            if graph is not None:
                graph.reset()
            dist.destroy_process_group()

    if rank != 0:
        teardown()
        return
    graphs = B * world * args.steps
    value = graphs / (ms / 1e3)
    e2e_value = graphs / (ms_e2e / 1e3)
    h2d = sum(host[k].numel() * host[k].element_size() for k in keys) + noise_h.numel() * 4
    d2h = 4 + N * 4
    peak, peak_src = peaks()
    fwd_b_un, bwd_b_un = edge_bytes(N, E, False)
    fwd_b_m, bwd_b_m = edge_bytes(N, E, True)
    roof = None
    if bf16 and world == 1:
        import importlib.util

        spec = importlib.util.spec_from_file_location("bench_edge", os.path.join(ROOT, "scripts", "bench_edge.py"))
        be = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(be)
        t = be.run_point(B, 20, 150, False, 20, B, seed=seed, bf16=True)
        key, kb = ("bwd_ms", "bwd_b") if train else ("fwd_ms", "fwd_b")
        ach = t[kb] / (t[key] * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "isg_gat_edge_%s, bf16 storage (head-pair ring kernels, 16-byte accesses), timed "
                                          "alone on a batch of the workload's shape, L2 flushed" % ("bwd" if train else "fwd"),
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": t[kb], "ms_per_launch": t[key],
                "launches_timed": 20}
    elif train and "isg_gat_edge_bwd" in tsum:
        calls, tot = tsum["isg_gat_edge_bwd"]
        per_launch_ms = tot / calls
        alg = (3 * bwd_b_un + bwd_b_m) / 4.0  # 3 unmasked layers + 1 masked layer per step
        ach = alg / (per_launch_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "isg_gat_edge_bwd (gat_edge_bwd_dst_ring + gat_edge_bwd_src_ring with the g_att fold as block roles; "
                                          "heavy-first task schedule)",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": measured_traffic("isg_gat_edge_bwd", args.workload),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "ms_per_launch": per_launch_ms,
                "launches_timed": calls}
    elif "isg_gat_edge_fwd" in tsum:
        calls, tot = tsum["isg_gat_edge_fwd"]
        per_launch_ms = tot / calls
        alg = (3 * fwd_b_un + fwd_b_m) / 4.0
        ach = alg / (per_launch_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "isg_gat_edge_fwd (gat_edge_fwd_kernel)", "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "traffic": measured_traffic("isg_gat_edge_fwd", args.workload),
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg, "ms_per_launch": per_launch_ms, "launches_timed": calls}
    edge_fwd = None
    if "isg_gat_edge_fwd" in tsum:
        calls, tot = tsum["isg_gat_edge_fwd"]
        alg = (3 * fwd_b_un + fwd_b_m) / 4.0
        edge_fwd = {"ms_per_launch": tot / calls, "achieved_GBps": alg / (tot / calls * 1e-3) / 1e9,
                    "frac": alg / (tot / calls * 1e-3) / 1e9 / peak}
    # projections: tensor-pipe roofline.  useful = algorithmic fp32 FLOPs / summed launch time of all isg_linear_*
    # calls; issued = 3x that in mode 1 (three TF32 products per fp32 product).  Peak: MEASURED_PEAKS.json has no TF32
    # entry; kind::tf32 runs at half the bf16 rate, so peak = bf16_tflops_sustained / 2 (kernels timed inside a step).
    roof_gemm = None
    lin = [tsum[k] for k in ("isg_linear_fwd", "isg_linear_dgrad", "isg_linear_wgrad") if k in tsum]
    if lin:
        lin_ms = sum(t for _, t in lin) / args.steps
        fl = gemm_flops(N, E, B, train)
        useful = fl / (lin_ms * 1e-3) / 1e12
        mult = {0: 1.0, 1: 3.0, 2: 1.0}[args.gemm_mode]
        pk_tf = None
        try:
            pk_tf = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]) / 2.0
        except Exception:
            pass
        roof_gemm = {"bound": "tensor", "kernel": "tc_gemm_kernel (all isg_linear_fwd/dgrad/wgrad launches of the step)",
                     "achieved": useful * mult, "achieved_useful_fp32": useful, "peak": pk_tf, "unit": "TFLOP/s",
                     "frac": (useful * mult / pk_tf) if pk_tf else None,
                     "frac_useful": (useful / pk_tf) if pk_tf else None,
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (kind::tf32 rate)",
                     "algorithmic_flops_per_step": fl, "ms_per_step": lin_ms,
                     "launches_per_step": sum(c for c, _ in lin) // args.steps,
                     "tensor_pipe_active_pct_ncu": measured_traffic("tc_gemm_tensor_pipe_pct", args.workload)}
    samp = None
    if world == 1 and sampler in ("imle", "aimle"):
        from isg_b200.graph import get_graph_index

        gi = get_graph_index(resident["edge_index"], resident["batch"], B)
        t_ms = sampler_study(N, E, B, nmax, gi, dev)
        sb = sampler_bytes(N, E, B, nmax)
        samp = {"kernel": "sampler_fused_fwd_kernel (gate dot + dropout + tau*noise + top-k + node mask + edge mask, "
                          "one launch)", "ms_per_launch": t_ms, "algorithmic_bytes_per_launch": sb,
                "achieved_GBps": sb / (t_ms * 1e-3) / 1e9, "frac": sb / (t_ms * 1e-3) / 1e9 / peak,
                "note": "latency-bound at this size (SURVEY.md section 8d: ~20 B per node): one CTA per graph, one warp "
                        "per node row for the gate dot products; 6.8 MB per launch"}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r, kind, what = run_cpu_arm(sampler, train, B, 2, 1)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": kind,
               "sample": f"{what}; the workload's own {B}-graph batch (N={r['N']}, E={r['E']}), 2 steps after 1 "
                         f"warm-up, {r['ms_per_step']:.0f} ms/step"}
    study = None
    if world == 1 and not args.no_edge_study:
        del resident, flush
        torch.cuda.empty_cache()
        study = edge_study()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if bf16 else "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "graphs_per_gpu": B, "nodes": N, "edges": E, "nmax": nmax,
                   "channels": CHANNELS, "heads": HEADS, "layers": LAYERS, "sample_k": K_SAMPLE,
                   "step": "MGAT forward+backward" + (
                       (" + NCCL gradient all-reduce (42 MB in %d buckets, issued by gradient hooks during backward, "
                        "captured in the step graph)" % n_buckets) if n_buckets else
                       (" + NCCL gradient all-reduce in place on the executor's flat gradient buffer (no pack/unpack), "
                        + ("one slice per layer issued during the backward pass" if args.dp_overlap else
                           "one 42 MB collective after the backward pass") + ", captured in the step graph")
                       if dp_mode == "layer" else
                       " + NCCL gradient all-reduce (42 MB flat bucket after the step)" if world > 1 else "")
                   if train else "MGAT forward (no_grad)",
                   "l2": "256 MiB buffer written between timed iterations (L2 flush); per-step working set "
                         f"~{(4 * 4 * HEADS * CHANNELS * (3 * E + 8 * N)) / 1e9:.2f} GB also exceeds the 126 MB L2",
                   "gemm_mode": gemm_desc,
                   "host_path": "layer executor: one C call per MGAT layer and direction (isg_mgat_layer_fwd/_bwd)",
                   "cuda_graph": ("value: K replays of one captured step; the same step issued eagerly through the "
                                  "nn.Module takes %.3f ms/step; e2e is eager" % (ms_eager / args.steps))
                   if graph is not None else "off",
                   "optimizer": "not in the timed step (SURVEY.md section 8 f4: isg_b200.optim)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps,
                "host_ms_per_step": {k: round(e2e_prof[k] / max(e2e_prof["n"], 1), 3)
                                     for k in ("stage_ms", "enqueue_ms", "wait_ms")},
                "host_wall_ms_per_step": round(wall_e2e / args.steps, 3),
                "what": "every step: pinned host batch -> H2D + CSR build on a copy stream one batch ahead (isg_b200."
                        "loader.DevicePrefetcher) -> MGAT fwd+bwd (eager, nn.Module API) -> D2H copy of loss + node "
                        "mask into pinned memory, read by the host one step later (isg_b200.loader.HostResults, lag 1); "
                        "all K results are on the host before the closing event; one event pair around the K steps"},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": roof,
        "roofline_gemm": roof_gemm,
        "sampler": samp,
        "edge_fwd": edge_fwd,
        "edge_roofline_study": study,
        "cpu_baseline": cpu,
    }
    if args.breakdown:
        line["breakdown_ms_per_step"] = {k: round(v[1] / args.steps, 4)
                                         for k, v in sorted(tall.items(), key=lambda kv: -kv[1][1])}
    if os.environ.get("ISG_BENCH_PER_STEP") == "1":  # diagnostics: [eager pass, replay pass, per-operator pass]
        line["per_step_ms"] = per_step_log
    out_stream.write(json.dumps(line) + "\n")
    out_stream.flush()
    teardown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="isg", choices=["isg", "reference"])
    ap.add_argument("--breakdown", action="store_true", help="add per-entry-point CUDA-event times to the JSON line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time the resident-input step eagerly (no CUDA graph)")
    ap.add_argument("--overlap", action="store_true",
                    help="multi-GPU (experimental): bucketed all-reduces issued during the backward pass and captured "
                         "in the step graph, instead of one flat all-reduce after the step")
    ap.add_argument("--dp-overlap", action="store_true",
                    help="multi-GPU: issue each layer's all-reduce during the backward pass (communication stream)")
    ap.add_argument("--dp-layer", action="store_true",
                    help="multi-GPU: the in-place reducer on the executor's flat gradient buffer (LayerGradAllReduce)")
    ap.add_argument("--dp-flat", action="store_true", help="(default) flat 42 MB all-reduce with pack/unpack after each replay")
    ap.add_argument("--no-edge-study", action="store_true", help="skip the batch-4096 edge-kernel roofline point")
    ap.add_argument("--gemm-mode", type=int, default=1, choices=[0, 1, 2, 3],
                    help="projection arithmetic: 0 fp32 FFMA, 1 tcgen05 3xTF32 (default), 2 tcgen05 1xTF32")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        main_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl isg needs a CUDA device (no CPU fallback)")
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    main_isg(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
