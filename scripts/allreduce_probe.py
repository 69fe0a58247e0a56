"""All-reduce latency probe (one process per GPU under torchrun): NCCL AVG all-reduce of the MGAT gradient bucket
(42 MB fp32) and of one layer slice (10.5 MB), eager and CUDA-graph replayed, next to torch symmetric-memory
variants when available.  Usage: torchrun --nproc-per-node N scripts/allreduce_probe.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
res = {"world": world, "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}}


def timeit(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t.item()) * 1e3, 1)


for name, n in (("layer_10MB", 2_630_000), ("bucket_42MB", 10_520_000)):
    buf = torch.randn(n, device=dev)
    res[name + "_nccl_us"] = timeit(lambda: dist.all_reduce(buf, op=dist.ReduceOp.AVG))
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        dist.all_reduce(buf, op=dist.ReduceOp.AVG)
    torch.cuda.synchronize()
    try:
        with torch.cuda.graph(g, stream=s):
            dist.all_reduce(buf, op=dist.ReduceOp.AVG)
        res[name + "_nccl_graph_us"] = timeit(g.replay)
    except Exception as exc:  # noqa: BLE001
        res[name + "_nccl_graph_us"] = f"capture failed: {type(exc).__name__}"
try:
    import torch.distributed._symmetric_memory as symm

    n = 10_520_000
    t = symm.empty(n, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
    t.normal_()
    res["symm_multicast"] = bool(getattr(hdl, "multicast_ptr", 0))
    for op in ("one_shot_all_reduce", "two_shot_all_reduce_", "multimem_all_reduce_"):
        try:
            fn = getattr(torch.ops.symm_mem, op)
            res["bucket_42MB_symm_" + op + "_us"] = timeit(lambda: fn(t, "sum", dist.group.WORLD.group_name))
        except Exception as exc:  # noqa: BLE001
            res["bucket_42MB_symm_" + op + "_us"] = f"{type(exc).__name__}: {str(exc)[:80]}"
except Exception as exc:  # noqa: BLE001
    res["symm"] = f"{type(exc).__name__}: {str(exc)[:120]}"
if rank == 0:
    print(json.dumps(res))
dist.barrier()
dist.destroy_process_group()
