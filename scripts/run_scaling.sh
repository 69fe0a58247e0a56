#!/bin/bash
# weak-scaling record with the default reducer: N = $1 GPUs (run under gpurun --gpus N)
N=$1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
timeout 240 $T 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r3k_bench_${N}gpu.json 2> gpurun_out/r3k_${N}.err; echo rc=$?
timeout 240 $T 29532 bench.py --gpus $N --steps 20 --warmup 5 --gemm-mode 3 > gpurun_out/r3k_bench_${N}gpu_bf16.json 2>> gpurun_out/r3k_${N}.err; echo rc=$?
if [ "$N" = "8" ]; then timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-edge-study > gpurun_out/r3k_bench_1gpu_samebox.json 2>/dev/null; fi
