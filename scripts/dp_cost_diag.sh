T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
ISG_BENCH_NO_REDUCE=1 timeout 200 $T 29541 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/r3l_2gpu_noreduce.json 2>/dev/null; echo rc=$?
timeout 200 $T 29542 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/r3l_2gpu_flat.json 2>/dev/null; echo rc=$?
timeout 200 $T 29543 bench.py --gpus 2 --steps 40 --warmup 5 --dp-layer > gpurun_out/r3l_2gpu_layer.json 2>/dev/null; echo rc=$?
NCCL_MAX_NCHANNELS=4 timeout 200 $T 29544 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/r3l_2gpu_flat_4ch.json 2>/dev/null; echo rc=$?
timeout 120 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-edge-study > gpurun_out/r3l_1gpu.json 2>/dev/null
