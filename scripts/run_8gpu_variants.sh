T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
timeout 240 $T 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2u_bench_8gpu.json 2> gpurun_out/r2u_8.err; echo rc=$?
timeout 240 $T 29522 bench.py --gpus 8 --steps 20 --warmup 5 --dp-flat > gpurun_out/r2u_bench_8gpu_flat.json 2> gpurun_out/r2u_8flat.err; echo rc=$?
timeout 240 $T 29523 bench.py --gpus 8 --steps 20 --warmup 5 --dp-overlap > gpurun_out/r2u_bench_8gpu_overlap.json 2> gpurun_out/r2u_8ov.err; echo rc=$?
