#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus8.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_8gpu.log 2>&1
tail -c 1200 gpurun_out/bench_8gpu.log | head -c 1200; grep -o '"value": [0-9.]*' gpurun_out/bench_8gpu.log | head -2
