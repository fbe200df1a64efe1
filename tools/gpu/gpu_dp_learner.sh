#!/bin/bash
# 2-GPU check of the data-parallel optimiser step (gpurun --gpus 2 -- bash tools/gpu/gpu_dp_learner.sh)
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_learner_check.py 2>&1 | grep -E "^world|Error|assert" | tee gpurun_out/dp_learner.log
