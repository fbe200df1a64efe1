#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mcts.py tests/test_gpu_agent.py -m gpu -q --no-header -rN --tb=short -x 2>&1 | tail -3
for e in 32 16 8; do
BL_BACKUP_ENVS=$e timeout 600 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_c2.log 2>&1
echo "envs/CTA $e: $(tail -c 3500 gpurun_out/bench_c2.log | grep -o '"value": [0-9.]*\|"ms_per_move_by_kernel": {[^}]*}' | head -3 | tr '\n' ' ')"
done
