#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_net.py tests/test_gpu_agent.py -m gpu -q --no-header -rN --tb=short -x -k "512" 2>&1 | tail -8
timeout 600 python bench.py --config c3 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_c3.log 2>&1
echo "c3: $(tail -c 4000 gpurun_out/bench_c3.log | grep -o '"value": [0-9.]*' | head -1) $(tail -c 4000 gpurun_out/bench_c3.log | grep -o '"ms_per_move_by_kernel": {[^}]*}') $(grep -i error gpurun_out/bench_c3.log | tail -1 | cut -c1-200)"
timeout 300 python tools/descend_phases.py c3 > gpurun_out/phases_c3.log 2>&1; tail -14 gpurun_out/phases_c3.log
