#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q --no-header -rN --tb=short -x -k "512" -s 2>&1 | grep "max |dlogit|" > gpurun_out/wide_err.log
cat gpurun_out/wide_err.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 --no-header -rN --tb=short 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -3
grep -E "^(FAILED|ERROR)|^E  .*Error|^_{5,}" gpurun_out/pytest_gpu.log | cut -c1-200 | head -40
timeout 900 python bench.py --config c3 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_c3.log 2>&1
echo "c3: $(tail -c 4000 gpurun_out/bench_c3.log | grep -o '"value": [0-9.]*' | head -1) $(tail -c 4000 gpurun_out/bench_c3.log | grep -o '"ms_per_move_by_kernel": {[^}]*}') $(grep -i error gpurun_out/bench_c3.log | tail -1 | cut -c1-200)"
timeout 600 python tools/descend_phases.py c3 2>&1 | tail -16
