#!/bin/bash
# first GPU trip: parity tests, smoke, small + full bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -x --no-header -rN 2>&1 | tail -80 > gpurun_out/pytest_gpu.log
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 300 python bench.py --config c1 --steps 3 --warmup 3 > gpurun_out/bench_c1.log 2>&1
timeout 900 python bench.py --config c2 --steps 3 --warmup 3 > gpurun_out/bench_c2.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; tail -c 1500 gpurun_out/bench_c1.log; tail -c 3000 gpurun_out/bench_c2.log
