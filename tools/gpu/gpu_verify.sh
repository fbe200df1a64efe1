#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 --no-header -rN --tb=short 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -3
grep -E "^(FAILED|ERROR)|^E  .*Error|^_{5,}" gpurun_out/pytest_gpu.log | cut -c1-200 | head -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_c2.log 2>&1
tail -c 3000 gpurun_out/bench_c2.log
