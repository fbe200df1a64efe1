#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_learner.py tests/test_gpu_agent.py tests/test_gpu_net.py -m gpu -q --maxfail=40 --no-header -rN --tb=short 2>&1 | tail -60 > gpurun_out/pytest_fix2.log
grep -E "passed|failed" gpurun_out/pytest_fix2.log | tail -3
grep -E "^(FAILED|ERROR)|^E  |^_{5,}" gpurun_out/pytest_fix2.log | cut -c1-260 | head -30
