#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_arena.py -m gpu -q --maxfail=10 --no-header -rN --tb=short 2>&1 | tail -60 > gpurun_out/pytest_arena.log
grep -E "passed|failed" gpurun_out/pytest_arena.log | tail -3
grep -E "^(FAILED|ERROR)|^E  |^_{5,}" gpurun_out/pytest_arena.log | cut -c1-250 | head -40
