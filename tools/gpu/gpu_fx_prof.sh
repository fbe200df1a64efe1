#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:descend_fx -s 39 -c 1 -o gpurun_out/prof_fx -f python tools/profile_move.py c2 1 > gpurun_out/ncu_fx.log 2>&1
tail -2 gpurun_out/ncu_fx.log
timeout 600 python -m pytest tests/test_gpu_fx.py -m gpu -q -s --no-header --tb=short -k "tree_mode" 2>&1 | tail -30 > gpurun_out/pytest_fx2.log
grep -E "passed|failed|stored values|FAILED|Error|assert" gpurun_out/pytest_fx2.log | cut -c1-250 | head
