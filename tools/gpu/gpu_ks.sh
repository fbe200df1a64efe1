#!/bin/bash
mkdir -p gpurun_out
BL_DESCEND_VARIANT=3 BL_MW_LANES=4 timeout 400 python -m pytest tests/test_gpu_mcts.py -m gpu -q --maxfail=10 --no-header -rN --tb=short -k "stepwise" 2>&1 | tail -30 > gpurun_out/pytest_ks.log
grep -E "passed|failed" gpurun_out/pytest_ks.log | tail -2
grep -E "^(FAILED|ERROR)|^E  |^_{5,}" gpurun_out/pytest_ks.log | cut -c1-250 | head -20
for cfg in c5-11 c5-13 c3; do echo "$cfg mw: $(timeout 300 python tools/descend_time.py $cfg 2>&1 | tail -1)"; done
