#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mcts.py -m gpu -q --maxfail=3 --no-header -rN --tb=short -k "stepwise and (False-4 or True-4)" 2>&1 | tail -40 > gpurun_out/pytest_pc.log
grep -E "passed|failed" gpurun_out/pytest_pc.log | tail -2
grep -E "^(FAILED|ERROR)|^E  |^_{5,}" gpurun_out/pytest_pc.log | cut -c1-250 | head -20
echo "c2 pc: $(BL_DESCEND_VARIANT=4 timeout 120 python tools/descend_time.py c2 2>&1 | tail -1)"
echo "c2 v3: $(BL_DESCEND_VARIANT=2 timeout 120 python tools/descend_time.py c2 2>&1 | tail -1)"
