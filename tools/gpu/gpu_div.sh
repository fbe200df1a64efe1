#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mcts.py tests/test_gpu_agent.py -m gpu -q --maxfail=20 --no-header -rN --tb=short 2>&1 | tail -60 > gpurun_out/pytest_div.log
grep -E "passed|failed" gpurun_out/pytest_div.log | tail -3
grep -E "^(FAILED|ERROR)|^E  |^_{5,}" gpurun_out/pytest_div.log | cut -c1-250 | head -30
echo "c2 v3:     $(BL_DESCEND_VARIANT=2 timeout 200 python tools/descend_time.py c2 2>&1 | tail -1)"
echo "c2 mw L=2: $(BL_DESCEND_VARIANT=3 timeout 200 python tools/descend_time.py c2 2>&1 | tail -1)"
echo "c5-13 mw:  $(timeout 200 python tools/descend_time.py c5-13 2>&1 | tail -1)"
