#!/bin/bash
# parity + timing of the packed descent (variant 7)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mcts.py -m gpu -q -x --no-header -k "stepwise and 7-" --tb=short 2>&1 | tail -30 > gpurun_out/pytest_pk_step.log
tail -5 gpurun_out/pytest_pk_step.log
if [ -n "$PK_FULL" ]; then
timeout 1200 python -m pytest tests/test_gpu_fx.py -m gpu -q -s --no-header --tb=short -k "variants" 2>&1 | tail -60 > gpurun_out/pytest_pk.log
grep -E "passed|failed|FAILED|Error|assert|differ" gpurun_out/pytest_pk.log | cut -c1-300 | head -20
fi
for p in ${PK_PASS:-4}; do
  echo "variant 7 pass warps $p: $(BL_DESCEND_VARIANT=7 BL_PK_PASS=$p timeout 300 python tools/descend_time.py c2 2>&1 | tail -1)"
done
echo "variant 2: $(BL_DESCEND_VARIANT=2 timeout 300 python tools/descend_time.py c2 2>&1 | tail -1)"
