#!/bin/bash
# parity + timing of the certified fast descent (variant 5)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mcts.py -m gpu -q -x --no-header -k "stepwise and 5-" --tb=short 2>&1 | tail -30 > gpurun_out/pytest_fx_step.log
tail -5 gpurun_out/pytest_fx_step.log
timeout 1500 python -m pytest tests/test_gpu_fx.py -m gpu -q -s --no-header --tb=short 2>&1 | tail -80 > gpurun_out/pytest_fx.log
grep -E "passed|failed|evaluations|stored values|more than|FAILED|Error|assert" gpurun_out/pytest_fx.log | cut -c1-300 | head -40
for n in 4 8 16 32; do
  echo "variant 5 epw $n: $(BL_DESCEND_VARIANT=5 BL_FX_EPW=$n timeout 300 python tools/descend_time.py c2 2>&1 | tail -1)"
done
echo "variant 5 c3: $(BL_DESCEND_VARIANT=5 timeout 300 python tools/descend_time.py c3 2>&1 | tail -1)"
if [ -n "$FX_PROF" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:descend_fx -s 39 -c 1 -o gpurun_out/prof_fx -f python tools/profile_move.py c2 1 > gpurun_out/ncu_fx.log 2>&1
tail -2 gpurun_out/ncu_fx.log
fi
