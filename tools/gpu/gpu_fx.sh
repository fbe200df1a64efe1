#!/bin/bash
# parity + timing of the certified fast descents (variants 5, 6)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mcts.py -m gpu -q -x --no-header -k "stepwise and (6- or 5-)" --tb=short 2>&1 | tail -30 > gpurun_out/pytest_fx_step.log
tail -5 gpurun_out/pytest_fx_step.log
timeout 1500 python -m pytest tests/test_gpu_fx.py -m gpu -q -s --no-header --tb=short -k "variants or t256" 2>&1 | tail -80 > gpurun_out/pytest_fx.log
grep -E "passed|failed|evaluations|stored values|more than|FAILED|Error|assert" gpurun_out/pytest_fx.log | cut -c1-300 | head -40
for c in 1 2 4; do
  echo "variant 6 ctas/SM $c: $(BL_DESCEND_VARIANT=6 BL_ALL_CTAS=$c timeout 300 python tools/descend_time.py c2 2>&1 | tail -1)"
done
echo "variant 6 c5-5: $(BL_DESCEND_VARIANT=6 timeout 300 python tools/descend_time.py c5-5 2>&1 | tail -1)"
echo "variant 6 c5-13: $(BL_DESCEND_VARIANT=6 timeout 300 python tools/descend_time.py c5-13 2>&1 | tail -1)"
if [ -n "$FX_PROF" ]; then
BL_DESCEND_VARIANT=6 timeout 900 ncu --set full --clock-control none --import-source on -k regex:eval_all -s 39 -c 1 -o gpurun_out/prof_all -f python tools/profile_move.py c2 1 > gpurun_out/ncu_all.log 2>&1
tail -2 gpurun_out/ncu_all.log
fi
