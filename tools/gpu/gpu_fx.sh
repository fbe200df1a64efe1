#!/bin/bash
# parity + timing of the certified fast descent (variant 5)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mcts.py -m gpu -q -x --no-header -k "stepwise and 5-" --tb=short 2>&1 | tail -30 > gpurun_out/pytest_fx_step.log
tail -5 gpurun_out/pytest_fx_step.log
timeout 1500 python -m pytest tests/test_gpu_fx.py -m gpu -q -s --no-header --tb=short 2>&1 | tail -80 > gpurun_out/pytest_fx.log
grep -E "passed|failed|evaluations|identical|FAILED|Error|assert" gpurun_out/pytest_fx.log | cut -c1-250 | head -40
for v in 2 5; do for l in 4 8; do
  if [ $v = 2 ] && [ $l = 8 ]; then continue; fi
  echo "variant $v lanes $l: $(BL_DESCEND_VARIANT=$v BL_FX_LANES=$l timeout 300 python tools/descend_time.py c2 2>&1 | tail -1)"
done; done
timeout 600 python bench.py --config c2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c2_fx.log 2>&1
tail -c 4000 gpurun_out/bench_c2_fx.log | grep -o '"value": [0-9.]*\|"ms_per_move_by_kernel": {[^}]*}' | head -3
