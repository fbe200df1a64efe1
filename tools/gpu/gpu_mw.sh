#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_mcts.py -m gpu -q --maxfail=10 --no-header -rN --tb=short -k "stepwise" 2>&1 | tail -60 > gpurun_out/pytest_mw.log
grep -E "passed|failed" gpurun_out/pytest_mw.log | tail -3
grep -E "^(FAILED|ERROR)|^E  |^_{5,}" gpurun_out/pytest_mw.log | cut -c1-220 | head -40
for v in 3 2; do
BL_DESCEND_VARIANT=$v timeout 300 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_v$v.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/bench_v$v.log'):
    if l.startswith('{'):
        d = json.loads(l); print('variant $v', round(d['value']/1e6,1), 'M sims/s', d['ms_per_step'], d['roofline']['ms_per_move_by_kernel'])
PY
tail -3 gpurun_out/bench_v$v.log | cut -c1-300 | grep -v '^{'
done
