#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 --no-header -rN --tb=short 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -3
grep -E "^(FAILED|ERROR)|^E  .*Error|^_{5,}" gpurun_out/pytest_gpu.log | cut -c1-200 | head -40
for cfg in c3 c5-13 c5-11; do
timeout 600 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_$cfg.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/bench_$cfg.log'):
    if l.startswith('{'):
        d = json.loads(l); print('$cfg', round(d['value']/1e6,2), 'M sims/s', round(d['ms_per_step'],1), 'ms/move', d['roofline']['ms_per_move_by_kernel'])
PY
done
