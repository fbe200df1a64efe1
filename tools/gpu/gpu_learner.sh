#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_learner.py -m gpu -q --maxfail=20 --no-header -rN --tb=short 2>&1 | tail -80 > gpurun_out/pytest_learner.log
grep -E "passed|failed" gpurun_out/pytest_learner.log | tail -3
grep -E "^(FAILED|ERROR)|^E  |^_{5,}" gpurun_out/pytest_learner.log | cut -c1-250 | head -40
timeout 900 python bench.py > gpurun_out/bench_c2.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/bench_c2.log'):
    if l.startswith('{'):
        d = json.loads(l); print('c2', round(d['value']/1e6,1), 'M sims/s; e2e', round(d['e2e']['value']/1e6,1), '; cpu', d['cpu_baseline']['value'], '; refcuda', d.get('reference_cuda'))
PY
tail -3 gpurun_out/bench_c2.log | grep -v '^{' | cut -c1-300
