#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_c2.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/bench_c2.log'):
    if l.startswith('{'):
        d = json.loads(l); print('c2', round(d['value']/1e6,2), 'M sims/s; e2e', round(d['e2e']['value']/1e6,2), d['e2e']['ms_per_step'], 'vs', d['ms_per_step'], '; cpu', d['cpu_baseline']['value'], d['cpu_baseline']['seconds'], '; refcuda', d['reference_cuda']['value'], '; clocks', d['clocks'])
PY
tail -2 gpurun_out/bench_c2.log | grep -v '^{' | cut -c1-300
