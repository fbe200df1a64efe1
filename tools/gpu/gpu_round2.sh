#!/bin/bash
# round-2 check: full GPU test suite, smoke, traffic capture at the current sources, the bench line (with extras), launch list
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --maxfail=30 --no-header -rN --tb=short 2>&1 | tail -120 > gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -3
grep -E "^(FAILED|ERROR)|^E  .*Error|^E  .*assert" gpurun_out/pytest_gpu.log | cut -c1-220 | head -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python tools/capture_traffic.py c2 > gpurun_out/traffic_c2.log 2>&1; tail -3 gpurun_out/traffic_c2.log | cut -c1-200
cp profiles/r02_traffic_c2.json gpurun_out/ 2>/dev/null
timeout 1500 python bench.py ${BENCH_ARGS} > gpurun_out/bench_c2.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/bench_c2.log'):
    if l.startswith('{'):
        d = json.loads(l)
        print('c2', round(d['value'] / 1e6, 2), 'M sims/s; e2e', round(d['e2e']['value'] / 1e6, 2), '; ms', d['roofline']['ms_per_move_by_kernel'])
        print({k: (round(v['frac'], 4), v.get('traffic')) for k, v in d['roofline']['by_kernel'].items()})
        print('cpu', d.get('cpu_baseline') and {k: d['cpu_baseline'].get(k) for k in ('value', 'seconds', 'python', 'unavailable')})
        print('refcuda', d.get('reference_cuda') and {k: d['reference_cuda'].get(k) for k in ('value', 'unavailable')})
        print('extra', {k: (round(v['value'] / 1e6, 2) if 'value' in v else v) for k, v in (d.get('extra') or {}).items()})
        print('clocks', d['clocks'])
PY
tail -3 gpurun_out/bench_c2.log | grep -v '^{' | cut -c1-300
