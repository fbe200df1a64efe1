#!/bin/bash
# selected GPU tests + a short bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --no-header -rN --tb=short -k "${TESTS}" 2>&1 | tail -100 > gpurun_out/pytest_quick.log
grep -E "passed|failed" gpurun_out/pytest_quick.log | tail -3
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/pytest_quick.log | cut -c1-260 | head -40
if [ -n "$BENCH" ]; then
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_quick.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/bench_quick.log'):
    if l.startswith('{'):
        d = json.loads(l)
        print('c2', round(d['value'] / 1e6, 2), 'M sims/s; e2e', round(d['e2e']['value'] / 1e6, 2), '; ms', d['roofline']['ms_per_move_by_kernel'], 'ms/step', d['ms_per_step'], 'launches', d['gpu_launches'])
PY
tail -3 gpurun_out/bench_quick.log | grep -v '^{' | cut -c1-300
fi
