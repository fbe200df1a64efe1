#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_agent.py tests/test_gpu_mcts.py -m gpu -q --maxfail=20 --no-header -rN --tb=short 2>&1 | tail -80 > gpurun_out/pytest_fuse.log
grep -E "passed|failed" gpurun_out/pytest_fuse.log | tail -3
grep -E "^(FAILED|ERROR)|^E  |^_{5,}" gpurun_out/pytest_fuse.log | cut -c1-250 | head -40
for f in 1 0; do
BL_FUSE_BACKUP=$f timeout 600 python bench.py --no-cpu > gpurun_out/bench_fuse$f.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/bench_fuse$f.log'):
    if l.startswith('{'):
        d = json.loads(l); print('fuse $f: c2', round(d['value']/1e6,1), 'M sims/s; e2e', round(d['e2e']['value']/1e6,1), d['ms_per_step'], d['roofline']['ms_per_move_by_kernel'], 'launches', d['gpu_launches'])
PY
tail -2 gpurun_out/bench_fuse$f.log | grep -v '^{' | cut -c1-300
done
