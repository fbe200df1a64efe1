#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 --no-header -rN --tb=short 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -3
grep -E "^(FAILED|ERROR)|^E  |^_{5,}" gpurun_out/pytest_gpu.log | cut -c1-220 | head -40
echo "c2 v3: $(BL_DESCEND_VARIANT=2 timeout 200 python tools/descend_time.py c2 2>&1 | tail -1)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_c2.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/bench_c2.log'):
    if l.startswith('{'):
        d = json.loads(l); print('c2', round(d['value']/1e6,1), 'M sims/s; e2e', round(d['e2e']['value']/1e6,1), d['roofline']['ms_per_move_by_kernel'], 'refcuda', (d.get('reference_cuda') or {}).get('value'))
PY
K='regex:descend_v3|descend_mw|expand_step|gather_leaves|fc_tc|set_eval|backup_kernel|reset_kernel|root_kernel'
for cfg in c2 c3; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 1200 --csv --log-file gpurun_out/launches_$cfg.csv python tools/profile_move.py $cfg 1 > gpurun_out/ncu_launch_$cfg.log 2>&1
echo "== $cfg"; python tools/launch_summary.py gpurun_out/launches_$cfg.csv | tee gpurun_out/launches_$cfg.txt | tail -12
done
