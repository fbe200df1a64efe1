#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 --no-header -rN --tb=short 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -3
grep -E "^(FAILED|ERROR)|^E  |^_{5,}" gpurun_out/pytest_gpu.log | cut -c1-220 | head -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_c2.log 2>&1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
tail -1 gpurun_out/bench_ref.log | cut -c1-400
for cfg in c5-5 c5-7 c5-9 c5-11 c5-13 c1; do
timeout 600 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_$cfg.log 2>&1
done
python - <<'PY'
import json, glob
for f in ['c2', 'c1', 'c5-5', 'c5-7', 'c5-9', 'c5-11', 'c5-13']:
    for l in open(f'gpurun_out/bench_{f}.log'):
        if l.startswith('{'):
            d = json.loads(l); r = d['roofline']
            sims = d['config']['envs_per_gpu'] * d['config']['n_nodes']
            tot = sum(r['algorithmic_bytes_per_move'].values())
            print(f, round(d['value']/1e6, 2), 'M sims/s; e2e', round(d['e2e']['value']/1e6, 2), '; ms', r['ms_per_move_by_kernel'], '; bytes/sim', round(tot / sims), '; dominant', r['kernel'], round(r['achieved'], 1), r['unit'], 'frac', round(r['frac'], 4), '; whole-move GB/s', round(tot / d['ms_per_step'] / 1e6, 1), '; tree', {k: round(v, 2) for k, v in r['tree_shape'].items()}, '; refcuda', (d.get('reference_cuda') or {}).get('value'), '; cpu', (d.get('cpu_baseline') or {}).get('value'))
PY
K='regex:descend_v3|descend_mw|expand_step|gather_leaves|fc_tc|set_eval|backup_kernel|reset_kernel|root_kernel'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 1200 --csv --log-file gpurun_out/launches_c2.csv python tools/profile_move.py c2 1 > gpurun_out/ncu_launch_c2.log 2>&1
python tools/launch_summary.py gpurun_out/launches_c2.csv | tee gpurun_out/launches_c2.txt | tail -10
timeout 600 ncu --set full --clock-control none --import-source on -k regex:descend_mw -s 100 -c 1 -o gpurun_out/prof_descend_mw_c3 -f python tools/profile_move.py c3 1 > gpurun_out/ncu_mw_c3.log 2>&1
ls -la gpurun_out/prof_descend_mw_c3.ncu-rep
