#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --maxfail=40 --no-header -rN --tb=short 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -3
grep -E "^(FAILED|ERROR)|^E  .*Error|^_{5,}" gpurun_out/pytest_gpu.log | cut -c1-200 | head -40
timeout 600 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_c2.log 2>&1
tail -c 3500 gpurun_out/bench_c2.log | grep -o '"value": [0-9.]*\|"ms_per_move_by_kernel": {[^}]*}' | head -3
K='regex:descend_v3|expand_step|gather_leaves|fc_tc|set_eval|backup_kernel|reset_kernel|root_kernel'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches.csv python tools/profile_move.py c2 1 > gpurun_out/ncu_launch.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv
