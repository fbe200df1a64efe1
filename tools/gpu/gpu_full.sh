#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 --no-header -rN --tb=short 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -3
grep -E "^(FAILED|ERROR)|^E  .*Error|^_{5,}" gpurun_out/pytest_gpu.log | cut -c1-200 | head -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_c2.log 2>&1
tail -c 3800 gpurun_out/bench_c2.log
K='regex:descend_v3|expand_step|gather_leaves|fc_tc|set_eval|backup_kernel|reset_kernel|root_kernel'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches.csv python tools/profile_move.py c2 1 > gpurun_out/ncu_launch.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:descend_v3 -s 40 -c 1 -o gpurun_out/prof_descend_r1c -f python tools/profile_move.py c2 1 > gpurun_out/ncu_descend.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fc_tc_kernel -s 40 -c 1 -o gpurun_out/prof_fc_tc_r1c -f python tools/profile_move.py c2 1 > gpurun_out/ncu_fc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:backup_kernel -s 40 -c 1 -o gpurun_out/prof_backup_r1c -f python tools/profile_move.py c2 1 > gpurun_out/ncu_backup.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
