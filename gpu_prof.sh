#!/bin/bash
mkdir -p gpurun_out
./tools/ubench > gpurun_out/ubench.log 2>&1
cat gpurun_out/ubench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 500 --csv --log-file gpurun_out/launches.csv python profile_move.py c2 1 > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log; wc -l gpurun_out/launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fc_tc -s 40 -c 1 -o gpurun_out/prof_fc_tc -f python profile_move.py c2 1 > gpurun_out/ncu_fc.log 2>&1
tail -3 gpurun_out/ncu_fc.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:backup_kernel -s 40 -c 1 -o gpurun_out/prof_backup -f python profile_move.py c2 1 > gpurun_out/ncu_backup.log 2>&1
tail -3 gpurun_out/ncu_backup.log
