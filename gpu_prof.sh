#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:descend_v2 -s 40 -c 1 -o gpurun_out/prof_descend_v2 -f python profile_move.py c2 1 > gpurun_out/ncu_descend.log 2>&1
tail -5 gpurun_out/ncu_descend.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 700 --csv --log-file gpurun_out/launches.csv python profile_move.py c2 1 > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log; wc -l gpurun_out/launches.csv
