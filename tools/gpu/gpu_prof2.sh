#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:descend_v3 -s 40 -c 1 -o gpurun_out/prof_descend_r1b -f python tools/profile_move.py c2 1 > gpurun_out/ncu_descend.log 2>&1
tail -2 gpurun_out/ncu_descend.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fc_tc_wide -s 40 -c 1 -o gpurun_out/prof_fc_wide -f python tools/profile_move.py c3 1 > gpurun_out/ncu_wide.log 2>&1
tail -2 gpurun_out/ncu_wide.log
