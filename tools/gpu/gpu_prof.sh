#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fc_tc -s 30 -c 1 -o gpurun_out/prof_fc_tc -f python tools/profile_move.py c2 1 > gpurun_out/ncu_fc.log 2>&1
tail -2 gpurun_out/ncu_fc.log
