#!/bin/bash
# ncu --set full captures of the small kernels the north star names beside descend / backup: the env transition and the root kernel
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:hex_step_kernel -s 200 -c 1 -o gpurun_out/prof_hex_step -f python tools/profile_move.py c2 2 > gpurun_out/ncu_hex.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:root_kernel -s 1 -c 1 -o gpurun_out/prof_root -f python tools/profile_move.py c2 2 > gpurun_out/ncu_root.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:backup_kernel -s 100 -c 1 -o gpurun_out/prof_backup -f python tools/profile_move.py c2 2 > gpurun_out/ncu_backup.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:fc_tc_kernel -s 100 -c 1 -o gpurun_out/prof_fc_tc -f python tools/profile_move.py c2 2 > gpurun_out/ncu_fc.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
