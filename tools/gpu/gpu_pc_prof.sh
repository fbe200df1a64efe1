#!/bin/bash
mkdir -p gpurun_out
BL_DESCEND_VARIANT=4 timeout 400 ncu --set full --clock-control none --import-source on -k regex:descend_pc -s 40 -c 1 -o gpurun_out/prof_descend_pc -f python tools/profile_move.py c2 1 > gpurun_out/ncu_pc.log 2>&1
ls -la gpurun_out/prof_descend_pc.ncu-rep
