#!/bin/bash
mkdir -p gpurun_out
BL_FX_EPW=${EPW:-32} timeout 900 ncu --set full --clock-control none --import-source on -k regex:descend_fx -s 39 -c 1 -o gpurun_out/prof_fx -f python tools/profile_move.py c2 1 > gpurun_out/ncu_fx.log 2>&1
tail -2 gpurun_out/ncu_fx.log
