#!/bin/bash
mkdir -p gpurun_out
export BL_DESCEND_VARIANT=3
for g in 0/1 1/4 1/2 3/4 1/1; do echo "gate $g fuse 1: $(BL_MW_GATE=$g timeout 200 python tools/descend_time.py c2 2>&1 | tail -1)"; done
echo "gate 1/2 fuse 0: $(BL_MW_FUSE=0 timeout 200 python tools/descend_time.py c2 2>&1 | tail -1)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:descend_mw -s 40 -c 1 -o gpurun_out/prof_descend_mw_a -f python tools/profile_move.py c2 1 > gpurun_out/ncu_mw.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
