#!/bin/bash
mkdir -p gpurun_out
for w in 1024 888 740 592 444; do
 for g in 1/2 1/3 1/4; do
  echo "warps $w gate $g: $(BL_DESCEND_GRID=$w BL_GATE=$g timeout 300 python tools/descend_phases.py c2 2>&1 | grep plain)"
 done
done
