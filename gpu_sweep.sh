#!/bin/bash
mkdir -p gpurun_out
for g in 1/3 1/2 2/3 1/4 1/6 0/1; do
  echo "gate $g: $(BL_GATE=$g timeout 300 python tools/descend_phases.py c2 2>&1 | grep plain)"
done
