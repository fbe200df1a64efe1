#!/bin/bash
mkdir -p gpurun_out
for g in 1/2 2/5 3/5; do echo "gate $g $(BL_GATE=$g timeout 300 python tools/descend_phases.py c2 2>&1 | grep -E "plain")"; done
