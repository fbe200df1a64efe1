#!/bin/bash
mkdir -p gpurun_out
for cfg in c5-5 c5-7 c5-11 c5-13 c3; do
echo "$cfg v3:      $(BL_DESCEND_VARIANT=2 timeout 300 python tools/descend_time.py $cfg 2>&1 | tail -1)"
echo "$cfg mw L=2:  $(BL_DESCEND_VARIANT=3 BL_MW_LANES=2 timeout 300 python tools/descend_time.py $cfg 2>&1 | tail -1)"
echo "$cfg mw L=4:  $(BL_DESCEND_VARIANT=3 BL_MW_LANES=4 timeout 300 python tools/descend_time.py $cfg 2>&1 | tail -1)"
done
