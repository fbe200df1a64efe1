#!/bin/bash
mkdir -p gpurun_out
export BL_DESCEND_VARIANT=3
for L in 2 4; do
BL_MW_LANES=$L timeout 300 python -m pytest tests/test_gpu_mcts.py -m gpu -q --maxfail=10 --no-header -rN --tb=short -k "stepwise" 2>&1 | tail -40 > gpurun_out/pytest_mw_L$L.log
echo "L=$L: $(grep -E 'passed|failed' gpurun_out/pytest_mw_L$L.log | tail -1)"
grep -E "^(FAILED|ERROR)|^E  |^_{5,}" gpurun_out/pytest_mw_L$L.log | cut -c1-220 | head -20
for g in 1/2 1/3; do echo "L $L gate $g: $(BL_MW_LANES=$L BL_MW_GATE=$g timeout 200 python tools/descend_time.py c2 2>&1 | tail -1)"; done
done
echo "L 2 gate 2/3: $(BL_MW_LANES=2 BL_MW_GATE=2/3 timeout 200 python tools/descend_time.py c2 2>&1 | tail -1)"
BL_MW_LANES=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:descend_mw -s 40 -c 1 -o gpurun_out/prof_descend_mw_c -f python tools/profile_move.py c2 1 > gpurun_out/ncu_mw.log 2>&1
