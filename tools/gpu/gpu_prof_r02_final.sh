#!/bin/bash
# round-2 final evidence: ncu --set full of the kernels at HEAD (descent, network, backup, learner GEMM), the launch list of one move,
# and — from an experimental build made on the box — the packed descent (variant 7): its counters and an ncu capture
mkdir -p gpurun_out
N="--set full --clock-control none --import-source on"
timeout 600 ncu $N -k regex:descend_v3 -s 39 -c 1 -o gpurun_out/prof_descend_v3 -f python tools/profile_move.py c2 1 > gpurun_out/ncu_v3.log 2>&1; tail -1 gpurun_out/ncu_v3.log
timeout 600 ncu $N -k regex:fc_tc_kernel -s 40 -c 1 -o gpurun_out/prof_fc_tc -f python tools/profile_move.py c2 1 > gpurun_out/ncu_fc.log 2>&1; tail -1 gpurun_out/ncu_fc.log
timeout 600 ncu $N -k regex:backup_kernel -s 39 -c 1 -o gpurun_out/prof_backup -f python tools/profile_move.py c2 1 > gpurun_out/ncu_backup.log 2>&1; tail -1 gpurun_out/ncu_backup.log
timeout 600 ncu $N -k regex:gemm_tc_kernel -s 3 -c 1 -o gpurun_out/prof_gemm -f python tools/gemm_time.py > gpurun_out/ncu_gemm.log 2>&1; tail -1 gpurun_out/ncu_gemm.log
K='regex:descend_v3|expand_step|fc_tc|set_eval|backup_kernel|reset_kernel|root_kernel|hex_'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_c2.csv python tools/profile_move.py c2 1 > gpurun_out/ncu_launch_c2.log 2>&1
python tools/launch_summary.py gpurun_out/launches_c2.csv | tee gpurun_out/launches_c2.txt | tail -12
# experimental build (variants 4, 6, 7) on the box only
cp boardlaw_b200/libboardlaw_b200.so /tmp/lib_product.so
BL_EXPERIMENTAL=1 timeout 900 python -m boardlaw_b200.build -f > gpurun_out/build_exp.log 2>&1; tail -1 gpurun_out/build_exp.log
for p in 4 6; do BL_PK_PASS=$p timeout 200 python tools/pk_phases.py c2 2>&1 | tail -6; done | tee gpurun_out/pk_phases.txt
BL_DESCEND_VARIANT=7 timeout 600 ncu $N -k regex:descend_pk -s 39 -c 1 -o gpurun_out/prof_descend_pk -f python tools/profile_move.py c2 1 > gpurun_out/ncu_pk.log 2>&1; tail -1 gpurun_out/ncu_pk.log
timeout 600 python -m pytest tests/test_gpu_mcts.py tests/test_gpu_fx.py -m gpu -q --no-header --tb=short -k "(stepwise and 7-) or variants" 2>&1 | tail -3 | tee gpurun_out/pytest_pk.log
ls -la gpurun_out/*.ncu-rep | tail -8
