#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q --no-header -rN --tb=short -x -k "512" 2>&1 | tail -30 > gpurun_out/pytest_wide.log
cat gpurun_out/pytest_wide.log | cut -c1-300 | tail -30
