#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_net.py -m gpu -q --no-header -rN --tb=short -s -x 2>&1 | tail -60 > gpurun_out/pytest_net.log
cat gpurun_out/pytest_net.log | cut -c1-300 | tail -30
