#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_net.py -m gpu -q --no-header -rN --tb=short -x 2>&1 | tail -5 > gpurun_out/pytest_net.log
cat gpurun_out/pytest_net.log | cut -c1-300 | tail -8
timeout 300 python tools/descend_phases.py c2 2>&1 | tail -11
