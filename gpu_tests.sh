#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --maxfail=40 --no-header -rN --tb=short 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
