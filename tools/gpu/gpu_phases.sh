#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/descend_phases.py c2 > gpurun_out/phases.log 2>&1
cat gpurun_out/phases.log | tail -20
