#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mcts.py tests/test_gpu_agent.py -m gpu -q --no-header -rN --tb=short -x 2>&1 | tail -5
timeout 300 python tools/descend_phases.py c2 2>&1 | grep -E "plain|visit|pass  "
timeout 300 python tools/descend_phases.py c3 2>&1 | grep -E "plain|visit|pass  |child"
timeout 300 python tools/descend_phases.py c5-13 2>&1 | grep -E "plain"
