#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_mcts.py tests/test_gpu_agent.py -m gpu -q --no-header -rN --tb=short -x 2>&1 | tail -3
timeout 300 python tools/descend_phases.py c2 2>&1 | head -20
timeout 300 python tools/descend_phases.py c3 2>&1 | grep plain
