#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_mcts.py tests/test_gpu_agent.py -m gpu -q --no-header -rN --tb=short -x 2>&1 | tail -3
timeout 300 python tools/descend_phases.py c2 2>&1 | grep -E "plain|adopt|trips"
timeout 300 python tools/descend_phases.py c3 2>&1 | grep -E "plain"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 | grep -o '"clocks": {[^}]*}'
