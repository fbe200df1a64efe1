#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_net.py -m gpu -q --no-header -rN --tb=short -x -k amp -s 2>&1 | grep -E "amp:|passed|failed|Error" | head
