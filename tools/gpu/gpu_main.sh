#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_learner.py tests/test_arena.py -m gpu -q --maxfail=20 --no-header -rN --tb=short 2>&1 | tail -60 > gpurun_out/pytest_main.log
grep -E "passed|failed" gpurun_out/pytest_main.log | tail -3
grep -E "^(FAILED|ERROR)|^E  |^_{5,}" gpurun_out/pytest_main.log | cut -c1-250 | head -30
timeout 600 python - <<'PY'
import time, torch, sys
sys.path.insert(0, '.')
from boardlaw_b200 import main
t0 = time.time()
def on_step(step, agent, out):
    torch.cuda.synchronize()
    print(f'step {step}: policy_loss {float(out.policy_loss):.4f} value_loss {float(out.value_loss):.4f}  t={time.time()-t0:.1f}s', flush=True)
# the reference's default run shape (boardlaw/main.py:147): 9x9, 32k envs, 64 nodes, buffer of 64 moves; 3 optimiser steps
agent, losses = main.run(boardsize=9, width=256, depth=4, nodes=64, n_envs=32768, buffer_len=64, mix_steps=200, max_steps=3, on_step=on_step)
print('done', time.time() - t0)
PY
