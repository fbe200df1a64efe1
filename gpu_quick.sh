#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_random_playout.py tests/test_gpu_hex.py -m gpu -q --no-header -rN --tb=short -x 2>&1 | tail -8
python - <<'PY'
import torch, time
from boardlaw_b200.hex import Hex
from boardlaw_b200.learning import mix
w = Hex.initial(32768, 9, device='cuda')
w = mix(w, 10)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
u = torch.rand((200, 32768), device='cuda')
e0.record()
for i in range(200):
    w, t = w.step_random(uniforms=u[i])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 200
print(f'step_random: {ms*1e3:.1f} us per move of 32768 envs (9x9): {32768/ms/1e3:.1f} M env-steps/s, {32768*(2*81+24)/ms/1e6:.1f} GB/s algorithmic')
PY
