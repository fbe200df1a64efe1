"""Kernel-time breakdown of one Learner.optimize step (torch.profiler): python tools/learner_prof.py [n_samples]"""
import sys
import torch
sys.path.insert(0, '.')
sys.argv = sys.argv[:1] + sys.argv[1:]
import runpy
from torch.profiler import profile, ProfilerActivity
ns = runpy.run_path('tools/learner_time.py', run_name='not_main') if False else None
from boardlaw_b200 import heads, arrdict
from boardlaw_b200.hex import Hex
from boardlaw_b200.learner import Learner
from boardlaw_b200.networks import FCModel, synthetic_state_dict
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
S, W, D = 9, 256, 4
A = S * S
dev = torch.device('cuda')
torch.manual_seed(0)
net = FCModel(heads.Tensor((S, S, 2)), heads.Masked(A), width=W, depth=D)
net.load_state_dict(synthetic_state_dict(S, W, D, 0))
net = net.to(dev)
worlds = Hex.initial(N, S, device=dev)
for _ in range(20):
    worlds, _ = worlds.step(torch.multinomial(worlds.valid.float(), 1).squeeze(-1))
logits = torch.log_softmax(torch.randn((N, A), device=dev).masked_fill(~worlds.valid, float('-inf')), -1).half()
batch = arrdict.arrdict(worlds=worlds, decisions=arrdict.arrdict(logits=logits), reward_to_go=(torch.rand((N, 2), device=dev) * 2 - 1).half())
L = Learner(net, lr=1e-4)
for _ in range(3):
    L.optimize(batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5):
        L.optimize(batch)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=25, max_name_column_width=70))
