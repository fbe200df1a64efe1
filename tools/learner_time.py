"""Time of one optimiser step (Learner.optimize: forward, loss, backward, Adam) on a c2-shaped batch, next to the same step through
torch autograd + cuBLAS (fp32, and TF32 allowed) — python tools/learner_time.py [n_samples]"""
import sys
import torch
sys.path.insert(0, '.')
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


def timed(f, n=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        f()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


L = Learner(net, lr=1e-4)
print(f'Learner.optimize (tcgen05 GEMMs), {N} samples: {timed(lambda: L.optimize(batch)):.3f} ms per step')

# the same step through autograd + cuBLAS
import copy
ref = copy.deepcopy(net)
opt = torch.optim.Adam(ref.parameters(), lr=1e-4)
obs, valid, seats = worlds.obs.reshape(N, -1).float(), worlds.valid, worlds.seats


def autograd_step():
    body = list(ref.body)
    x = torch.addmm(body[0].bias, obs, body[0].weight.t())
    for blk in body[1:]:
        x = x + getattr(blk, 'α') * torch.addmm(blk.bias, torch.relu(x), blk.weight.t())
    logp = torch.log_softmax(torch.addmm(ref.policy.core.bias, x, ref.policy.core.weight.t()).masked_fill(~valid, float('-inf')), -1)
    t = torch.tanh(torch.addmv(ref.value.core.bias, x, ref.value.core.weight[0]))
    v = torch.where(seats[:, None] == 0, torch.stack([t, -t], -1), torch.stack([-t, t], -1))
    tl = logits.float()
    pl = -(tl.exp() * logp).where(valid, torch.zeros_like(logp)).sum(-1).mean()
    vl = (batch.reward_to_go.float() - v).pow(2).mean()
    opt.zero_grad()
    (pl + vl).backward()
    opt.step()


for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    print(f'autograd + cuBLAS ({"TF32" if tf32 else "fp32"}): {timed(autograd_step):.3f} ms per step')
