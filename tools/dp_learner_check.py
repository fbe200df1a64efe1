"""Data-parallel learner check, 2+ ranks under torchrun: on ONE gathered chunk of self-play records, the sharded optimiser step (each
rank: forward/backward on its own envs' samples + one all-reduce of the flat gradient) against the replicated one (every rank runs the
whole batch): same losses and weights up to fp32 summation order over three steps, and identical replicas on every rank.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_learner_check.py"""
import os
import sys
import torch
import torch.distributed as dist
sys.path.insert(0, '.')
from boardlaw_b200 import learner, main, selfplay

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=dev)
S, B, T = 5, 512, 8
pool = selfplay.TrajectoryPool()
worlds, network, agent = main.setup(S, 64, 2, nodes=16, n_envs=B, mix_steps=20, device=dev, seed=3, rank=rank)
records = []
for _ in range(T):
    pool.sync_before_overwrite()
    decisions, worlds, transition, rec = agent.play(worlds, record=True)
    pool.gather(rec)
    records.append(pool.wait().clone())
n_all = B * world
g = torch.Generator(device=dev).manual_seed(0)
idxs = (torch.randint(T, (n_all,), device=dev, generator=g), torch.arange(n_all, device=dev))
from boardlaw_b200.networks import FCModel
sd = {k: v.clone() for k, v in network.state_dict().items()}
nets = {}
for m in ('replicated', 'sharded'):
    nets[m] = FCModel(worlds.obs_space, worlds.action_space, width=64, depth=2).to(dev)
    nets[m].load_state_dict(sd)
L = {m: learner.Learner(nets[m], lr=1e-3) for m in nets}
full, _ = learner.chunk_from_records(records, S, n_all)
mine, _ = learner.chunk_from_records(records, S, B, shard=rank)
worst = 0.
for step in range(3):
    a = L['replicated'].optimize(full[idxs])
    b = L['sharded'].optimize(mine[(idxs[0][rank * B:(rank + 1) * B], idxs[1][:B])], group=pool.group, world=world)
    worst = max(worst, abs(float(a.policy_loss) - float(b.policy_loss)), abs(float(a.value_loss) - float(b.value_loss)))
    selfplay.check_replicas(nets['sharded'], pool)
    selfplay.check_replicas(nets['replicated'], pool)
dw = float((L['replicated'].flat - L['sharded'].flat).abs().max())
if rank == 0:
    print(f'world {world}: losses {float(a.policy_loss):.6f}/{float(a.value_loss):.6f} vs {float(b.policy_loss):.6f}/{float(b.value_loss):.6f}; '
          f'max |loss diff| {worst:.2e}, max |weight diff| after 3 steps {dw:.2e}; replicas identical on every rank')
assert worst < 1e-5 and dw < 1e-5, (worst, dw)
dist.barrier()
dist.destroy_process_group()
