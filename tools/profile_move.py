"""One eager (un-graphed) move at a bench config, for ncu: python profile_move.py [config] [n_moves]"""
import sys
import torch
sys.path.insert(0, '.')
import bench
from boardlaw_b200 import heads
from boardlaw_b200.mcts import MCTSAgent, engine_for
from boardlaw_b200.networks import FCModel, synthetic_state_dict

config = sys.argv[1] if len(sys.argv) > 1 else 'c2'
n_moves = int(sys.argv[2]) if len(sys.argv) > 2 else 1
S, B, T, W, D = bench.CONFIGS[config]
net = FCModel(heads.Tensor((S, S, 2)), heads.Masked(S * S), width=W, depth=D)
net.load_state_dict(synthetic_state_dict(S, W, D, 0))
net = net.cuda()
worlds = bench.make_worlds(S, B, torch.device('cuda', 0), 0)
agent = MCTSAgent(net, n_nodes=T)
torch.manual_seed(0)
for _ in range(n_moves):
    d = agent(worlds, use_graph=False)
    worlds, _ = worlds.step(d.actions)
torch.cuda.synchronize()
print('done', float(d.v.float().mean()))
