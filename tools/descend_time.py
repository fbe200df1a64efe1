"""Descent time of one eager move (CUDA events around every descend launch): python tools/descend_time.py [config]
Tuning knobs come from the environment (BL_DESCEND_VARIANT, BL_GATE, BL_MW_GATE, BL_MW_FUSE ...)."""
import sys
import torch
sys.path.insert(0, '.')
import bench
from boardlaw_b200 import heads
from boardlaw_b200.mcts import MCTSAgent, engine_for, dirichlet_mix
from boardlaw_b200.networks import FCModel, synthetic_state_dict

config = sys.argv[1] if len(sys.argv) > 1 else 'c2'
S, B, T, W, D = bench.CONFIGS[config]
net = FCModel(heads.Tensor((S, S, 2)), heads.Masked(S * S), width=W, depth=D)
net.load_state_dict(synthetic_state_dict(S, W, D, 0))
net = net.cuda()
worlds = bench.make_worlds(S, B, torch.device('cuda', 0), 0)
agent = MCTSAgent(net, n_nodes=T)
torch.manual_seed(0)
d = agent(worlds, use_graph=False)          # warm-up move
worlds, _ = worlds.step(d.actions)
eng = engine_for(worlds, T)
cp = net.packed()
for rep in range(2):
    eng.reset(worlds.board, worlds.seats, 1 / 16)
    eng.eval_root(cp)
    eng.set_eval(0, dirichlet_mix(eng.root_logits, worlds.valid, .25, 10), eng.root_v)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(T)]
    for sim in range(1, T):
        ev[sim][0].record(); eng.descend_expand(sim); ev[sim][1].record()
        eng.eval_leaves(cp, sim)
        eng.backup(sim)
    torch.cuda.synchronize()
    ms = [ev[s][0].elapsed_time(ev[s][1]) for s in range(1, T)]
print(f'descend+expand {sum(ms):.2f} ms/move; sims 5/20/40/60: ' + ' '.join(f'{ms[s - 1] * 1e3:.0f}us' for s in (5, 20, 40, 60) if s < T))
