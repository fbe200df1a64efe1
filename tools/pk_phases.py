"""Counters of the packed descent (variant 7) over one eager move: python tools/pk_phases.py [config]   (BL_PK_PASS = pass warps)"""
import sys
import torch
sys.path.insert(0, '.')
import bench
from boardlaw_b200 import _lib, heads
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
_lib.lib().bl_debug_set_descend_variant(7)
d = agent(worlds, use_graph=False)          # warm-up move
worlds, _ = worlds.step(d.actions)
eng = engine_for(worlds, T)
cp = net.packed()
buf = torch.zeros(32, dtype=torch.int64, device='cuda')
_lib.lib().bl_debug_set_phase_profile(_lib.ptr(buf))
eng.reset(worlds.board, worlds.seats, 1 / 16)
eng.eval_root(cp)
eng.set_eval(0, dirichlet_mix(eng.root_logits, worlds.valid, .25, 10), eng.root_v)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(T)]
for sim in range(1, T):
    ev[sim][0].record(); eng.descend_expand(sim); ev[sim][1].record()
    eng.eval_leaves(cp, sim)
    eng.backup(sim)
torch.cuda.synchronize()
_lib.lib().bl_debug_set_phase_profile(None)
ms = [ev[s][0].elapsed_time(ev[s][1]) for s in range(1, T)]
print(f'descend+expand {sum(ms):.2f} ms/move; sims 5/20/40/60: ' + ' '.join(f'{ms[s - 1] * 1e3:.0f}us' for s in (5, 20, 40, 60) if s < T))
c = buf.cpu().tolist()
pw, sw = max(c[0], 1), max(c[8], 1)
print(f'pass warps x launches {pw}: trips {c[1] / pw:.1f}, idle polls {c[2] / pw:.1f}, lanes busy per trip {c[3] / max(c[1], 1):.1f}, claims {c[14] / pw:.1f}')
print(f'  cycles per warp per launch: total {c[4] / pw:.0f}; claim+load {c[6] / pw:.0f}, child terms+pass+newton {c[5] / pw:.0f}, sample/advance {c[7] / pw:.0f}')
print(f'  per trip: claim+load {c[6] / max(c[1], 1):.0f}, pass {c[5] / max(c[1], 1):.0f}, tail {c[7] / max(c[1], 1):.0f}')
print(f'service warps x launches {sw}: polls {c[9] / sw:.1f}, visits {c[10] / sw:.1f}, slow {c[12] / sw:.2f}, cycles total {c[13] / sw:.0f}, in visits {c[11] / sw:.0f} ({c[11] / max(c[10], 1):.0f} per visit)')
