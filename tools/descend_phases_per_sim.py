"""Phase clock of the one-lane descent kernel for single simulations of one eager move: python tools/descend_phases_per_sim.py [config] [sims...]"""
import sys
import torch
sys.path.insert(0, '.')
import bench
from boardlaw_b200 import _lib, heads
from boardlaw_b200.mcts import MCTSAgent, engine_for, dirichlet_mix
from boardlaw_b200.networks import FCModel, synthetic_state_dict

config = sys.argv[1] if len(sys.argv) > 1 else 'c2'
want = [int(x) for x in sys.argv[2:]] or [1, 2, 5, 20, 60]
S, B, T, W, D = bench.CONFIGS[config]
net = FCModel(heads.Tensor((S, S, 2)), heads.Masked(S * S), width=W, depth=D)
net.load_state_dict(synthetic_state_dict(S, W, D, 0))
net = net.cuda()
worlds = bench.make_worlds(S, B, torch.device('cuda', 0), 0)
agent = MCTSAgent(net, n_nodes=T)
torch.manual_seed(0)
d = agent(worlds, use_graph=False)
worlds, _ = worlds.step(d.actions)
eng = engine_for(worlds, T)
cp = net.packed()
names = ['head', 'sample+advance', 'finish/fetch', 'gap after service', 'child terms', 'pass', 'newton/tail', 'visit: loads issued', 'visit: cp.async wait', 'visit: adopt', 'visit: lambda+scale', '(lanes outside visit)', 'visit: alpha seed']
buf = torch.zeros(32, dtype=torch.int64, device='cuda')
eng.reset(worlds.board, worlds.seats, 1 / 16)
eng.eval_root(cp)
eng.set_eval(0, dirichlet_mix(eng.root_logits, worlds.valid, .25, 10), eng.root_v)
for sim in range(1, T):
    prof = sim in want
    if prof:
        buf.zero_()
        _lib.lib().bl_debug_set_phase_profile(_lib.ptr(buf))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.descend_expand(sim); e1.record()
    torch.cuda.synchronize()
    if prof:
        _lib.lib().bl_debug_set_phase_profile(None)
        c = buf.cpu().tolist()
        nw = max(c[15], 1)
        tot = sum(c[:11]) + c[12]
        print(f'sim {sim}: {e0.elapsed_time(e1) * 1e3:.0f} us (clocked build); cycles per warp {tot / nw:.0f}; trips {c[13] / nw:.1f} service / {c[14] / nw:.1f} pass')
        print('   ' + '  '.join(f'{n} {c[k] / nw:.0f}' for k, n in enumerate(names)) + f'  expand/tail {c[6] / nw:.0f}')
    eng.eval_leaves(cp, sim)
    eng.backup(sim)
