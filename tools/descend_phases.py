"""Phase clock of the descent kernel over one eager move: python tools/descend_phases.py [config]"""
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
d = agent(worlds, use_graph=False)          # warm-up move
worlds, _ = worlds.step(d.actions)
eng = engine_for(worlds, T)
cp = net.packed()
names = ['head', 'sample+advance', 'finish/fetch', 'gap after service', 'child terms', 'pass', 'newton/tail', 'visit: loads issued + parent scan', 'visit: cp.async wait', 'visit: adopt children', 'visit: lambda + scale row', '(lanes outside the visit)', 'visit: alpha seed, child tops']
for label, prof_on in (('plain', False), ('clocked', True)):
    buf = torch.zeros(32, dtype=torch.int64, device='cuda')
    _lib.lib().bl_debug_set_phase_profile(_lib.ptr(buf) if prof_on else None)
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
    print(f'{label}: descend+expand total {sum(ms):.2f} ms/move; sims 5/20/40/60: ' + ' '.join(f'{ms[s - 1] * 1e3:.0f}us' for s in (5, 20, 40, 60) if s < T))
    if prof_on:
        c = buf.cpu().tolist()
        nw = max(c[15], 1)
        tot = sum(c[:11]) + c[12]
        print(f'warps x launches = {nw}; avg cycles per warp per launch = {tot / nw:.0f}')
        for k, n in enumerate(names):
            print(f'  {n:16s} {c[k] / nw:9.0f} cycles  {100 * c[k] / tot:5.1f}%')
        print(f'  trips per warp per launch: {c[13] / nw:.1f} with the service block, {c[14] / nw:.1f} with a pass')
        ncta = max(c[31], 1)
        print(f'network kernel: CTAs x launches = {ncta}')
        for k, n in ((16, 'mma: wait operand'), (17, 'mma: wait weights'), (18, 'mma: issue'), (20, 'epi: board staging'), (21, 'epi: one-hot operand'),
                     (22, 'epi: wait accumulator'), (23, 'epi: layer epilogue'), (25, 'heads: wait accumulator'), (26, 'heads: acc -> registers'),
                     (27, 'heads: legal-move mask'), (28, 'heads: max'), (29, 'heads: sum exp'), (30, 'heads: logits/pi/stores')):
            print(f'  {n:24s} {c[k] / ncta:9.0f} cycles per CTA per launch')
_lib.lib().bl_debug_set_phase_profile(None)
