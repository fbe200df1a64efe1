"""GPU parity at the BENCHMARKED shapes: the descent variants against each other on full-size batches,
the engine's own tree-mode network against the oracle's fp32 network at W256/D4, the queue path of the one-lane descent,
and T = 256.  Indices and tree statistics are bit-exact; network rows are within 1e-5 before their half rounding."""
import numpy as np
import pytest
import torch

import gpu_util as gu

pytestmark = pytest.mark.gpu


def _model(S, W, D, seed=0, precision='fp32'):
    from boardlaw_b200 import heads
    from boardlaw_b200.networks import FCModel
    from oracle import pyref
    sd = pyref.synth_state_dict(S, W, D, seed=seed)
    net = FCModel(heads.Tensor((S, S, 2)), heads.Masked(S * S), width=W, depth=D, precision=precision)
    net.load_state_dict(sd)
    return net.cuda(), sd


def _positions(S, B, seed):
    """B positions spread over all game phases (2*A uniformly random moves with auto-resets) on the GPU env."""
    from boardlaw_b200.hex import Hex
    g = torch.Generator(device='cuda').manual_seed(seed)
    worlds = Hex.initial(B, S, device='cuda')
    for _ in range(2 * S * S):
        worlds, _ = worlds.step(torch.multinomial(worlds.valid.float(), 1, generator=g).squeeze(-1))
    return worlds


TREE_FIELDS = ('node', 'aux', 'kids', 'board', 'leaf', 'leaf_parent', 'leaf_action', 'prior', 'qrange', 'parent_of')


def _search(eng, worlds, net, variant, draw, grid=0):
    """One whole eager search with the in-kernel random stream of move 0; returns clones of the tree and the outputs."""
    from boardlaw_b200 import _lib
    lib = _lib.lib()
    if lib.bl_debug_set_descend_variant(variant) != 0:
        return None                                                # not compiled in (BL_EXPERIMENTAL=1 at build time)
    lib.bl_debug_set_descend_grid(grid)
    try:
        eng.move = 0
        eng.ws.counters.zero_()
        out = eng.search(worlds.board, worlds.seats, net, noise=draw, use_graph=False)
        torch.cuda.synchronize()
        state = {k: eng.ws[k].clone() for k in TREE_FIELDS}
        state['pi'] = eng.ws.pi.view(torch.int32).sum(dtype=torch.int64).item()     # checksum: 0.7 GB at c2
        state['out'] = [x.clone() for x in out]
        state['counters'] = eng.ws.counters.cpu().tolist()
        return state
    finally:
        lib.bl_debug_set_descend_variant(0)
        lib.bl_debug_set_descend_grid(0)


def _assert_same(a, b, what):
    if b is None:
        return
    for k in TREE_FIELDS:
        assert torch.equal(a[k], b[k]), f'{what}: {k} differs'
    assert a['pi'] == b['pi'], f'{what}: pi rows differ'
    for x, y in zip(a['out'], b['out']):
        assert torch.equal(x.view(torch.int16) if x.dtype == torch.float16 else x, y.view(torch.int16) if y.dtype == torch.float16 else y), f'{what}: outputs differ'
    assert a['counters'][:5] == b['counters'][:5], f'{what}: traffic counters differ'


@pytest.mark.parametrize('S,B,T,W,D', [(9, 32768, 64, 256, 4), (5, 32768, 64, 256, 4), (13, 8192, 64, 256, 4)])
def test_descent_variants_full_batch(S, B, T, W, D):
    """c2 (and the ends of the c5 sweep) at full size: the certified fast descent (5), the one-lane exact descent (2) — resident
    and through its env queue — the multi-lane exact descent (3) and the lock-step reference loops (1) build bit-identical
    trees over a whole move with the engine's own tree-mode network and in-kernel random numbers."""
    from boardlaw_b200.engine import SearchEngine
    net, _ = _model(S, W, D)
    worlds = _positions(S, B, seed=1)
    A = S * S
    torch.manual_seed(2)
    draw = torch.distributions.Dirichlet(torch.full((A,), 10 / A)).sample((B,))
    eng = SearchEngine(B, S, T, 'cuda', fast_descent=True)
    ref = _search(eng, worlds, net, 2, draw)
    fx = _search(eng, worlds, net, 5, draw)
    _assert_same(ref, fx, 'variant 5 vs 2')
    _assert_same(ref, _search(eng, worlds, net, 6, draw), 'variant 6 vs 2')
    _assert_same(ref, _search(eng, worlds, net, 7, draw), 'variant 7 (packed) vs 2')
    evals = fx['counters'][0]
    flagged = sum(fx['counters'][8:11])
    print(f'\nS{S} B{B}: {evals} evaluations, {flagged} sent to the exact path ({100 * flagged / max(evals, 1):.3f} %: stop '
          f'{fx["counters"][8]}, sample {fx["counters"][9]}, guard/tiny {fx["counters"][10]}), {fx["counters"][11]} exact passes')
    assert flagged / evals < .2
    _assert_same(ref, _search(eng, worlds, net, 2, draw, grid=(B // 32) * 5 // 8), 'variant 2 through the env queue')
    _assert_same(ref, _search(eng, worlds, net, 3, draw), 'variant 3 vs 2')
    _assert_same(ref, _search(eng, worlds, net, 1, draw), 'variant 1 vs 2')
    assert ref['counters'][5] == 0


def _half_neighbours(h):
    """The binary16 values just below and above each entry of h (float16 tensor, finite entries)."""
    bits = h.view(torch.int16).to(torch.int32) & 0xFFFF
    neg = bits >= 0x8000
    mag = bits & 0x7FFF
    up_mag = torch.where(neg, mag - 1, mag + 1)
    dn_mag = torch.where(neg, mag + 1, mag - 1)
    sign = torch.where(neg, 0x8000, 0)
    mk = lambda m: ((m.clamp(min=0) | sign) & 0xFFFF).to(torch.int16).view(torch.float16)
    return mk(dn_mag), mk(up_mag)


@pytest.mark.parametrize('S,B,T,W,D', [(9, 4096, 64, 256, 4)])
def test_tree_mode_network_vs_oracle_c2_shape(S, B, T, W, D):
    """The exact kernel instantiation bench.py times — fc_tc in TREE mode at W256 D4 S9 — on every leaf of a real search,
    against pyref.Tree with the oracle's fp32 network, injected Dirichlet draw and random numbers.  Each half logit / value the
    kernel stores in the tree equals the oracle's, or lies within 1e-5 + half a half-step of the oracle's fp32 value (i.e. the
    values agree to 1e-5 BEFORE rounding: the tree only ever sees halves).  After the check
    the oracle's rows replace ours (transition_q's min/max couples all envs of a batch, so one flipped half would make every
    later comparison meaningless), and the tree — links, counts, values, boards — must then be bit-identical for all envs
    after every simulation and at the root."""
    from boardlaw_b200.engine import SearchEngine
    from boardlaw_b200.mcts import dirichlet_mix
    from oracle import pyref
    net, sd = _model(S, W, D, seed=4)
    cp = net.packed()
    w0 = gu.start_position(S, B, S * S // 3, seed=5)
    A = S * S
    torch.manual_seed(6)
    fp32_rows = {}

    class Rec:
        """The oracle's network, recording its fp32 outputs."""
        def __init__(self):
            self.net = pyref.FCNet(sd)

        def __call__(self, world):
            r = self.net(world)
            fp32_rows['last'] = (r.logits.clone(), r.v.clone())
            return r

    onet = Rec()
    o = pyref.Tree(w0, n_nodes=T)
    eng = SearchEngine(B, S, T, 'cuda', mirror_logits=True)
    o.initialize(onet)
    eng.reset(w0.board.cuda(), w0.seats.cuda(), 1 / 16)
    rl, rv = eng.eval_root(cp)
    ol, ov = fp32_rows['last']
    fin = torch.isfinite(ol)
    assert torch.equal(torch.isfinite(rl.cpu()), fin)
    assert (rl.cpu()[fin] - ol[fin]).abs().max() <= 1e-5 and (rv.cpu() - ov).abs().max() <= 1e-5
    eng.set_eval(0, o.logits[:, 0].cuda(), o.v[:, 0].cuda())
    envs = torch.arange(B)
    worst, flips, total = 0., 0, 0
    for sim in range(1, T):
        rands = gu.peek(lambda: torch.rand((B, T), dtype=torch.half))
        eng.descend_expand(sim, rands.cuda())
        eng.eval_leaves(cp, sim)                     # OUR network, tree mode
        o.simulate(onet)
        ol, ov = fp32_rows['last']                   # the oracle's fp32 evaluation of the same leaves
        ws = eng.ws
        leaf = ws.leaf.cpu().long()
        assert (leaf >= 0).all()
        assert torch.equal(ws.parents.cpu(), o.parents) and torch.equal(ws.relation.cpu(), o.relation), f'links differ at sim {sim}'
        gl, gv = ws.logits.cpu()[envs, leaf], ws.v.cpu()[envs, leaf]
        for got, want32 in ((gl, ol), (gv, ov)):
            want16 = want32.half()
            fin = torch.isfinite(want32)
            assert torch.equal(torch.isfinite(got.float()), fin), f'masked entries differ at sim {sim}'
            diff = fin & (got.view(torch.int16) != want16.view(torch.int16))
            total += int(fin.sum())
            if diff.any():
                flips += int(diff.sum())
                # |ours32 - oracle32| <= 1e-5 and both rounded to half  =>  |half(ours) - oracle32| <= 1e-5 + half a half-step of ours
                g32 = got.float()
                dn, up = _half_neighbours(got)
                half_step = torch.maximum((g32 - dn.float()).abs(), (up.float() - g32).abs()) / 2
                excess = ((g32 - want32).abs() - half_step)[diff]
                worst = max(worst, float(excess.max()))
                assert float(excess.max()) <= 1e-5, f'pre-rounding values differ by more than 1e-5 at sim {sim}: {float(excess.max())}'
        eng.set_eval(-1, ol.half().cuda(), ov.half().cuda())
        eng.backup(sim)
        assert torch.equal(ws.n.cpu(), o.n) and torch.equal(ws.w.cpu().view(torch.int16), o.w.view(torch.int16)), f'statistics differ at sim {sim}'
    populated = (o.parents >= 0)
    populated[:, 0] = True
    gb = eng.ws.board.cpu()[:, :, :A].reshape(B, T, S, S)
    assert torch.equal(gb[populated], o.board[populated])
    logits, v, n_leaves = eng.root(T)
    orr = o.root()
    assert torch.equal(logits.cpu().view(torch.int16), orr.logits.view(torch.int16))
    assert torch.equal(n_leaves.cpu(), o.n_leaves())
    print(f'\nS{S} W{W} D{D} B{B}: {total} stored values checked, {flips} one-half-step flips ({100 * flips / total:.4f} %), '
          f'largest |ours - oracle| beyond the rounding step among them {worst:.2e} (tolerance 1e-5)')


def test_tree_mode_equals_plain_mode_w256():
    """W256 D4 S9: the tree-mode kernel (leaf boards from the tree, pi rows / summaries / values into it) stores exactly
    half(plain-mode fp32 output) -> exp table, for a batch that spans several tiles per CTA."""
    from boardlaw_b200 import _lib
    from boardlaw_b200.engine import SearchEngine
    S, B, T, W, D = 9, 20000, 4, 256, 4
    net, _ = _model(S, W, D, seed=7)
    cp = net.packed()
    worlds = _positions(S, B, seed=8)
    eng = SearchEngine(B, S, T, 'cuda', mirror_logits=True)
    eng.reset(worlds.board, worlds.seats, 1 / 16)
    rl, rv = eng.eval_root(cp)
    eng.set_eval(0, rl, rv)
    eng.descend_expand(1)
    eng.eval_leaves(cp, 1)
    board, seats = eng.leaf_worlds()
    from boardlaw_b200.hex import Hex
    plain = net(Hex(board=board, seats=seats))
    leaf = eng.ws.leaf.long()
    envs = torch.arange(B, device='cuda')
    got_l = eng.ws.logits[envs, leaf]
    got_v = eng.ws.v[envs, leaf]
    assert torch.equal(got_l.view(torch.int16), plain.logits.half().view(torch.int16))
    assert torch.equal(got_v.view(torch.int16), plain.v.half().view(torch.int16))
    lut = _lib.exp_lut(torch.device('cuda', 0))
    want_pi = lut[got_l.view(torch.int16).long() & 0xFFFF]
    assert torch.equal(eng.ws.pi[envs, leaf][:, :S * S], want_pi)


@pytest.mark.parametrize('variant', [6, 5, 3, 2])
def test_engine_stepwise_vs_oracle_t256(variant):
    """c3's tree depth: S11, T = 256 (four 64-bit words of children mask per node) stepwise against the oracle."""
    import test_gpu_mcts as tm
    from boardlaw_b200 import _lib
    if _lib.lib().bl_debug_set_descend_variant(variant) != 0:
        pytest.skip(f'descent variant {variant} is not compiled in')
    try:
        tm._engine_stepwise_vs_oracle(11, 40, 256, 32, 2, False)
    finally:
        _lib.lib().bl_debug_set_descend_variant(0)
