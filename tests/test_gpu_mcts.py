"""GPU parity: MCTS descend / root / backup through the C ABI (reference tensor layout), the op-level ``MCTS`` class
and the fused engine, against the oracle and the golden vectors.  Indices and tree statistics are bit-exact."""
import numpy as np
import pytest
import torch

import golden_checks as gc
import gpu_util as gu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('S,T', [(5, 16), (9, 64)])
def test_mcts_calls_golden(S, T):
    gc.check_mcts_calls(gc.cuda_ops(), S, T)


def test_mcts_kats():
    gc.check_mcts_kats(gc.cuda_ops())


def test_exp_and_log_tables():
    """The device tables behind exp(logit) and log(prob) hold the host library's values for every binary16 input."""
    import oracle
    from boardlaw_b200 import _lib
    dev = torch.device('cuda', 0)
    assert np.array_equal(_lib.exp_lut(dev).cpu().numpy().view(np.uint32), oracle.exp_table().view(np.uint32))
    every = torch.arange(65536, dtype=torch.int32).to(torch.int16).view(torch.float16)
    got, want = _lib.log_lut(dev).cpu(), every.float().log().half()
    nan = torch.isnan(want)                                    # log of a negative input: the NaN payload is not pinned by IEEE
    assert torch.equal(torch.isnan(got), nan)
    assert torch.equal(got[~nan].view(torch.int16), want[~nan].view(torch.int16))


def test_transition_q_vs_oracle():
    import oracle
    from boardlaw_b200.mcts import cuda as mc
    g = torch.Generator().manual_seed(0)
    B, T, Sn, A = 300, 17, 2, 4
    w = (torch.randn((B, T, Sn), generator=g) * 3).half()
    n = torch.randint(0, 40, (B, T), generator=g).short()
    dummy = dict(logits=torch.zeros((B, T, A)).half(), c_puct=torch.ones(B).half(), seats=torch.zeros((B, T)).short(),
                 terminal=torch.zeros((B, T), dtype=torch.bool), children=torch.full((B, T, A), -1).short())
    om = oracle.MCTS(dummy['logits'], w, n, dummy['c_puct'], dummy['seats'], dummy['terminal'], dummy['children'])
    oq, omm = oracle.transition_q(om, return_range=True)
    gm = mc.MCTS(dummy['logits'].cuda(), w.cuda(), n.cuda(), dummy['c_puct'].cuda(), dummy['seats'].cuda(),
                 dummy['terminal'].cuda(), dummy['children'].cuda())
    gq, gmm = mc.transition_q(gm)
    assert torch.equal(gmm.cpu(), omm)
    assert torch.equal(gq.cpu().view(torch.int16), oq.view(torch.int16))


def _compare_tree(o, g_children, g, sim, what):
    """o: oracle pyref.Tree; g: dict of CPU tensors in reference layout."""
    for name in ['parents', 'relation', 'n', 'terminal']:
        assert torch.equal(getattr(o, name), g[name]), f'{what}: {name} differs after sim {sim}'
    assert torch.equal(o.children, g_children), f'{what}: children differ after sim {sim}'
    for name in ['w', 'rewards']:
        assert torch.equal(getattr(o, name).view(torch.int16), g[name].view(torch.int16)), f'{what}: {name} differs after sim {sim}'


@pytest.mark.parametrize('S,B,T,W,D', [(5, 256, 16, 32, 2), (9, 192, 64, 64, 4), (11, 64, 96, 32, 2)])
def test_op_level_search_vs_oracle(S, B, T, W, D):
    """The op-level ``MCTS`` class on the GPU (reference tensor layout, reference-shaped ops) against the oracle's
    restated search, fed the same network outputs, Dirichlet draw and random numbers: the whole tree after every
    simulation, and the agent outputs, are identical."""
    from boardlaw_b200.hex import Hex
    from boardlaw_b200.mcts import MCTS
    from oracle import pyref
    sd = pyref.synth_state_dict(S, W, D, seed=1)
    w0 = gu.start_position(S, B, S * S // 3, seed=2)
    net = gu.OracleFedNet(sd)
    torch.manual_seed(3)

    o = pyref.Tree(w0, n_nodes=T)
    gm = MCTS(Hex(board=w0.board.cuda(), seats=w0.seats.cuda()), n_nodes=T)
    A = S * S
    draw = gu.peek(lambda: torch.distributions.Dirichlet(torch.full((A,), 10 / A)).sample((B,)))
    o.initialize(pyref.FCNet(sd))
    gm.initialize(net, noise=draw)
    assert torch.equal(gm.decisions.logits[:, 0].cpu().view(torch.int16), o.logits[:, 0].view(torch.int16))
    for sim in range(1, T):
        rands = gu.peek(lambda: torch.rand((B, T), dtype=torch.half))
        o.simulate(pyref.FCNet(sd))
        gm.simulate(net, rands=rands.cuda())
        g = dict(parents=gm.tree.parents.cpu(), relation=gm.tree.relation.cpu(), n=gm.stats.n.cpu(),
                 terminal=gm.transitions.terminal.cpu(), w=gm.stats.w.cpu(), rewards=gm.transitions.rewards.cpu())
        _compare_tree(o, gm.tree.children.cpu(), g, sim, 'op-level')
        assert torch.equal(gm.worlds.board.cpu(), o.board) and torch.equal(gm.worlds.seats.cpu(), o.seats)
    orr, grr = o.root(), gm.root()
    assert torch.equal(grr.logits.cpu().view(torch.int16), orr.logits.view(torch.int16))
    assert torch.equal(gm.n_leaves().cpu(), o.n_leaves())


class ExtremeNet:
    """Wraps the oracle network and pushes a few valid logits of every row into the range where exp() is denormal or
    zero in fp32 (-70, -95, -110): exercises the engine's exact serial fallback for denormal-range lambda*pi."""

    def __init__(self, net):
        self.net = net

    def __call__(self, world):
        r = self.net(world)
        logits = r.logits.clone()
        for k, val in enumerate((-70., -95., -110.)):
            idx = torch.where(world.valid, torch.arange(world.valid.shape[1])[None], 10 ** 6).kthvalue(k + 2, -1).indices
            rows = torch.arange(logits.shape[0])
            keep = world.valid[rows, idx]
            logits[rows[keep], idx[keep]] = val
        r.logits = logits
        return r


@pytest.mark.parametrize('variant', [7, 6, 5, 4, 3, 2, 1])
@pytest.mark.parametrize('S,B,T,W,D,extreme', [(5, 256, 16, 32, 2, False), (9, 200, 64, 64, 4, False), (11, 70, 128, 32, 2, False),
                                               (13, 33, 32, 16, 1, False), (7, 130, 48, 32, 2, True)])
def test_engine_stepwise_vs_oracle(S, B, T, W, D, extreme, variant):
    from boardlaw_b200 import _lib
    if _lib.lib().bl_debug_set_descend_variant(variant) != 0:
        pytest.skip(f'descent variant {variant} is not compiled in (BL_EXPERIMENTAL=1 at build time)')
    try:
        _engine_stepwise_vs_oracle(S, B, T, W, D, extreme)
    finally:
        _lib.lib().bl_debug_set_descend_variant(0)


def test_shared_reciprocal_division():
    """q0 = n*y, r = fma(-b, q0, n), q = fma(r, y, q0) with y = RN(1/b) against the IEEE division on 2^32 operand pairs
    from the descent's ranges (divisors 2^-27..2^2 incl. all-ones / all-zeros significands, numerators 2^-100..2^-3)."""
    from boardlaw_b200 import _lib
    bad = torch.zeros(1, dtype=torch.int64, device='cuda')
    dev = bad.device
    _lib.check(_lib.lib().bl_selftest_division(1234, 1 << 16, 1 << 16, _lib.ptr(bad), _lib.stream_for(dev)), 'selftest')
    torch.cuda.synchronize()
    assert int(bad) == 0, f'{int(bad)} mismatching quotients'


def _engine_stepwise_vs_oracle(S, B, T, W, D, extreme):
    """The fused engine's kernels (private layout: fp32 pi rows, child lists, per-sim q-range) stepped one simulation
    at a time with the oracle's network outputs injected: leaves, links, boards, statistics and the final root
    policy are identical to the oracle's after every simulation."""
    from boardlaw_b200.engine import SearchEngine
    from boardlaw_b200.mcts import dirichlet_mix
    from oracle import pyref
    sd = pyref.synth_state_dict(S, W, D, seed=4)
    w0 = gu.start_position(S, B, S * S // 3, seed=5)
    onet = pyref.FCNet(sd)
    if extreme:
        onet = ExtremeNet(onet)
    torch.manual_seed(6)
    A = S * S

    o = pyref.Tree(w0, n_nodes=T)
    eng = SearchEngine(B, S, T, 'cuda', mirror_logits=True, fast_descent=True)
    eng.reset(w0.board.cuda(), w0.seats.cuda(), 1 / 16)
    o.initialize(onet)
    # the engine receives the oracle's (noised) root evaluation verbatim
    eng.set_eval(0, o.logits[:, 0].cuda(), o.v[:, 0].cuda())
    envs = torch.arange(B)
    for sim in range(1, T):
        rands = gu.peek(lambda: torch.rand((B, T), dtype=torch.half))
        eng.descend_expand(sim, rands.cuda())
        board, seats = eng.leaf_worlds()
        r = onet(pyref.HexWorld(board.cpu(), seats.cpu()))
        eng.set_eval(-1, r.logits.half().cuda(), r.v.half().cuda())
        eng.backup(sim)
        o.simulate(onet)

        ws = eng.ws
        g = dict(parents=ws.parents.cpu(), relation=ws.relation.cpu(), n=ws.n.cpu(), terminal=ws.terminal.cpu().bool(),
                 w=ws.w.cpu(), rewards=ws.rewards.cpu())
        _compare_tree(o, eng.children_dense().cpu(), g, sim, 'engine')
        populated = (o.parents >= 0)
        populated[:, 0] = True
        gb = ws.board.cpu()[:, :, :A].reshape(B, T, S, S)
        assert torch.equal(gb[populated], o.board[populated]), f'boards differ after sim {sim}'
        assert torch.equal(ws.seats.cpu().int()[populated], o.seats[populated])
        assert torch.equal(ws.logits.cpu()[populated].view(torch.int16), o.logits[populated].view(torch.int16))
        assert torch.equal(ws.v.cpu()[populated].view(torch.int16), o.v[populated].view(torch.int16))
    logits, v, n_leaves = eng.root(T)
    orr = o.root()
    assert torch.equal(logits.cpu().view(torch.int16), orr.logits.view(torch.int16))
    assert torch.equal(v.cpu().view(torch.int16), orr.v.view(torch.int16))
    assert torch.equal(n_leaves.cpu(), o.n_leaves())
    assert torch.equal(eng.ws.prior.cpu().view(torch.int16), orr.prior.view(torch.int16))
    assert int(eng.ws.counters[5]) == 0     # no error events
    # the engine's traffic counters match the tree it built
    assert int(eng.ws.counters[3]) == B * (T - 1)


def test_engine_counters_vs_oracle():
    """Policy-evaluation / children / Newton-iteration counters (the roofline accounting of SURVEY.md 8d) equal the
    oracle's own counts on the same search."""
    import oracle
    from boardlaw_b200.mcts import cuda as mc
    z = np.load(gc.GOLDEN / 'mcts_calls_S9_T64.npz')
    names = ['logits', 'w', 'n', 'c_puct', 'seats', 'terminal', 'children']
    k = 62
    om = oracle.MCTS(*[gc.t(z[f'd{k}_in_{x}'], 'cpu') for x in names])
    cnt = torch.zeros(3, dtype=torch.int64)
    oracle.descend(om, gc.t(z[f'd{k}_in_rands'], 'cpu'), counters=cnt)
    gm = mc.MCTS(*[gc.t(z[f'd{k}_in_{x}'], 'cuda') for x in names])
    c = mc.counters(gm.device)
    c.zero_()
    mc.descend(gm, gc.t(z[f'd{k}_in_rands'], 'cuda'))
    assert c[:3].cpu().tolist() == cnt.tolist()
