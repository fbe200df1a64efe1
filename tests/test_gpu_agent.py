"""GPU: MCTSAgent end to end (fused engine behind the reference's surface) — outputs, determinism of the injected
mode, agreement with the oracle's search when both run their own fp32 networks, and the counters bench.py reads."""
import pytest
import torch

import gpu_util as gu

pytestmark = pytest.mark.gpu


def make_agent(S, W, D, T, seed=0):
    from boardlaw_b200 import heads
    from boardlaw_b200.mcts import MCTSAgent
    from boardlaw_b200.networks import FCModel
    from oracle import pyref
    sd = pyref.synth_state_dict(S, W, D, seed=seed)
    net = FCModel(heads.Tensor((S, S, 2)), heads.Masked(S * S), width=W, depth=D)
    net.load_state_dict(sd)
    return MCTSAgent(net.cuda(), n_nodes=T), sd


def test_agent_surface():
    """config 1 of BASELINE.json: Hex 5x5, 256 envs, 16 sims/move, 2-layer net."""
    from boardlaw_b200.hex import Hex
    agent, _ = make_agent(5, 32, 2, 16)
    worlds = Hex.initial(256, 5, device='cuda')
    torch.manual_seed(0)
    for _ in range(12):
        d = agent(worlds)
        assert d.logits.shape == (256, 25) and d.logits.dtype == torch.float16
        assert d.prior.shape == (256, 25) and d.v.shape == (256, 2)
        assert d.n_sims.tolist() == [17] * 256                     # reference counts n_nodes+1 (mcts/__init__.py:226)
        assert (d.n_leaves >= 1).all() and (d.n_leaves <= 15).all()
        assert d.actions.dtype == torch.int64
        # sampled actions are valid moves
        assert worlds.valid.gather(1, d.actions[:, None]).all()
        probs = d.logits.float().exp()
        assert torch.allclose(probs.sum(-1), torch.ones(256, device='cuda'), atol=2e-2)
        worlds, tr = worlds.step(d.actions)
        assert int(worlds.errors) == 0
    sd = agent.state_dict()
    assert 'network.body.0.weight' in sd and 'kwargs.n_nodes' in sd and 'network.body.1.α' in sd
    agent.load_state_dict(sd)
    # eval mode is the argmax of the root policy
    d = agent(worlds, eval=True)
    assert torch.equal(d.actions, d.logits.argmax(-1))


@pytest.mark.parametrize('S,W,D,B,T', [(7, 64, 3, 300, 32), (11, 512, 2, 300, 24)])
def test_fused_vs_op_level_same_device(S, W, D, B, T):
    """The fused engine and the op-level ``MCTS`` loop, both on the GPU with the same network kernels and injected
    randomness, build the same search (they share no tree code above mcts_core.cuh).  W = 512 covers the wide network
    kernel's tree mode (net_tc_wide.cu) against its plain mode."""
    from boardlaw_b200.engine import SearchEngine
    from boardlaw_b200.hex import Hex
    from boardlaw_b200.mcts import MCTS
    agent, sd = make_agent(S, W, D, T, seed=3)
    net = agent.network
    w0 = gu.start_position(S, B, 10, seed=1)
    world = Hex(board=w0.board.cuda(), seats=w0.seats.cuda())
    g = torch.Generator().manual_seed(5)
    draw = torch.distributions.Dirichlet(torch.full((S * S,), 10 / (S * S))).sample((B,))
    rands = [torch.rand((B, T), generator=g).half().cuda() for _ in range(T)]

    m = MCTS(world, n_nodes=T)
    m.initialize(net, noise=draw.cuda())
    for sim in range(1, T):
        m.simulate(net, rands=rands[sim])
    r = m.root()

    eng = SearchEngine(B, S, T, 'cuda')
    eng.reset(world.board, world.seats, 1 / 16)
    cp = net.packed()
    logits, v = eng.eval_root(cp)
    from boardlaw_b200.mcts import dirichlet_mix
    eng.set_eval(0, dirichlet_mix(logits, world.valid, .25, 10, draw.cuda()), v)
    for sim in range(1, T):
        eng.descend_expand(sim, rands[sim])
        eng.eval_leaves(cp, sim)
        eng.backup(sim)
    el, ev, en = eng.root(T)
    assert torch.equal(eng.ws.parents, m.tree.parents) and torch.equal(eng.ws.relation, m.tree.relation)
    assert torch.equal(eng.ws.n, m.stats.n) and torch.equal(eng.ws.w.view(torch.int16), m.stats.w.view(torch.int16))
    assert torch.equal(eng.children_dense(), m.tree.children)
    assert torch.equal(el.view(torch.int16), r.logits.view(torch.int16))
    assert torch.equal(en, m.n_leaves())


def test_graph_replay_matches_eager():
    """CUDA-graph replay of a whole move gives the same result as eager launches (same in-kernel random stream)."""
    from boardlaw_b200.engine import SearchEngine
    S, B, T = 5, 512, 16
    agent, _ = make_agent(S, 32, 2, T)
    w0 = gu.start_position(S, B, 6, seed=2)
    board, seats = w0.board.cuda(), w0.seats.cuda()
    draw = torch.distributions.Dirichlet(torch.full((S * S,), 10 / (S * S))).sample((B,)).cuda()
    outs = []
    for use_graph in (False, True):
        eng = SearchEngine(B, S, T, 'cuda', seed=11)
        for move in range(3):                       # move 0 captures, moves 1-2 replay
            l, p, v, n = eng.search(board, seats, agent.network, noise=draw, use_graph=use_graph)
            outs.append((l.clone(), n.clone(), eng.ws.parents.clone(), eng.ws.n.clone()))
    for k in range(3):
        for x, y in zip(outs[k], outs[3 + k]):
            assert torch.equal(x, y), f'graph replay differs from eager at move {k}'
    assert not torch.equal(outs[0][2], outs[1][2])     # a fresh random stream every move


def test_fused_agent_vs_oracle_agreement():
    """Both sides run their OWN fp32 network (GPU kernels vs CPU oracle); logits differ by ~1e-6 before the tree rounds
    them to half, so a small fraction of envs may see one differently-rounded prior.  The searches must agree on
    almost every env: reported, and bounded."""
    from boardlaw_b200.engine import SearchEngine
    from boardlaw_b200.mcts import dirichlet_mix
    from oracle import pyref
    S, B, T, W, D = 9, 256, 64, 64, 4
    agent, sd = make_agent(S, W, D, T, seed=7)
    w0 = gu.start_position(S, B, 20, seed=3)
    torch.manual_seed(9)
    draw = gu.peek(lambda: torch.distributions.Dirichlet(torch.full((S * S,), 10 / (S * S))).sample((B,)))
    o = pyref.Tree(w0, n_nodes=T)
    onet = pyref.FCNet(sd)
    o.initialize(onet)
    eng = SearchEngine(B, S, T, 'cuda')
    eng.reset(w0.board.cuda(), w0.seats.cuda(), 1 / 16)
    cp = agent.network.packed()
    logits, v = eng.eval_root(cp)
    valid = w0.valid.cuda()
    eng.set_eval(0, dirichlet_mix(logits, valid, .25, 10, draw.cuda()), v)
    for sim in range(1, T):
        rands = gu.peek(lambda: torch.rand((B, T), dtype=torch.half))
        eng.descend_expand(sim, rands.cuda())
        eng.eval_leaves(cp, sim)
        eng.backup(sim)
        o.simulate(onet)
    same_tree = ((eng.ws.parents.cpu() == o.parents) & (eng.ws.relation.cpu() == o.relation)).all(-1)
    el, ev, en = eng.root(T)
    same_policy = (el.cpu().view(torch.int16) == o.root().logits.view(torch.int16)).all(-1)
    frac_tree, frac_policy = same_tree.float().mean().item(), same_policy.float().mean().item()
    print(f'identical trees: {frac_tree:.4f}, identical root policies: {frac_policy:.4f}')
    assert frac_tree >= 0.9


def test_rank_seeds_split():
    """main.setup: ranks share the network's initial weights and differ in everything that is drawn during play."""
    from boardlaw_b200 import main
    w0, n0, a0 = main.setup(5, 32, 2, nodes=8, n_envs=64, mix_steps=12, seed=3, rank=0)
    w1, n1, a1 = main.setup(5, 32, 2, nodes=8, n_envs=64, mix_steps=12, seed=3, rank=1)
    for (k, p), (_, q) in zip(n0.state_dict().items(), n1.state_dict().items()):
        assert torch.equal(p, q), f'{k} differs between ranks'
    assert not torch.equal(w0.board, w1.board), 'both ranks decorrelated their worlds with the same playouts'
    # same position, same torch seed: the engines' in-kernel random streams still differ by rank
    torch.manual_seed(0); d0 = a0(w0)
    torch.manual_seed(0); d1 = a1(w0)
    assert not torch.equal(d0.logits, d1.logits)


def test_hex_error_word_persists_and_check_raises():
    from boardlaw_b200.hex import Hex
    w = Hex.initial(8, 5, device='cuda')
    a = torch.zeros(8, dtype=torch.long, device='cuda')
    w1, _ = w.step(a)
    w1.check()
    w2, _ = w1.step((a + 1))
    assert w2.errors.data_ptr() == w1.errors.data_ptr(), 'one error word per chain of worlds'
    w3, _ = w2.step(a)                                    # cell 0 is occupied (black's first move; white's frame is transposed: (0,0) again)
    w4, _ = w3.step(a + 7)                                # a legal move afterwards does not clear the record
    with pytest.raises(AssertionError):
        w4.check()


def test_graph_cache_survives_model_replacement():
    """Graphs are keyed on a per-model token, not id(): a new model of the same shape (possibly at the same address) gets its own."""
    import gc
    from boardlaw_b200.hex import Hex
    from boardlaw_b200.mcts import engine_for
    worlds = Hex.initial(64, 5, device='cuda')
    outs = []
    for seed in (1, 2, 3, 4, 5, 6):
        agent, _ = make_agent(5, 32, 2, 8, seed=seed)
        torch.manual_seed(0)
        d = agent(worlds, eval=True)                      # captures graphs for this model ...
        d = agent(worlds, eval=True)                      # ... and replays them
        # the root value is the model's own value output at the root, whatever the search draws: a replay against another model's
        # (freed) operands would not reproduce it
        ref = agent.network(worlds)
        assert torch.equal(d.v.view(torch.int16), ref.v.half().view(torch.int16))
        outs.append(d.v.clone())
        del agent
        gc.collect()
    eng = engine_for(worlds, 8)
    assert len(eng._graph_lru) <= eng.MAX_GRAPH_KEYS and len(eng._graphs) <= 2 * eng.MAX_GRAPH_KEYS
    assert not all(torch.equal(outs[0], o) for o in outs[1:])


def test_large_board_falls_back_to_op_level():
    """16x16 has 256 actions, one more than the engine's node records hold: mcts() must take the op-level path, not fail."""
    from boardlaw_b200.hex import Hex
    from boardlaw_b200.mcts import MCTS, mcts
    agent, _ = make_agent(16, 32, 1, 4)
    worlds = Hex.initial(6, 16, device='cuda')
    m = mcts(worlds, agent.network, n_nodes=4)
    assert isinstance(m, MCTS)
    d = agent(worlds)
    assert worlds.valid.gather(1, d.actions[:, None]).all()
