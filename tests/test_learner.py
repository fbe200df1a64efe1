"""Learner-side kernels: reward-to-go scan (bit-exact), fused policy/value loss + gradient, Adam, and the
written-out backward of FCModel — against the oracle's CPU restatement (torch autograd / torch.optim.Adam in fp32), which is
itself pinned against the reference's own ``learning.reward_to_go`` and its known answers (boardlaw/learning.py:83-94)."""
import numpy as np
import pytest
import torch

from oracle import pyref, refpy


def _rtg_case(T, B, Sn, seed, p_term=.1):
    g = torch.Generator().manual_seed(seed)
    reward = (torch.rand((T, B, Sn), generator=g) < .1).float() * (torch.randint(0, 2, (T, B, Sn), generator=g) * 2 - 1)
    value = torch.rand((T, B, Sn), generator=g) * 2 - 1
    terminal = torch.rand((T, B), generator=g) < p_term
    return reward, value, terminal


def test_oracle_reward_to_go_known_answers():
    """boardlaw/learning.py:83-94."""
    reward, value = torch.tensor([1., 2., 3.]), torch.tensor([4., 5., 6.])
    out = pyref.reward_to_go(reward, value, torch.tensor([False, False, False]), 1.)
    assert out.tolist() == [9., 8., 6.]
    out = pyref.reward_to_go(reward, value, torch.tensor([False, True, False]), 1.)
    assert out.tolist() == [3., 2., 6.]


@pytest.mark.reference
@pytest.mark.skipif(not refpy.present(), reason='reference sources not present')
def test_oracle_reward_to_go_vs_reference():
    import sys
    refpy.load()
    import boardlaw.learning as rl
    for T, B, Sn, seed, gamma in [(64, 96, 2, 0, 1.), (9, 33, 2, 1, .97), (2, 5, 1, 2, 1.)]:
        reward, value, terminal = _rtg_case(T, B, Sn, seed)
        term = torch.stack([terminal] * Sn, -1)
        want = rl.reward_to_go(reward.clone(), value.clone(), term, gamma)
        got = pyref.reward_to_go(reward, value, term, gamma)
        assert torch.equal(want.view(torch.int32), got.view(torch.int32))


@pytest.mark.gpu
@pytest.mark.parametrize('T,B,Sn,seed,gamma', [(64, 4096, 2, 0, 1.), (9, 33, 2, 1, .97), (2, 5, 1, 2, 1.), (1, 7, 2, 3, 1.), (64, 1, 2, 4, .5)])
def test_reward_to_go_bit_exact(T, B, Sn, seed, gamma):
    from boardlaw_b200 import learner
    reward, value, terminal = _rtg_case(T, B, Sn, seed)
    term = torch.stack([terminal] * Sn, -1)
    want = pyref.reward_to_go(reward, value, term, gamma)
    v_dev = value.cuda()
    got = learner.reward_to_go(reward.cuda(), v_dev, term.cuda(), gamma)
    assert torch.equal(got.cpu().view(torch.int32), want.view(torch.int32))
    assert torch.equal(v_dev.cpu(), value)                                     # inputs untouched
    got_h = learner.reward_to_go(reward.cuda(), value.cuda(), terminal.cuda(), gamma, half=True)      # terminal without the seat axis
    assert torch.equal(got_h.cpu().view(torch.int16), want.half().view(torch.int16))


@pytest.mark.gpu
def test_reward_to_go_known_answers_and_empty():
    from boardlaw_b200 import learner
    reward, value = torch.tensor([1., 2., 3.]).cuda(), torch.tensor([4., 5., 6.]).cuda()
    assert learner.reward_to_go(reward, value, torch.tensor([False, False, False]).cuda()).tolist() == [9., 8., 6.]
    assert learner.reward_to_go(reward, value, torch.tensor([False, True, False]).cuda()).tolist() == [3., 2., 6.]
    e = torch.zeros((4, 0, 2)).cuda()
    assert learner.reward_to_go(e, e, torch.zeros((4, 0), dtype=torch.bool).cuda()).shape == (4, 0, 2)


def _batch(S, N, seed):
    """N positions from random playouts, search-like targets: a noisy softmax over the legal moves, reward-to-go in [-1, 1]."""
    import gpu_util as gu
    w = gu.start_position(S, N, S * S // 3, seed=seed)
    g = torch.Generator().manual_seed(seed + 100)
    raw = torch.randn((N, S * S), generator=g) * 2
    tl = torch.log_softmax(raw.masked_fill(~w.valid, -np.inf), -1).half()
    t = torch.rand((N,), generator=g) * 2 - 1
    tv = torch.stack([t, -t], -1).half()
    return w, tl, tv


@pytest.mark.gpu
@pytest.mark.parametrize('S,W,D,N', [(5, 32, 2, 300), (9, 64, 3, 1000)])
def test_loss_and_gradients_vs_autograd(S, W, D, N):
    """Fused loss kernel + written-out backward against torch autograd on the oracle's fp32 CPU network.  Tolerance: rtol 1e-3 /
    atol 1e-6 on gradients (different summation orders over N samples in fp32), 1e-5 relative on the losses."""
    from boardlaw_b200 import arrdict, heads, learner
    from boardlaw_b200.hex import Hex
    from boardlaw_b200.networks import FCModel
    sd = pyref.synth_state_dict(S, W, D, seed=7)
    w, tl, tv = _batch(S, N, 11)
    _, grads, (pl, vl) = pyref.learner_step(sd, w, tl, tv)
    net = FCModel(heads.Tensor((S, S, 2)), heads.Masked(S * S), width=W, depth=D)
    net.load_state_dict(sd)
    L = learner.Learner(net.cuda())
    batch = arrdict.arrdict(worlds=Hex(board=w.board.cuda(), seats=w.seats.cuda()), decisions=arrdict.arrdict(logits=tl.cuda()),
                            reward_to_go=tv.cuda())
    gpl, gvl = L.forward_backward(batch)
    assert abs(float(gpl) - float(pl)) <= 1e-5 * abs(float(pl)) + 1e-6
    assert abs(float(gvl) - float(vl)) <= 1e-5 * abs(float(vl)) + 1e-6
    for k, want in grads.items():
        got = L.g[k].cpu()
        assert got.shape == want.shape, k
        err = (got - want).abs()
        assert bool((err <= 1e-3 * want.abs() + 1e-6).all()), f'{k}: max err {float(err.max()):.3e} (max |g| {float(want.abs().max()):.3e})'


@pytest.mark.gpu
def test_adam_vs_torch():
    """bl_adam_step against torch.optim.Adam on identical injected gradients, 5 steps."""
    from boardlaw_b200 import heads, learner
    from boardlaw_b200.networks import FCModel
    S, W, D = 5, 32, 2
    sd = pyref.synth_state_dict(S, W, D, seed=3)
    net = FCModel(heads.Tensor((S, S, 2)), heads.Masked(S * S), width=W, depth=D)
    net.load_state_dict(sd)
    L = learner.Learner(net.cuda(), lr=1e-3)
    ref = L.flat.detach().cpu().clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3)
    g = torch.Generator().manual_seed(0)
    for _ in range(5):
        grad = torch.randn(ref.shape, generator=g) * 10 ** torch.randint(-6, 1, ref.shape, generator=g).float()
        ref.grad = grad.clone()
        opt.step()
        L.grad.copy_(grad)
        L.apply()
        assert torch.allclose(L.flat.cpu(), ref.detach(), rtol=1e-6, atol=1e-8)
    assert L.step == 5


@pytest.mark.gpu
def test_optimize_updates_the_network_the_search_uses():
    """Three optimiser steps: parameters track the oracle's (autograd + torch Adam), the loss goes down, and the inference
    kernels (the self-play path) serve the updated weights — the packed operands are restaged."""
    from boardlaw_b200 import arrdict, heads, learner
    from boardlaw_b200.hex import Hex
    from boardlaw_b200.networks import FCModel
    S, W, D, N = 5, 32, 2, 512
    sd = pyref.synth_state_dict(S, W, D, seed=5)
    w, tl, tv = _batch(S, N, 21)
    new_sd, _, (pl0, vl0) = pyref.learner_step(sd, w, tl, tv, lr=1e-3, steps=3)
    net = FCModel(heads.Tensor((S, S, 2)), heads.Masked(S * S), width=W, depth=D)
    net.load_state_dict(sd)
    net = net.cuda()
    worlds = Hex(board=w.board.cuda(), seats=w.seats.cuda())
    before = net(worlds).logits.clone()
    L = learner.Learner(net, lr=1e-3)
    batch = arrdict.arrdict(worlds=worlds, decisions=arrdict.arrdict(logits=tl.cuda()), reward_to_go=tv.cuda())
    losses = [L.optimize(batch) for _ in range(3)]
    assert float(losses[-1].policy_loss + losses[-1].value_loss) < float(losses[0].policy_loss + losses[0].value_loss)
    assert abs(float(losses[0].policy_loss) - float(pl0)) < 1e-4 and abs(float(losses[0].value_loss) - float(vl0)) < 1e-4
    for k, want in new_sd.items():
        got = dict(net.named_parameters())[k].detach().cpu()
        # Adam's first steps move every weight by ~lr whatever the gradient's size: 3 steps of 1e-3, agreement to 2 % of a step
        assert float((got - want).abs().max()) < 6e-5, k
    after = net(worlds)
    want = pyref.FCNet(new_sd)(w)
    assert not torch.equal(after.logits, before)
    fin = torch.isfinite(want.logits)
    assert float((after.logits.cpu()[fin] - want.logits[fin]).abs().max()) < 2e-3      # weights agree to 6e-5, not bit for bit


@pytest.mark.gpu
def test_actor_to_learner_round_trip():
    """A few moves of self-play (fused engine) -> packed trajectory records (what the all-gather carries) -> chunk ->
    one optimiser step on (t, env) samples drawn as boardlaw/main.py:170: the chunk built from the records equals the one
    built from the actor's buffer, reward-to-go equals the oracle's on the same tensors, and the step runs."""
    from boardlaw_b200 import arrdict, heads, learner, selfplay
    from boardlaw_b200.hex import Hex
    from boardlaw_b200.mcts import MCTSAgent
    from boardlaw_b200.networks import FCModel, synthetic_state_dict
    S, W, D, B, T = 5, 32, 2, 128, 12
    net = FCModel(heads.Tensor((S, S, 2)), heads.Masked(S * S), width=W, depth=D)
    net.load_state_dict(synthetic_state_dict(S, W, D, 1))
    agent = MCTSAgent(net.cuda(), n_nodes=16)
    worlds = Hex.initial(B, S, device='cuda')
    torch.manual_seed(0)
    pool = selfplay.TrajectoryPool()
    buffer, records = [], []
    for _ in range(T):
        d = agent(worlds, value=True)
        new_worlds, tr = worlds.step(d.actions)
        buffer.append(arrdict.arrdict(worlds=worlds, decisions=d.half(), transitions=arrdict.arrdict(
            terminal=tr.terminal, rewards=tr.rewards.half())))
        pool.gather(selfplay.pack_records(worlds, d, tr))
        records.append(pool.wait())
        worlds = new_worlds
    chunk, rest = learner.as_chunk(buffer, B)
    chunk2, rest2 = learner.chunk_from_records(records, S, B)
    assert len(rest) == T - 1 and len(rest2) == T - 1
    assert torch.equal(chunk.reward_to_go.view(torch.int16), chunk2.reward_to_go.view(torch.int16))
    assert torch.equal(chunk.worlds.board, chunk2.worlds.board) and torch.equal(chunk.worlds.seats, chunk2.worlds.seats)
    assert torch.equal(chunk.decisions.logits.view(torch.int16), chunk2.decisions.logits.view(torch.int16))
    term = torch.stack([chunk.transitions.terminal.cpu()] * 2, -1)
    want = pyref.reward_to_go(chunk.transitions.rewards.float().cpu(), chunk.decisions.v.float().cpu(), term, 1.).half()
    assert torch.equal(chunk.reward_to_go.cpu().view(torch.int16), want.view(torch.int16))
    assert bool(chunk.transitions.terminal.any())                                      # 12 moves of 5x5: some games ended
    idxs = (torch.randint(T, (B,), device='cuda'), torch.arange(B, device='cuda'))
    L = learner.Learner(agent.network, lr=1e-3)
    out = L.optimize(chunk2[idxs])
    assert torch.isfinite(out.policy_loss) and torch.isfinite(out.value_loss)
    d = agent(worlds, value=True)                                                       # the search runs on the updated weights
    assert bool(torch.isfinite(d.v).all())


@pytest.mark.gpu
def test_main_run_loop():
    """boardlaw_b200.main.run: two rounds of (8 moves of self-play -> chunk -> optimiser step) at a toy size."""
    from boardlaw_b200 import main
    agent, losses = main.run(boardsize=5, width=32, depth=2, nodes=8, n_envs=64, buffer_len=8, mix_steps=10, max_steps=2)
    assert len(losses) == 2
    for l in losses:
        assert torch.isfinite(l.policy_loss) and torch.isfinite(l.value_loss) and float(l.policy_loss) > 0
    # ReZero gates start at 0 (networks.py:15) and must have moved
    assert float(dict(agent.network.named_parameters())['body.1.α'].abs()) > 0


@pytest.mark.gpu
def test_graph_replay_serves_updated_weights():
    """The search's captured CUDA graphs hold the addresses of the staged network operands; an optimiser step writes the
    parameters through raw pointers, so the operands are refreshed IN PLACE (same addresses): the replayed graph must
    evaluate the updated network, not a stale or freed copy."""
    from boardlaw_b200 import arrdict, heads, learner
    from boardlaw_b200.hex import Hex
    from boardlaw_b200.mcts import MCTSAgent
    from boardlaw_b200.networks import FCModel, synthetic_state_dict
    S, W, D, N = 5, 32, 2, 256
    w, tl, tv = _batch(S, N, 31)
    net = FCModel(heads.Tensor((S, S, 2)), heads.Masked(S * S), width=W, depth=D)
    net.load_state_dict(synthetic_state_dict(S, W, D, 8))
    net = net.cuda()
    worlds = Hex(board=w.board.cuda(), seats=w.seats.cuda())
    agent = MCTSAgent(net, n_nodes=8, noise_eps=0.)
    torch.manual_seed(0)
    d0 = agent(worlds)                      # captures the graphs
    agent(worlds)                           # replays them
    gen = net._pack_gen
    L = learner.Learner(net, lr=1e-2)
    batch = arrdict.arrdict(worlds=worlds, decisions=arrdict.arrdict(logits=tl.cuda()), reward_to_go=tv.cuda())
    for _ in range(3):
        L.optimize(batch)
        junk = torch.full((1 << 22,), float('nan'), device='cuda')          # anything freed by the update gets overwritten
        d1 = agent(worlds)
        del junk
        want = net(worlds).logits           # eager forward with the current weights
        assert net._pack_gen == gen         # same shapes: refreshed in place, the staged operands never moved
        assert bool(torch.isfinite(d1.v).all())
        fin = torch.isfinite(want)
        assert torch.equal(torch.isfinite(d1.prior), fin)
        assert float((d1.prior.float()[fin] - want[fin]).abs().max()) < 4e-3      # prior = half(log(exp(logits))) of the root evaluation
    assert float((d1.prior.float()[fin] - d0.prior.float()[fin]).abs().max()) > 1e-2


@pytest.mark.reference
@pytest.mark.skipif(not refpy.present(), reason='reference sources not present')
@pytest.mark.parametrize('S,W,D,B', [(5, 32, 2, 96), (9, 64, 3, 48)])
def test_oracle_optimise_step_vs_reference_main_optimize(S, W, D, B):
    """pyref.learner_step (the oracle the GPU learner is checked against) pinned against the reference's OWN ``main.optimize``
    (boardlaw/main.py:75-101), run on the CPU from where it lies: its function body is compiled out of the reference file and
    executed with the reference's ``learning`` module, its ``FCModel`` and ``Hex``, torch's Adam and GradScaler (disabled on a
    machine without CUDA, as is autocast: fp32 throughout).  Losses and the parameters after one and after three steps agree."""
    import ast
    import contextlib
    import types
    import numpy as np
    r = refpy.load()
    import boardlaw.learning as rlearning
    src = (refpy.REFERENCE / 'boardlaw' / 'main.py').read_text()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == 'optimize')
    recorded = {}

    class Stats:                                     # pavlov.stats: the side channel; records what optimize reports
        defer = staticmethod(contextlib.nullcontext)

        def __getattr__(self, kind):
            return lambda name, *vals: recorded.__setitem__(name, vals)
    ns = dict(torch=torch, np=np, learning=rlearning, stats=Stats())
    exec(compile(ast.Module(body=[fn], type_ignores=[]), str(refpy.REFERENCE / 'boardlaw' / 'main.py'), 'exec'), ns)

    sd = pyref.synth_state_dict(S, W, D, seed=31)
    g = torch.Generator().manual_seed(32)
    w = pyref.HexWorld.initial(B, S)
    for _ in range(S * S // 3):
        w, _ = w.step(torch.multinomial(w.valid.float(), 1, generator=g).squeeze(-1))
    A = S * S
    target_logits = torch.log_softmax(torch.randn((B, A), generator=g).masked_fill(~w.valid, -np.inf), -1).half()
    prior = torch.log_softmax(torch.randn((B, A), generator=g).masked_fill(~w.valid, -np.inf), -1).half()
    target_v = (torch.rand((B, 2), generator=g) * 2 - 1).half()

    rworlds = r.Hex(board=w.board.clone(), seats=w.seats.clone())
    rnet = r.FCModel(rworlds.obs_space, rworlds.action_space, width=W, depth=D)
    rnet.load_state_dict(sd)
    opt = torch.optim.Adam(rnet.parameters(), lr=1e-3)
    scaler = torch.cuda.amp.GradScaler()
    batch = r.arrdict.arrdict(
        worlds=rworlds,
        decisions=r.arrdict.arrdict(logits=target_logits, prior=prior, v=target_v),
        transitions=r.arrdict.arrdict(terminal=torch.zeros(B, dtype=torch.bool)),
        reward_to_go=target_v)
    for steps in (1, 3):
        if steps == 3:
            ns['optimize'](rnet, scaler, opt, batch); ns['optimize'](rnet, scaler, opt, batch)
        else:
            ns['optimize'](rnet, scaler, opt, batch)
            pl_ref, vl_ref = float(recorded['loss.policy'][0]), float(recorded['loss.value'][0])
        new, _, (pl, vl) = pyref.learner_step(sd, w, target_logits, target_v, lr=1e-3, steps=steps)
        if steps == 1:
            assert abs(float(pl) - pl_ref) <= 1e-6 * abs(pl_ref) and abs(float(vl) - vl_ref) <= 1e-6 * abs(vl_ref)
        for k, p in rnet.state_dict().items():
            assert torch.allclose(new[k], p, rtol=0, atol=2e-7), f'{k} after {steps} step(s): {float((new[k] - p).abs().max())}'
