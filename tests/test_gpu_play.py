"""GPU: the whole move as one captured graph — in-kernel Dirichlet noise mixed into the root prior, the agent's
action drawn in the root kernel, the env transition and the packed trajectory record — against the separate calls it replaces."""
import numpy as np
import pytest
import torch

import gpu_util as gu

pytestmark = pytest.mark.gpu


def _setup(S, B, T, W=32, D=2, seed=0):
    from boardlaw_b200 import heads
    from boardlaw_b200.engine import SearchEngine
    from boardlaw_b200.hex import Hex
    from boardlaw_b200.networks import FCModel, synthetic_state_dict
    net = FCModel(heads.Tensor((S, S, 2)), heads.Masked(S * S), width=W, depth=D)
    net.load_state_dict(synthetic_state_dict(S, W, D, seed=seed))
    w0 = gu.start_position(S, B, S * S // 3, seed=seed + 1)
    worlds = Hex(board=w0.board.cuda(), seats=w0.seats.cuda())
    return net.cuda(), worlds, SearchEngine(B, S, T, 'cuda', mirror_logits=True)


@pytest.mark.parametrize('S', [5, 9, 13])
def test_root_prior_kernel_vs_torch_mix(S):
    """bl_tree_set_root_prior with an injected Dirichlet sample against the torch expression it replaces (dirichlet_mix + set_eval): the
    stored half logits are identical except where the fp32 value sits on a half rounding boundary (expf/logf differ in the last
    ulp between the two code paths), and then one half step apart."""
    from boardlaw_b200.mcts import dirichlet_mix
    B, T = 3000, 4
    net, worlds, eng = _setup(S, B, T)
    A = S * S
    cp = net.packed()
    torch.manual_seed(3)
    draw = torch.distributions.Dirichlet(torch.full((A,), 10 / A)).sample((B,)).cuda()
    eng.reset(worlds.board, worlds.seats, 1 / 16)
    eng.eval_root(cp)
    eng.set_eval(0, dirichlet_mix(eng.root_logits, worlds.valid, .25, 10, draw), eng.root_v)
    want, want_pi, want_aux = eng.ws.prior.clone(), eng.ws.pi[:, 0].clone(), eng.ws.aux[:, 0].clone()
    eng.ws.prior.zero_(); eng.ws.pi[:, 0].zero_()
    eng.set_root_prior(.25, 10, draw)
    got = eng.ws.prior
    fin = torch.isfinite(want.float())
    assert torch.equal(torch.isfinite(got.float()), fin)
    diff = (got.view(torch.int16) != want.view(torch.int16)) & fin
    frac = float(diff.float().sum() / fin.float().sum())
    assert frac < 2e-3, frac
    assert ((got.view(torch.int16).int() - want.view(torch.int16).int()).abs()[diff] <= 1).all()
    same_rows = ~diff.any(-1)
    assert torch.equal(eng.ws.pi[:, 0][same_rows], want_pi[same_rows]) and torch.equal(eng.ws.aux[:, 0][same_rows], want_aux[same_rows])


def test_in_kernel_dirichlet_noise_distribution():
    """The in-kernel draw (Marsaglia-Tsang gammas on Philox, normalised over the legal moves) has the Dirichlet(alpha) law: recovered
    from the stored prior, each env's noise sums to 1 over the legal moves, vanishes elsewhere, and its components have the Beta
    marginal's mean 1/n and variance (1/n)(1-1/n)/(n alpha + 1)."""
    S, B, T = 9, 16384, 4
    net, _, eng = _setup(S, B, T)
    from boardlaw_b200.hex import Hex
    worlds = Hex.initial(B, S, device='cuda')              # empty boards: n = 81 legal moves everywhere
    A, eps, alpha_scale = S * S, .25, 10.
    cp = net.packed()
    eng.reset(worlds.board, worlds.seats, 1 / 16)
    eng.eval_root(cp)
    eng.set_root_prior(eps, alpha_scale)
    noise = (eng.ws.prior.float().exp() - (1 - eps) * eng.root_logits.exp()) / eps      # (half rounding of the stored prior: ~1e-3 relative)
    assert (noise.sum(-1) - 1).abs().max() < 2e-2
    alpha, n = alpha_scale / A, A
    mean, var = noise.mean().item(), noise.var().item()
    assert abs(mean - 1 / n) < 1e-4
    want_var = (1 / n) * (1 - 1 / n) / (n * alpha + 1)
    assert abs(var - want_var) / want_var < .05, (var, want_var)
    # a different move index draws different noise; the same (seed, move) the same
    first = eng.ws.prior.clone()
    eng.reset(worlds.board, worlds.seats, 1 / 16); eng.eval_root(cp); eng.set_root_prior(eps, alpha_scale)
    assert not torch.equal(first, eng.ws.prior)
    # masked: occupied cells get no noise
    w2 = gu.start_position(S, 256, 30, seed=5)
    net2, worlds2, eng2 = _setup(S, 256, T)
    eng2.reset(worlds2.board, worlds2.seats, 1 / 16); eng2.eval_root(net2.packed()); eng2.set_root_prior(eps, alpha_scale)
    assert torch.equal(torch.isfinite(eng2.ws.prior.float()), worlds2.valid)


def test_root_action_draw_vs_numpy():
    """bl_tree_root_act: greedy = argmax of the root policy; otherwise the inverse-CDF draw over exp(half logits) with injected uniforms,
    reproduced in numpy."""
    S, B, T = 7, 2048, 16
    net, worlds, eng = _setup(S, B, T)
    cp = net.packed()
    eng.reset(worlds.board, worlds.seats, 1 / 16)
    eng.eval_root(cp)
    eng.set_root_prior(.25, 10)
    for sim in range(1, T):
        eng.descend_expand(sim); eng.eval_leaves(cp, sim); eng.backup(sim)
    logits, v, n_leaves = [x.clone() for x in eng.root(T)]
    l2, v2, nl2, greedy = [x.clone() for x in eng.root_act(T, greedy=True)]
    assert torch.equal(l2.view(torch.int16), logits.view(torch.int16)) and torch.equal(nl2, n_leaves)
    assert torch.equal(greedy, logits.float().argmax(-1))
    u = torch.rand(B, device='cuda')
    acts = eng.root_act(T, greedy=False, uniforms=u)[3].cpu().numpy()
    w = np.exp(logits.float().cpu().numpy().astype(np.float32)).astype(np.float32)
    cum = np.cumsum(w, -1, dtype=np.float32)
    target = (u.cpu().numpy() * cum[:, -1]).astype(np.float32)
    want = np.array([int(np.argmax((cum[b] >= target[b]) & (w[b] > 0))) if ((cum[b] >= target[b]) & (w[b] > 0)).any() else int(np.nonzero(w[b] > 0)[0][-1])
                     for b in range(B)])
    # (expf on the device and numpy's exp may differ in the last ulp: a draw that lands on a boundary can move by one legal action)
    assert (acts == want).mean() > .999
    assert worlds.valid.cpu().numpy()[np.arange(B), acts].all()
    # in-kernel uniforms: the empirical action distribution follows the root policy
    counts = torch.zeros((B, S * S), device='cuda')
    for rep in range(64):
        eng.ws.counters[6] += 1                             # a new move index: a new draw
        counts[torch.arange(B, device='cuda'), eng.root_act(T, greedy=False)[3]] += 1
    p = logits.float().exp()
    assert ((counts / 64).mean(0) - p.mean(0)).abs().max() < .01


@pytest.mark.parametrize('S,B,T,W,D', [(5, 512, 16, 32, 2), (9, 1024, 32, 64, 2)])
def test_play_is_the_separate_calls_in_one_graph(S, B, T, W, D):
    """MCTSAgent.play = agent(worlds) + worlds.step(actions) + pack_records: the stepped worlds, the transition and the record bytes
    are those of the separate calls applied to play's own decisions; eager and captured execution agree; every launch of the move is
    one of this library's kernels (engine.launches accounts for all of them)."""
    from boardlaw_b200.hex import Hex
    from boardlaw_b200.mcts import MCTSAgent, engine_for
    from boardlaw_b200.selfplay import pack_records, record_width
    net, worlds, _ = _setup(S, B, T, W, D)
    agent = MCTSAgent(net, n_nodes=T)
    for rep in range(3):                                   # first call captures, later calls replay
        d, nw, tr, rec = agent.play(worlds, record=True)
        d, tr, rec = d.clone(), tr.clone(), rec.clone()
        nb, ns = nw.board.clone(), nw.seats.clone()
        assert worlds.valid.gather(1, d.actions[:, None]).all()
        probs = d.logits.float().exp()
        assert torch.allclose(probs.sum(-1), torch.ones(B, device='cuda'), atol=2e-2)
        want_w, want_t = worlds.step(d.actions)
        assert torch.equal(nb, want_w.board) and torch.equal(ns, want_w.seats)
        assert torch.equal(tr.terminal, want_t.terminal) and torch.equal(tr.rewards, want_t.rewards)
        assert rec.shape == (B, record_width(S * S)) and torch.equal(rec, pack_records(worlds, d, tr))
        assert torch.equal(d.n_sims, torch.full((B,), T + 1, device='cuda'))
        worlds = Hex(board=nb, seats=ns)
    eng = engine_for(worlds, T)
    before = eng.launches
    agent.play(worlds, record=True)
    per_move = eng.launches - before
    assert per_move == 3 * (T - 1) + 7, per_move           # reset, root gather + network, root prior, 3 per simulation, root + action, transition, record
    assert int(eng.errors) == 0
    # greedy play picks the argmax of the root policy
    d, _, _ = agent.play(worlds, eval=True)
    assert torch.equal(d.actions, d.logits.argmax(-1))
