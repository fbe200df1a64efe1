"""arena.neural (Tracker, ChunkEvaluator) and the one-player Hex wrappers (Solitaire / Lazy / Random) against fixtures produced by
the REFERENCE's own classes (tests/golden/make_golden_chunk.py): CPU tests drive the device-agnostic host logic on a CPU world
and pin the oracle's restatements; GPU tests run the real env and kernels."""
import numpy as np
import pytest
import torch

import golden_checks as gc

Z = np.load(gc.GOLDEN / 'chunk.npz')
CHUNKS = [(5, 6, (3, 7, 11)), (7, 4, (2, 5, 9, 13))]


def _run_chunk(world_cls, S, n_envs_per, mults, device):
    from boardlaw_b200.arena import neural
    agents = {f'kth{m}': gc.KthValid(m) for m in mults}
    board, seats = torch.from_numpy(Z[f'S{S}_board']).to(device), torch.from_numpy(Z[f'S{S}_seats']).to(device)

    def worldfunc(n):
        assert n == int(Z[f'S{S}_n_envs']) == board.shape[0]
        return world_cls(board=board.clone(), seats=seats.clone())
    ev = neural.ChunkEvaluator(worldfunc, agents, n_envs_per=n_envs_per, device=device)
    results, suggest = [], []
    while not ev.finished():
        name, mask, _ = ev.tracker.suggest(ev.worlds.seats)
        suggest.append((list(agents).index(name), int(mask.sum())))
        results.extend(ev.step())
    results.sort(key=lambda r: r.names)
    assert np.array_equal(np.array(suggest), Z[f'S{S}_suggest']), 'the order of play differs from the reference\'s'
    assert ev.steps == int(Z[f'S{S}_steps'])
    assert [list(r.names) for r in results] == Z[f'S{S}_names'].tolist()
    assert np.array_equal(np.array([r.wins for r in results]), Z[f'S{S}_wins'])
    assert np.array_equal(np.array([r.moves for r in results]), Z[f'S{S}_moves'])
    assert np.array_equal(np.array([r.games for r in results]), Z[f'S{S}_games'])
    assert all(r.boardsize == S for r in results) and len(results) == len(mults) * (len(mults) - 1)


@pytest.mark.parametrize('S,n_envs_per,mults', CHUNKS)
def test_chunk_evaluator_host_logic_vs_reference(S, n_envs_per, mults):
    from cpu_world import CpuHex
    _run_chunk(CpuHex, S, n_envs_per, mults, 'cpu')


def test_tracker_with_games_already_played():
    from boardlaw_b200.arena import neural
    tr = neural.Tracker(3, {'a': {'a': 0, 'b': 2, 'c': 0}, 'b': {'a': 1, 'b': 0, 'c': 3}, 'c': {'a': 0, 'b': 0, 'c': 0}}, device='cpu')
    assert np.array_equal(tr.live.numpy(), Z['tracker_live'])
    assert tr.n_envs == len(Z['tracker_live']) and not tr.finished()


@pytest.mark.parametrize('S', [5, 9])
def test_oracle_lazy_vs_reference(S):
    from oracle import pyref
    w = pyref.LazyWorld(torch.from_numpy(Z[f'lazy_S{S}_board0']).clone(), torch.from_numpy(Z[f'lazy_S{S}_seats0']).clone())
    for t in range(len(Z[f'lazy_S{S}_actions'])):
        a = gc.KthValid(5)(w).actions
        assert np.array_equal(a.numpy(), Z[f'lazy_S{S}_actions'][t])
        w, tr = w.step(a)
        assert np.array_equal(w.board.numpy(), Z[f'lazy_S{S}_boards'][t])
        assert np.array_equal(tr.rewards.numpy(), Z[f'lazy_S{S}_rewards'][t]) and np.array_equal(tr.terminal.numpy(), Z[f'lazy_S{S}_terminal'][t])


@pytest.mark.gpu
@pytest.mark.parametrize('S,n_envs_per,mults', CHUNKS)
def test_gpu_chunk_evaluator_vs_reference(S, n_envs_per, mults):
    from boardlaw_b200.hex import Hex
    _run_chunk(Hex, S, n_envs_per, mults, 'cuda')


@pytest.mark.gpu
@pytest.mark.parametrize('S', [5, 9])
def test_gpu_lazy_vs_reference(S):
    from boardlaw_b200.hex import Lazy
    w = Lazy(board=torch.from_numpy(Z[f'lazy_S{S}_board0']).cuda(), seats=torch.from_numpy(Z[f'lazy_S{S}_seats0']).cuda())
    assert w.n_seats == 1
    for t in range(len(Z[f'lazy_S{S}_actions'])):
        a = gc.KthValid(5)(w).actions
        assert np.array_equal(a.cpu().numpy(), Z[f'lazy_S{S}_actions'][t])
        w, tr = w.step(a)
        assert isinstance(w, Lazy) and tr.rewards.shape == (w.n_envs, 1)
        assert np.array_equal(w.board.cpu().numpy(), Z[f'lazy_S{S}_boards'][t])
        assert np.array_equal(tr.rewards.cpu().numpy(), Z[f'lazy_S{S}_rewards'][t]) and np.array_equal(tr.terminal.cpu().numpy(), Z[f'lazy_S{S}_terminal'][t])
    w.check()


@pytest.mark.gpu
def test_gpu_random_solitaire_plays_legal_games():
    """``Random``: the opponent's reply is a uniformly drawn legal move (no reference pin is possible for the draw itself — it
    would need torch's Categorical stream — so this checks the rules: the player is always to move, stones only ever appear,
    two per step unless the game ended, and wins/losses are +-1 in the single reward column)."""
    from boardlaw_b200.hex import Random
    torch.manual_seed(0)
    w = Random.initial(256, 7, device='cuda')
    for _ in range(40):
        a = torch.multinomial(w.valid.float(), 1).squeeze(-1)
        before = (w.board != 0).flatten(1).sum(-1)
        w2, tr = w.step(a)
        assert (w2.seats == 0).all() and tr.rewards.shape == (256, 1)
        after = (w2.board != 0).flatten(1).sum(-1)
        assert ((after == before + 2) | tr.terminal).all() and (after[tr.terminal] == 0).all()
        assert set(tr.rewards.unique().tolist()) <= {-1., 0., 1.} and (tr.rewards[~tr.terminal] == 0).all() and (tr.rewards[tr.terminal] != 0).all()
        w = w2
    w.check()


@pytest.mark.gpu
def test_gpu_chunk_evaluator_with_search_agents():
    """MCTSAgents on mask-selected sub-batches of one env pool: every pairing completes with n_envs_per games, both colours."""
    from boardlaw_b200 import heads
    from boardlaw_b200.arena import neural
    from boardlaw_b200.hex import Hex
    from boardlaw_b200.mcts import MCTSAgent
    from boardlaw_b200.networks import FCModel, synthetic_state_dict
    S, n_envs_per = 5, 24
    agents = {}
    for k in range(3):
        net = FCModel(heads.Tensor((S, S, 2)), heads.Masked(S * S), width=32, depth=2)
        net.load_state_dict(synthetic_state_dict(S, 32, 2, seed=k))
        agents[f'net{k}'] = MCTSAgent(net.cuda(), n_nodes=8)
    torch.manual_seed(0)
    results = neural.evaluate_chunk(lambda n: Hex.initial(n, S, device='cuda'), lambda name: agents[name], list(agents), n_envs_per=n_envs_per)
    assert len(results) == 6 and {r.names for r in results} == {(a, b) for a in agents for b in agents if a != b}
    assert all(r.games == n_envs_per and sum(r.wins) == n_envs_per and r.moves >= n_envs_per * (2 * S - 1) for r in results)
