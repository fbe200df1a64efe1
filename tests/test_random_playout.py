"""Random playouts (learning.mix, boardlaw/learning.py:6-10) — the oracle's restatement of the draw on CPU, and the
fused GPU kernel against it bit for bit on injected uniforms."""
import numpy as np
import pytest
import torch

from oracle import pyref


def test_oracle_draw_is_legal_and_uniform():
    g = torch.Generator().manual_seed(0)
    w = pyref.HexWorld.initial(4000, 5)
    for _ in range(7):
        a = pyref.random_actions(w.valid, torch.rand((w.n_envs,), generator=g))
        assert w.valid.gather(1, a[:, None]).all()
        w, _ = w.step(a)
    # first move on an empty 5x5 board: 25 equiprobable cells
    w0 = pyref.HexWorld.initial(50000, 5)
    a0 = pyref.random_actions(w0.valid, torch.rand((50000,), generator=g))
    counts = np.bincount(a0.numpy(), minlength=25)
    assert counts.min() > 0.85 * 2000 and counts.max() < 1.15 * 2000
    # edge cases of the index arithmetic: u = 0 -> first legal move, u -> 1 -> last legal move
    v = w.valid[:3]
    first = pyref.random_actions(v, torch.zeros(3))
    last = pyref.random_actions(v, torch.full((3,), np.nextafter(np.float32(1), np.float32(0))))
    for b in range(3):
        legal = torch.nonzero(v[b]).squeeze(-1)
        assert first[b] == legal[0] and last[b] == legal[-1]


@pytest.mark.gpu
@pytest.mark.parametrize('S,B', [(3, 257), (5, 1000), (9, 4096), (11, 777), (13, 300)])
def test_random_transition_vs_oracle(S, B):
    """Whole playouts with auto-resets: boards, seats, actions, rewards and terminal flags equal the oracle's at every step."""
    from boardlaw_b200.hex import Hex
    g = torch.Generator().manual_seed(S)
    o = pyref.HexWorld.initial(B, S)
    w = Hex.initial(B, S, device='cuda')
    n_terminal = 0
    for step in range(3 * S * S):
        u = torch.rand((B,), generator=g)
        if step == 1:
            u[:4] = torch.tensor([0., np.nextafter(np.float32(1), np.float32(0)), .5, .25])
        a = pyref.random_actions(o.valid, u)
        o, ot = o.step(a)
        w, t = w.step_random(uniforms=u.cuda())
        assert int(w.errors) == 0
        assert torch.equal(t.actions.cpu(), a), f'actions differ at step {step}'
        assert torch.equal(w.board.cpu(), o.board) and torch.equal(w.seats.cpu(), o.seats.int())
        assert torch.equal(t.rewards.cpu(), ot.rewards) and torch.equal(t.terminal.cpu(), ot.terminal)
        n_terminal += int(ot.terminal.sum())
    assert n_terminal > 0                                          # games did finish and reset along the way


@pytest.mark.gpu
def test_mix_decorrelates():
    """``learning.mix``: after 2A random moves the envs are spread over game phases (stone counts differ) and all boards are legal
    positions (stone counts of the two colours differ by at most one)."""
    from boardlaw_b200.hex import Hex
    from boardlaw_b200.learning import mix
    w = mix(Hex.initial(2048, 7, device='cuda'), T=98, generator=torch.Generator(device='cuda').manual_seed(1))
    b = w.board.cpu().reshape(2048, -1)
    black = ((b == 1) | (b == 3) | (b == 4)).sum(-1)
    white = ((b == 2) | (b == 5) | (b == 6)).sum(-1)
    assert ((black - white) >= 0).all() and ((black - white) <= 1).all()
    assert torch.equal((black - white).int(), w.seats.cpu().int())     # black to move iff equal counts
    assert (black + white).unique().numel() > 10
