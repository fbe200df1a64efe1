"""Checks shared by the CPU tests (oracle vs golden) and the GPU tests (CUDA path vs golden): each takes an ``ops``
namespace with the reference's op signatures and the device to run on."""
import types
from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / 'golden'


def oracle_ops():
    import oracle
    return types.SimpleNamespace(step=oracle.step, observe=oracle.observe, MCTS=oracle.MCTS, Backup=oracle.Backup,
                                 descend=lambda m, rands: oracle.descend(m, rands), root=oracle.root, backup=oracle.backup,
                                 device='cpu')


def cuda_ops():
    from boardlaw_b200.hex import cuda as hc
    from boardlaw_b200.mcts import cuda as mc
    return types.SimpleNamespace(step=hc.step, observe=hc.observe, MCTS=mc.MCTS, Backup=mc.Backup,
                                 descend=mc.descend, root=mc.root, backup=mc.backup, device='cuda')


def t(a, device):
    if a.dtype == np.float16:
        return torch.from_numpy(a.view(np.int16).copy()).view(torch.float16).to(device)
    return torch.from_numpy(a.copy()).to(device)


def bits(x):
    return x.view(torch.int16) if x.dtype == torch.float16 else x


def check_hex_playout(ops, S):
    """Replays the recorded actions of a reference playout: boards, rewards and observations must match exactly."""
    z = np.load(GOLDEN / f'hex_playout_S{S}.npz')
    dev = ops.device
    steps, B = z['actions'].shape
    board = torch.zeros((B, S, S), dtype=torch.uint8, device=dev)
    for k in range(steps):
        seats = t(z['seats'][k], dev)
        if k % 8 == 0:
            obs = ops.observe(board, seats)
            assert torch.equal(obs.cpu().to(torch.uint8), torch.from_numpy(z['obs'][k // 8])), f'obs differs at step {k}'
        rewards = ops.step(board, seats, t(z['actions'][k], dev))
        assert torch.equal(board.cpu(), torch.from_numpy(z['boards'][k])), f'board differs at step {k}'
        assert torch.equal(rewards.cpu(), torch.from_numpy(z['rewards'][k])), f'rewards differ at step {k}'
        board[(rewards > 0).any(-1)] = 0
    return steps * B


def check_mcts_calls(ops, S, T):
    z = np.load(GOLDEN / f'mcts_calls_S{S}_T{T}.npz')
    dev = ops.device
    names = ['logits', 'w', 'n', 'c_puct', 'seats', 'terminal', 'children']
    n_checked = 0
    for k in range(1, T):
        if f'd{k}_in_logits' not in z:
            continue
        m = ops.MCTS(*[t(z[f'd{k}_in_{x}'], dev) for x in names])
        r = ops.descend(m, t(z[f'd{k}_in_rands'], dev))
        assert torch.equal(r.parents.cpu(), torch.from_numpy(z[f'd{k}_out_parents'])), f'descend {k}: parents differ'
        assert torch.equal(r.actions.cpu(), torch.from_numpy(z[f'd{k}_out_actions'])), f'descend {k}: actions differ'
        n_checked += 1
    assert n_checked >= 3
    for k in range(1, T):
        if f'b{k}_in_v' not in z:
            continue
        g = lambda x: t(z[f'b{k}_in_{x}'], dev)
        w, n = g('w'), g('n')
        bk = ops.Backup(g('v'), w, n, g('rewards'), g('parents'), g('terminal'))
        ops.backup(bk, g('leaves'))
        assert torch.equal(n.cpu(), torch.from_numpy(z[f'b{k}_out_n'])), f'backup {k}: n differs'
        assert torch.equal(bits(w).cpu(), bits(t(z[f'b{k}_out_w'], 'cpu'))), f'backup {k}: w differs'
    m = ops.MCTS(*[t(z[f'r_in_{x}'], dev) for x in names])
    p = ops.root(m)
    assert torch.equal(bits(p).cpu(), bits(t(z['r_out_probs'], 'cpu'))), 'root probs differ'
    return z


# ---- the reference's own known-answer tests, restated on an ops namespace ---------------------------------------

B_, W_ = 0, 1
EMPTY, BLACK, WHITE, TOP, BOT, LEFT, RIGHT = range(7)
TL, TC, TR, CL, CC, CR, BL, BC, BR = range(9)


def tokened(dev, *moves):
    board = torch.zeros((1, 3, 3), dtype=torch.uint8, device=dev)
    for ij, v in moves:
        board[:, ij // 3, ij % 3] = v
    return board


def apply(ops, seat, action, board):
    dev = ops.device
    return ops.step(board, torch.tensor([seat], dtype=torch.int32, device=dev), torch.tensor([action], dtype=torch.int32, device=dev))


def check_hex_kats(ops):
    """boardlaw/hex/tests.py:58-91 (single moves incl. white mirroring, wins, floods) and the two regression
    boards of boardlaw/hex/__init__.py:274-297."""
    dev = ops.device

    def move(seat, action, initial, expected):
        r = apply(ops, seat, action, initial)
        assert torch.equal(initial, expected), (seat, action, initial, expected)
        assert (r == 0).all()

    e = lambda: tokened(dev)
    move(B_, CC, e(), tokened(dev, (CC, BLACK)))
    move(B_, TL, e(), tokened(dev, (TL, TOP)))
    move(B_, BR, e(), tokened(dev, (BR, BOT)))
    move(B_, TR, e(), tokened(dev, (TR, TOP)))
    move(W_, CC, e(), tokened(dev, (CC, WHITE)))
    move(W_, TL, e(), tokened(dev, (TL, LEFT)))
    move(W_, BR, e(), tokened(dev, (BR, RIGHT)))
    move(W_, TR, e(), tokened(dev, (BL, LEFT)))       # mirrored

    r = apply(ops, B_, CC, tokened(dev, (TC, TOP), (BC, BOT)))
    assert r.cpu().tolist() == [[1., -1.]]
    r = apply(ops, W_, CC, tokened(dev, (CL, LEFT), (CR, RIGHT)))
    assert r.cpu().tolist() == [[-1., 1.]]

    move(B_, BC, tokened(dev, (CL, BLACK), (CC, BLACK)), tokened(dev, (CL, BOT), (CC, BOT), (BC, BOT)))
    move(W_, TC, tokened(dev, (TC, WHITE), (CC, WHITE)), tokened(dev, (TC, LEFT), (CC, LEFT), (CL, LEFT)))

    # test_bug: actions 5,5,6,1 from the empty board, seats alternating
    board = torch.zeros((1, 3, 3), dtype=torch.uint8, device=dev)
    for k, a in enumerate([5, 5, 6, 1]):
        apply(ops, k % 2, a, board)
    assert board[0].cpu().tolist() == [[0, 0, 0], [5, 0, 1], [4, 2, 0]]
    # test_bug_2
    board = torch.tensor([[[0, 6, 6], [1, 1, 1], [0, 2, 0]]], dtype=torch.uint8, device=dev)
    apply(ops, 0, 6, board)
    assert board[0].cpu().tolist() == [[0, 6, 6], [4, 4, 4], [4, 2, 0]]


def tree3(dev, c_puct, terminal=(False, False, False), reps=1024):
    """The three-node tree of boardlaw/mcts/tests.py:47-63, replicated over envs."""
    d = dict(
        logits=torch.tensor([[1 / 3, 2 / 3], [1 / 4, 3 / 4], [1 / 5, 4 / 5]]).log().half(),
        w=torch.tensor([[0.], [0.], [1.]]).half(),
        n=torch.tensor([2, 1, 1]).short(),
        c_puct=torch.tensor(c_puct).half(),
        seats=torch.tensor([0, 0, 0]).short(),
        terminal=torch.tensor(list(terminal)),
        children=torch.tensor([[1, 2], [-1, -1], [-1, -1]]).short())
    return {k: v[None].repeat_interleave(reps, 0).contiguous().to(dev) for k, v in d.items()}


def assert_distribution(xs, freqs):
    xs = xs.cpu()
    for i, freq in enumerate(freqs):
        actual = (xs == i).float().mean()
        ci = 3 * (freq * (1 - freq) / len(xs)) ** .5
        assert abs(actual - freq) <= ci + 1e-9, f'Expected {freq:.2f}±{ci:.2f} to be {i}, got {actual:.2f}'


def check_mcts_kats(ops):
    """boardlaw/mcts/tests.py:11-24 (root), :34-150 (descend distributions), :189-234 (backup; cast to the dtypes the
    op checks for)."""
    dev = ops.device
    names = ['logits', 'w', 'n', 'c_puct', 'seats', 'terminal', 'children']
    gen = torch.Generator().manual_seed(0)
    rands = lambda B, T: torch.rand((B, T), generator=gen).half().to(dev)

    one = dict(logits=torch.tensor([[1 / 3, 2 / 3]]).log().half(), w=torch.tensor([[0.]]).half(), n=torch.tensor([0]).short(),
               c_puct=torch.tensor(1.).half(), seats=torch.tensor([0]).short(), terminal=torch.tensor([False]),
               children=torch.tensor([[-1, -1]]).short())
    m = ops.MCTS(*[one[k][None].to(dev) for k in names])
    p = ops.root(m)
    assert torch.allclose(p.float().cpu(), torch.tensor([[1 / 3, 2 / 3]]), rtol=1e-3, atol=1e-3)

    m = ops.MCTS(*[one[k][None].repeat_interleave(1024, 0).contiguous().to(dev) for k in names])
    r = ops.descend(m, rands(1024, 1))
    assert_distribution(r.parents, [1])
    assert_distribution(r.actions, [1 / 3, 2 / 3])

    d = tree3(dev, 1000.)
    r = ops.descend(ops.MCTS(*[d[k] for k in names]), rands(1024, 3))
    assert_distribution(r.parents, [0, 1 / 3, 2 / 3])
    assert_distribution(r.actions, [1 / 3 * 1 / 4 + 2 / 3 * 1 / 5, 1 / 3 * 3 / 4 + 2 / 3 * 4 / 5])

    d = tree3(dev, .001)
    r = ops.descend(ops.MCTS(*[d[k] for k in names]), rands(1024, 3))
    assert_distribution(r.parents, [0, 0, 1])
    assert_distribution(r.actions, [1 / 5, 4 / 5])

    d = tree3(dev, 1000., terminal=(False, True, False))
    r = ops.descend(ops.MCTS(*[d[k] for k in names]), rands(1024, 3))
    assert_distribution(r.parents, [1 / 3, 0, 2 / 3])
    assert_distribution(r.actions, [1 / 3 + 2 / 3 * 1 / 5, 2 / 3 * 4 / 5])

    def bk(v, w, n, rewards, parents, terminal, leaf):
        f = lambda x: torch.tensor(x)[None].half().to(dev)
        w_, n_ = f(w), torch.tensor(n)[None].short().to(dev)
        b = ops.Backup(f(v), w_, n_, f(rewards), torch.tensor(parents)[None].short().to(dev),
                       torch.tensor(terminal)[None].to(dev))
        ops.backup(b, torch.tensor([leaf]).short().to(dev))
        return w_.float().cpu()[0, :, 0].tolist(), n_.cpu()[0].tolist()

    w, n = bk([[1.], [2.]], [[3.], [4.]], [5, 6], [[0.], [0.]], [-1, 0], [False, False], 1)
    assert w == [5., 6.] and n == [6, 7]
    w, n = bk([[0.], [0.]], [[0.], [0.]], [0, 0], [[0.], [1.]], [-1, 0], [False, False], 1)
    assert w == [1., 1.] and n == [1, 1]
    w, n = bk([[0.], [1.], [2.]], [[0.], [0.], [0.]], [0, 0, 0], [[0.], [3.], [0.]], [-1, 0, 1], [False, True, False], 2)
    assert w == [3., 3., 2.]


class KthValid:
    """Deterministic agent for the arena fixtures: plays the k-th legal move with k = (mult * sum(board) + 5 * seat) mod
    #legal — a function of the position only, in integer arithmetic, so the reference's CPU env and the GPU env play the
    same games.  Works on any world with ``board``, ``seats`` and ``valid``."""

    def __init__(self, mult):
        self.mult = mult

    def __call__(self, world, eval=False, value=True):
        valid = world.valid
        n = valid.sum(-1)
        k = (self.mult * world.board.long().flatten(1).sum(-1) + 5 * world.seats.long()) % n.clamp(min=1)
        hit = (valid.long().cumsum(-1) == (k + 1)[:, None]) & valid
        return types.SimpleNamespace(actions=hit.long().argmax(-1))
