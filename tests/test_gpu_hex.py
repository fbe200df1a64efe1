"""GPU parity: Hex step / observe / valid / transition through the C ABI against the oracle, the golden vectors
generated from the reference, and the reference's known-answer tests.  Bit-exact."""
import numpy as np
import pytest
import torch

import golden_checks as gc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('S', [3, 5, 9, 11, 13])
def test_hex_playout_golden(S):
    gc.check_hex_playout(gc.cuda_ops(), S)


def test_hex_kats():
    gc.check_hex_kats(gc.cuda_ops())


@pytest.mark.parametrize('S,B', [(3, 1000), (5, 4096), (9, 4096), (11, 2048), (13, 1024), (17, 300)])
def test_step_observe_vs_oracle(S, B):
    """Random playouts (ragged B, auto-resets, both seats, wins): every board and reward equal to the oracle's."""
    import oracle
    from boardlaw_b200.hex import cuda as hc
    g = torch.Generator().manual_seed(S)
    board = torch.zeros((B, S, S), dtype=torch.uint8)
    seats = torch.zeros((B,), dtype=torch.int32)
    gboard = board.cuda()
    n_wins = 0
    for t in range(2 * S * S):
        obs = oracle.observe(board, seats)
        if t % 5 == 0:
            gobs = hc.observe(gboard, seats.cuda())
            assert torch.equal(gobs.cpu(), obs)
            assert torch.equal(hc.valid(gboard, seats.cuda()).cpu(), (obs == 0).all(-1).reshape(B, -1))
        valid = (obs == 0).all(-1).reshape(B, -1)
        actions = torch.multinomial(valid.float(), 1, generator=g).squeeze(-1).int()
        rew = oracle.step(board, seats, actions)
        grew = hc.step(gboard, seats.cuda(), actions.cuda())
        assert torch.equal(gboard.cpu(), board), f'board differs at step {t}'
        assert torch.equal(grew.cpu(), rew), f'rewards differ at step {t}'
        term = (rew > 0).any(-1)
        n_wins += int(term.sum())
        board[term] = 0
        gboard[term.cuda()] = 0
        seats = 1 - seats
        seats[term] = 0
    assert n_wins > 0


@pytest.mark.parametrize('S', [5, 9])
def test_hex_class_vs_oracle(S):
    """``Hex.step`` (fused transition: clone, move, auto-reset, seat flip) against the restated reference semantics."""
    from boardlaw_b200.hex import Hex
    from oracle import pyref
    B = 777
    g = torch.Generator().manual_seed(1)
    ow = pyref.HexWorld.initial(B, S)
    gw = Hex.initial(B, S, device='cuda')
    for t in range(2 * S * S):
        assert torch.equal(gw.valid.cpu(), ow.valid)
        a = torch.multinomial(ow.valid.float(), 1, generator=g).squeeze(-1)
        ow, ot = ow.step(a)
        gw, gt = gw.step(a.cuda())
        assert torch.equal(gw.board.cpu(), ow.board) and torch.equal(gw.seats.cpu(), ow.seats)
        assert torch.equal(gt.rewards.cpu(), ot.rewards) and torch.equal(gt.terminal.cpu(), ot.terminal)
        assert int(gw.errors) == 0
    # 2-D actions, mask get/set, clone, stack (arena-style use, boardlaw/arena/common.py:90-94)
    a2 = torch.stack([a // S, a % S], -1)
    mask = torch.arange(B) % 3 == 0
    sub = gw[mask.cuda()]
    assert sub.n_envs == int(mask.sum())
    valid = sub.valid
    a_sub = torch.multinomial(valid.float().cpu(), 1, generator=g).squeeze(-1).cuda()
    sub2, _ = sub.step(torch.stack([a_sub // S, a_sub % S], -1))
    gw[mask.cuda()] = sub2
    assert torch.equal(gw.board[mask.cuda()], sub2.board)
    # occupied cell -> device-side error word, no host assert
    w = Hex.initial(4, S, device='cuda')
    w, _ = w.step(torch.zeros(4, dtype=torch.long, device='cuda'))
    w = Hex(board=w.board, seats=torch.zeros_like(w.seats))
    w2, _ = w.step(torch.zeros(4, dtype=torch.long, device='cuda'))
    assert int(w2.errors) == 2


def test_edge_shapes():
    from boardlaw_b200.hex import cuda as hc
    import oracle
    # empty batch
    b = torch.zeros((0, 5, 5), dtype=torch.uint8, device='cuda')
    s = torch.zeros((0,), dtype=torch.int32, device='cuda')
    assert hc.step(b, s, s).shape == (0, 2)
    assert hc.observe(b, s).shape == (0, 5, 5, 2)
    # leading batch dims (B,T,S,S) as MCTS.worlds uses them (boardlaw/hex/cpp/cuda.cu:200-215)
    g = torch.Generator().manual_seed(0)
    board = torch.randint(0, 7, (6, 4, 7, 7), generator=g, dtype=torch.uint8)
    seats = torch.randint(0, 2, (6, 4), generator=g, dtype=torch.int32)
    assert torch.equal(hc.observe(board.cuda(), seats.cuda()).cpu(), oracle.observe(board, seats))
    # dtype / contiguity errors mirror TensorProxy (boardlaw/cpp/common.h:34-38)
    with pytest.raises(TypeError):
        hc.step(torch.zeros((2, 3, 3), dtype=torch.int32, device='cuda'), torch.zeros(2, dtype=torch.int32, device='cuda'),
                torch.zeros(2, dtype=torch.int32, device='cuda'))
    with pytest.raises(RuntimeError):
        hc.step(torch.zeros((2, 3, 6), dtype=torch.uint8, device='cuda')[:, :, ::2], torch.zeros(2, dtype=torch.int32, device='cuda'),
                torch.zeros(2, dtype=torch.int32, device='cuda'))


def test_full_size_properties():
    """BASELINE config 2 size (32768 envs, 9x9): size-independent invariants + oracle agreement on a sample."""
    import oracle
    from boardlaw_b200.hex import Hex
    B, S = 32768, 9
    w = Hex.initial(B, S, device='cuda')
    g = torch.Generator(device='cuda').manual_seed(0)
    sample = torch.arange(0, B, 16)
    for t in range(60):
        a = torch.multinomial(w.valid.float(), 1, generator=g).squeeze(-1)
        before = (w.board != 0).sum((1, 2))
        w2, tr = w.step(a)
        after = (w2.board != 0).sum((1, 2))
        assert int(w2.errors) == 0
        assert torch.equal(after[~tr.terminal], before[~tr.terminal] + 1)        # exactly one new stone
        assert (after[tr.terminal] == 0).all() and (w2.seats[tr.terminal] == 0).all()
        assert torch.equal(w2.seats[~tr.terminal], 1 - w.seats[~tr.terminal])
        assert torch.equal(tr.rewards.sum(-1), torch.zeros(B, device='cuda'))     # zero-sum
        ob = w.board[sample.cuda()].cpu().clone()
        orew = oracle.step(ob, w.seats[sample.cuda()].cpu(), a[sample.cuda()].int().cpu())
        ob[(orew > 0).any(-1)] = 0
        assert torch.equal(w2.board[sample.cuda()].cpu(), ob) and torch.equal(tr.rewards[sample.cuda()].cpu(), orew)
        w = w2
