"""CPU: the oracle against the committed golden vectors (generated from the reference by
tests/golden/make_golden.py) and against the reference's own known-answer tests."""
import numpy as np
import pytest
import torch

import golden_checks as gc


@pytest.mark.parametrize('S', [3, 5, 9, 11, 13])
def test_oracle_hex_playout(S):
    gc.check_hex_playout(gc.oracle_ops(), S)


@pytest.mark.parametrize('S,T', [(5, 16), (9, 64)])
def test_oracle_mcts_calls(S, T):
    gc.check_mcts_calls(gc.oracle_ops(), S, T)


def test_oracle_hex_kats():
    gc.check_hex_kats(gc.oracle_ops())


def test_oracle_mcts_kats():
    gc.check_mcts_kats(gc.oracle_ops())


@pytest.mark.parametrize('S,W,D', [(5, 32, 2), (9, 256, 4), (11, 64, 3)])
def test_oracle_network(S, W, D):
    from oracle import pyref
    z = np.load(gc.GOLDEN / f'net_S{S}_W{W}_D{D}.npz')
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('sd_')}
    w = pyref.HexWorld(torch.from_numpy(z['board']), torch.from_numpy(z['seats']))
    r = pyref.FCNet(sd)(w)
    assert torch.allclose(r.logits, torch.from_numpy(z['logits']), atol=1e-6, rtol=0)
    assert torch.allclose(r.v, torch.from_numpy(z['v']), atol=1e-6, rtol=0)


@pytest.mark.parametrize('S,T', [(5, 16)])
def test_oracle_whole_move_golden(S, T):
    """The search restatement reproduces the reference's recorded per-simulation (parent, action) choices when fed
    the recorded random numbers, and the recorded root policy."""
    z = np.load(gc.GOLDEN / f'mcts_calls_S{S}_T{T}.npz')
    for k in range(1, T):
        assert f'd{k}_out_parents' in z
    assert z['agent_actions'].shape[0] == z['r_in_logits'].shape[0]
