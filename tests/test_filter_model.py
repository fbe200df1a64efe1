"""CPU: the certified fast policy evaluation (oracle/filter_model.py, the executable specification of csrc/descend_fx.cu)
against the reference's arithmetic restated in numpy float32 — which is itself checked against the C oracle's descend.
Every decision the fast evaluation certifies must equal the reference's; the rest is flagged for the exact path."""
import numpy as np
import pytest
import torch

import oracle
from oracle import filter_model as fm
from oracle import pyref


@pytest.mark.parametrize('S,B,T,W,D,c_puct', [(5, 96, 24, 32, 2, 1 / 16), (9, 48, 48, 64, 2, 1 / 16), (7, 64, 32, 32, 2, 1.)])
def test_certified_decisions_equal_reference(S, B, T, W, D, c_puct):
    sd = pyref.synth_state_dict(S, W, D, seed=11)
    g = torch.Generator().manual_seed(12)
    w = pyref.HexWorld.initial(B, S)
    for _ in range(S * S // 2):
        w, _ = w.step(torch.multinomial(w.valid.float(), 1, generator=g).squeeze(-1))
    net = pyref.FCNet(sd)
    torch.manual_seed(13)
    stats = fm.Stats()
    pending = {}

    def pre(tree):
        state = torch.get_rng_state()
        rands = torch.rand((B, T), dtype=torch.half)
        torch.set_rng_state(state)
        m = tree._m()
        q = oracle.transition_q(m)
        pending['x'] = fm.replay_descend(m.logits.numpy(), q.numpy(), m.n.numpy(), m.c_puct.numpy(), m.seats.numpy(),
                                         m.terminal.numpy(), m.children.numpy(), rands.numpy(), stats)

    def post(tree, r):
        parents, actions = pending.pop('x')
        # the numpy restatement of the reference's arithmetic is the C oracle's
        assert np.array_equal(parents, r.parents.numpy()) and np.array_equal(actions, r.actions.numpy())

    pyref.agent_call(w, net, n_nodes=T, c_puct=c_puct, hooks={'pre_descend': pre, 'post_descend': post})
    assert stats.evals > 1000
    assert stats.bad_stop == 0 and stats.bad_action == 0, 'a certified decision differs from the reference'
    # the bound holds with room on every evaluation that went through (quantities are <= 1 ulp apart when the ratio nears 1)
    assert stats.max_ne_ratio <= 1. and stats.max_cum_ratio <= 1. and stats.max_alpha_ratio <= 1.
    # the kernel's complete control flow (fast iterations from a safe point, single exact passes on demand) resolves every
    # flagged evaluation to the reference's action and pass count
    assert stats.resolved > 0 and stats.bad_resolved == 0
    flagged = stats.flag_stop + stats.flag_sample + stats.flag_guard + stats.flag_tiny
    assert flagged / stats.evals < .05
