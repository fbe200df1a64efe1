"""Offline study (CPU, numpy) of the CERTIFIED FAST descent: can the regularised-policy evaluation
be decided from closed-form sums over the children only, with a running error bound, and how often does the bound
fail to certify the decision (=> exact fallback)?

Runs the oracle's search at a c2-like shape, captures the inputs of mctscuda.descend before sampled simulations,
and replays every policy evaluation twice: (1) the reference's fp32 sequential arithmetic (bit-exact restatement,
checked against the oracle's own parents/actions), (2) the filtered evaluation.  Prints fallback rates and the
largest observed deviation relative to the bound (must stay < 1; any certified-but-different decision is a bug).

    python tools/filter_study.py [--envs 512] [--S 9] [--T 64] [--every 4]
"""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests'))

import oracle  # noqa: E402
from oracle import pyref  # noqa: E402
from oracle import filter_model as fm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--envs', type=int, default=512)
    ap.add_argument('--S', type=int, default=9)
    ap.add_argument('--T', type=int, default=64)
    ap.add_argument('--W', type=int, default=256)
    ap.add_argument('--D', type=int, default=4)
    ap.add_argument('--every', type=int, default=4)
    ap.add_argument('--moves', type=int, default=1)
    ap.add_argument('--c_puct', type=float, default=1 / 16)
    args = ap.parse_args()
    S, B, T = args.S, args.envs, args.T
    sd = pyref.synth_state_dict(S, args.W, args.D, seed=0)
    g = torch.Generator().manual_seed(0)
    w = pyref.HexWorld.initial(B, S)
    for _ in range(2 * S * S):
        w, _ = w.step(torch.multinomial(w.valid.float(), 1, generator=g).squeeze(-1))
    net = pyref.FCNet(sd)
    torch.manual_seed(0)
    stats = fm.Stats()
    pending = {}

    def pre(tree):
        if tree.sim % args.every and tree.sim != T - 1:
            return
        state = torch.get_rng_state()
        rands = torch.rand((B, T), dtype=torch.half)
        torch.set_rng_state(state)
        m = tree._m()
        q = oracle.transition_q(m)
        pending['x'] = fm.replay_descend(m.logits.numpy(), q.numpy(), m.n.numpy(), m.c_puct.numpy(), m.seats.numpy(),
                                         m.terminal.numpy(), m.children.numpy(), rands.numpy(), stats)

    def post(tree, r):
        if 'x' in pending:
            parents, actions = pending.pop('x')
            assert np.array_equal(parents, r.parents.numpy()) and np.array_equal(actions, r.actions.numpy()), \
                'the numpy restatement of the exact path disagrees with the oracle'

    for mv in range(args.moves):
        d = pyref.agent_call(w, net, n_nodes=T, c_puct=args.c_puct, hooks={'pre_descend': pre, 'post_descend': post})
        w, _ = w.step(d.actions)
    stats.report()


if __name__ == '__main__':
    main()
