"""Helpers for the GPU parity tests: seeded positions, an oracle-fed network, RNG peeking."""
import types

import torch

from oracle import pyref


def start_position(S, B, n_moves, seed):
    """B boards after n_moves uniformly random valid moves (with auto-resets), via the oracle."""
    g = torch.Generator().manual_seed(seed)
    w = pyref.HexWorld.initial(B, S)
    for _ in range(n_moves):
        a = torch.multinomial(w.valid.float(), 1, generator=g).squeeze(-1)
        w, _ = w.step(a)
    return w


def peek(draw):
    """Runs ``draw()`` against torch's global CPU generator and rewinds it, so the oracle's next internal draw
    returns the same numbers."""
    state = torch.get_rng_state()
    out = draw()
    torch.set_rng_state(state)
    return out


class OracleFedNet:
    """network(worlds) evaluated by the oracle's fp32 CPU forward and shipped to the world's device: makes the
    tree arithmetic of a GPU search comparable bit for bit with the oracle's."""

    def __init__(self, sd):
        self.net = pyref.FCNet(sd)

    def __call__(self, world):
        from boardlaw_b200 import arrdict
        w = pyref.HexWorld(world.board.cpu(), world.seats.cpu().int())
        r = self.net(w)
        return arrdict.arrdict(logits=r.logits.to(world.board.device), v=r.v.to(world.board.device))

    def on_cpu(self, board, seats):
        return self.net(pyref.HexWorld(board, seats))


def dense_children(parents, relation, A):
    """(B,T,A) children tensor implied by (parents, relation) — for comparing against the oracle's."""
    B, T = parents.shape
    ch = torch.full((B, T, A), -1, dtype=torch.int16)
    for k in range(T):
        has = parents[:, k] >= 0
        b = torch.nonzero(has).squeeze(-1)
        ch[b, parents[b, k].long(), relation[b, k].long()] = k
    return ch
