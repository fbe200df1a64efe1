"""A CPU stand-in for ``boardlaw_b200.hex.Hex`` built on the oracle's C ops — the same record surface (``board``/``seats`` fields,
indexing, ``valid``, ``step``) — so that device-agnostic host logic (arena.neural's Tracker / ChunkEvaluator) can be driven
without a GPU.  TEST INFRASTRUCTURE."""
import torch

import oracle
from boardlaw_b200 import arrdict


class CpuHex(arrdict.namedarrtuple('CpuHex', fields=('board', 'seats'))):

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if not isinstance(self.board, torch.Tensor):
            return
        self.n_seats = 2
        self.n_envs = self.board.shape[0]
        self.boardsize = self.board.shape[-1]
        self.device = self.board.device

    @property
    def valid(self):
        return (oracle.observe(self.board, self.seats) == 0).all(-1).reshape(self.n_envs, -1)

    def step(self, actions):
        assert self.valid.gather(1, actions[:, None].long()).all()
        new_board = self.board.clone()
        rewards = oracle.step(new_board, self.seats.int().contiguous(), actions.int().contiguous())
        terminal = (rewards > 0).any(-1)
        new_board[terminal] = 0
        new_seats = 1 - self.seats
        new_seats[terminal] = 0
        return type(self)(board=new_board, seats=new_seats), arrdict.arrdict(terminal=terminal, rewards=rewards)
