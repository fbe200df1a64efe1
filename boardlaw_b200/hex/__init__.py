"""The Hex environment record — same API as ``boardlaw.hex.Hex`` (boardlaw/hex/__init__.py:120-222).

``board`` is (B,S,S) uint8 in absolute coordinates with the cell codes ``.bwTBLR`` (0 empty, 1 black,
2 white, 3/4 black joined to top/bottom, 5/6 white joined to left/right); ``seats`` is (B,) int32.
Actions, ``obs`` and ``valid`` are in the mover's frame (white sees the transposed board).
"""
import torch

from .. import arrdict, heads
from . import cuda

CHARS = '.bwTBLR'
ORDS = {c: i for i, c in enumerate(CHARS)}

# Debug switch: re-enables the reference's host-side asserts in ``step`` (two device->host syncs per call,
# boardlaw/hex/__init__.py:174,179).  Off, violations are recorded in ``Hex.errors`` on the device.
STRICT = False


class Hex(arrdict.namedarrtuple('Hex', fields=('board', 'seats'))):

    @classmethod
    def initial(cls, n_envs, boardsize=11, device='cuda'):
        # Black plays first (boardlaw/hex/__init__.py:122-127)
        return cls(
            board=torch.full((n_envs, boardsize, boardsize), 0, device=device, dtype=torch.uint8),
            seats=torch.full((n_envs,), 0, device=device, dtype=torch.int))

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._restore()

    def _restore(self):
        if not isinstance(self.board, torch.Tensor):
            # intermediate records full of bound methods (``worlds.clone`` before the call)
            return
        self.n_seats = 2
        self.n_envs = self.board.shape[0]
        self.boardsize = self.board.shape[-1]
        self.device = self.board.device
        self.obs_space = heads.Tensor((self.boardsize, self.boardsize, 2))
        self.action_space = heads.Masked(self.boardsize * self.boardsize)
        self._obs = None
        self._valid = None

    @property
    def obs(self):
        if self._obs is None:
            self._obs = cuda.observe(self.board, self.seats)
        return self._obs

    @property
    def valid(self):
        if self._valid is None:
            self._valid = cuda.valid(self.board, self.seats)
        return self._valid

    def step(self, actions, reset=True):
        """actions: (n_envs,) flat cell indices or (n_envs, 2) (row, col) pairs, mover's frame, row-major from
        the top-left.  Returns (new_world, arrdict(terminal, rewards)) exactly as boardlaw/hex/__init__.py:161-195:
        finished games are reset to an empty board with black to move."""
        if self.board.ndim != 3:
            raise ValueError('You can only step a board with a single batch dimension')
        if actions.ndim == 2:
            actions = actions[..., 0] * self.boardsize + actions[:, 1]
        assert actions.shape == (self.n_envs,)
        actions = actions.long().contiguous()
        if STRICT:
            assert (0 <= actions).all(), 'You passed a negative action'
            assert self.valid.gather(1, actions[:, None]).squeeze(-1).all()

        errors = self._error_word()
        new_board, new_seats, rewards, terminal = cuda.transition(
            self.board.contiguous(), self.seats.int().contiguous(), actions, reset, errors)
        new_world = type(self)(board=new_board, seats=new_seats)
        new_world.errors = errors
        transition = arrdict.arrdict(terminal=terminal, rewards=rewards)
        return new_world, transition

    def _error_word(self):
        """One int32 on the device per chain of worlds: ``step`` ORs rule violations into it (bit 0: negative / out-of-range
        action, bit 1: occupied cell) and hands the same tensor to the world it returns, so a whole loop of steps costs one
        allocation and can be checked once, at a point where the host synchronises anyway (``check()``)."""
        errors = getattr(self, 'errors', None)
        if errors is None or errors.device != self.board.device:
            errors = self.board.new_zeros((), dtype=torch.int32)
        return errors

    def check(self):
        """Raises if any step of the chain that led to this world broke the rules (the reference asserts on the host inside every
        step, boardlaw/hex/__init__.py:174-179: two device syncs per step; ``STRICT = True`` restores that).  Synchronises."""
        errors = getattr(self, 'errors', None)
        if errors is not None and int(errors) != 0:
            raise AssertionError(f'invalid action(s) were played (error bits {int(errors):#x}: 1 = negative or out of range, 2 = occupied cell)')
        return self


    def step_random(self, uniforms=None, generator=None, reset=True):
        """One step of a uniformly random playout in one kernel: every env plays a uniformly drawn legal move
        (``Categorical(probs=worlds.valid.float()).sample()`` + ``worlds.step``, boardlaw/learning.py:8-9).  ``uniforms``
        (n_envs,) f32 in [0,1) may be injected; otherwise they are drawn on the device.  Returns
        (new_world, arrdict(terminal, rewards, actions))."""
        if uniforms is None:
            uniforms = torch.rand((self.n_envs,), device=self.device, generator=generator)
        errors = self._error_word()
        new_board, new_seats, actions, rewards, terminal = cuda.random_transition(
            self.board.contiguous(), self.seats.int().contiguous(), uniforms.float().contiguous(), reset, errors)
        new_world = type(self)(board=new_board, seats=new_seats)
        new_world.errors = errors
        return new_world, arrdict.arrdict(terminal=terminal, rewards=rewards, actions=actions)


class Solitaire(Hex):
    """One-player Hex: the env plays the other seat itself, so the caller only ever moves for seat 0 and receives a single
    reward column (boardlaw/hex/__init__.py:224-253)."""

    @classmethod
    def initial(cls, *args, seat=0, **kwargs):
        if seat == 1:
            raise ValueError('Can\'t do seat #1 right now')
        return super().initial(*args, **kwargs)

    def _restore(self):
        super()._restore()
        if isinstance(self.board, torch.Tensor):
            self.n_seats = 1

    def step(self, actions):
        worlds, transitions = Hex.step(self, actions)                      # (an instance of type(self), carrying the chain's error word)
        # the move may have ended the game (auto-reset: the player's seat is up again); otherwise the opponent replies until it is
        # the player's turn — in Hex exactly one reply, but the loop keeps the reference's general form
        while True:
            idx = (worlds.seats != self.seats).nonzero().squeeze(-1)
            if idx.numel() == 0:
                break
            sub = Hex(board=worlds.board[idx], seats=worlds.seats[idx])
            sub.errors = worlds.errors
            replied, other = self._play(sub)
            worlds.board[idx], worlds.seats[idx] = replied.board, replied.seats
            transitions.rewards[idx] += other.rewards
            transitions.terminal[idx] |= other.terminal
        envs = torch.arange(self.n_envs, device=self.device)
        transitions['rewards'] = transitions.rewards[envs, self.seats.long()][:, None]
        return worlds, transitions


class Lazy(Solitaire):
    """The opponent plays the first available action (boardlaw/hex/__init__.py:255-263): the 0-th legal move of the fused
    draw-and-step kernel (``bl_hex_random_transition`` with a zero uniform)."""

    @classmethod
    def _play(cls, worlds):
        return worlds.step_random(uniforms=torch.zeros((worlds.n_envs,), device=worlds.device))


class Random(Solitaire):
    """The opponent plays a uniformly random action (boardlaw/hex/__init__.py:265-271), drawn and stepped in one kernel."""

    @classmethod
    def _play(cls, worlds):
        return worlds.step_random()


def from_string(s, **kwargs):
    """Plays out a position drawn with 'b', 'w', '.' (as boardlaw/hex/tests.py:121-134)."""
    import numpy as np
    rows = [l.strip() for l in s.splitlines() if l.strip()]
    size = len(rows)
    grid = np.array([list(r) for r in rows])
    bs = np.stack(np.nonzero(grid == 'b'), -1)
    ws = np.stack(np.nonzero(grid == 'w'), -1)
    assert len(bs) - len(ws) in {0, 1}
    moves = []
    for i in range(len(ws)):
        moves.append([bs[i, 0], bs[i, 1]])
        moves.append([ws[i, 1], ws[i, 0]])
    if len(ws) < len(bs):
        moves.append([bs[-1, 0], bs[-1, 1]])
    worlds = Hex.initial(n_envs=1, boardsize=size, **kwargs)
    for a in torch.as_tensor(np.array(moves).reshape(-1, 2), device=worlds.device):
        worlds, _ = worlds.step(a[None])
    return worlds
