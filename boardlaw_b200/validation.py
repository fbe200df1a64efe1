"""Baseline agents for evaluation runs — the call surface of ``boardlaw/validation.py:23-76`` on the B200 env.

``RandomAgent`` plays a uniformly random legal move; ``MonteCarloAgent`` scores every first move by flat random
playouts.  The playouts run on ``Hex.step_random`` (one kernel per ply: draw the k-th legal move and step,
``bl_hex_random_transition``) instead of a ``Categorical`` sample + ``step`` pair per ply.
"""
import numpy as np
import torch

from . import arrdict


def uniform_logits(valid):
    """log(1/#legal) on legal moves, -inf elsewhere (boardlaw/validation.py:79-80)."""
    return torch.log(valid.float() / valid.sum(-1, keepdims=True))


class RandomAgent:
    """boardlaw/validation.py:23-30.  ``eval`` is accepted (arena callers pass it) and ignored, as a random agent has no mode."""

    def __call__(self, world, value=True, eval=False):
        valid = world.valid
        return arrdict.arrdict(
            logits=uniform_logits(valid),
            actions=torch.distributions.Categorical(probs=valid.float()).sample(),
            v=torch.zeros((valid.shape[0], world.n_seats), device=world.device))


class MonteCarloAgent:
    """Flat Monte-Carlo: ``n_rollouts`` uniformly random playouts per env, the first move of each credited with the
    playout's return (boardlaw/validation.py:32-76).  Policy = softmax(temperature * mean return of the mover)."""

    def __init__(self, n_rollouts, temperature=1.):
        self.n_rollouts = n_rollouts
        self.temperature = temperature

    def rollout(self, world):
        """One playout of every env to the end of its game: (return per seat (B, n_seats), first action (B,))."""
        B = world.n_envs
        live = torch.ones((B,), dtype=torch.bool, device=world.device)
        total = torch.zeros((B, world.n_seats), dtype=torch.float, device=world.device)
        first = None
        # a Hex game on an SxS board ends within S*S plies; the liveness test (a device->host sync) is only taken every S plies
        ply, every = 0, max(world.boardsize, 1)
        while True:
            if ply % every == 0 and not bool(live.any()):
                break
            world, t = world.step_random()
            if first is None:
                first = t.actions.long()
            total += t.rewards.float() * live[:, None].float()
            live = live & ~t.terminal
            ply += 1
        return total, first

    def __call__(self, world, value=True, eval=False):
        envs = torch.arange(world.n_envs, device=world.device)
        valid = world.valid
        totals = torch.zeros(valid.shape + (world.n_seats,), dtype=torch.float, device=world.device)
        counts = torch.zeros_like(totals)
        for _ in range(self.n_rollouts):
            r, a = self.rollout(world)
            totals[envs, a] += r
            counts[envs, a] += 1
        means = torch.where(counts > 0, totals / counts.clamp(min=1), torch.zeros_like(totals))
        mover = means[envs, :, world.seats.long()]
        logits = torch.log_softmax(self.temperature * mover, -1)
        logits[~valid] = -np.inf
        return arrdict.arrdict(
            logits=logits,
            actions=logits.argmax(-1) if eval else torch.distributions.Categorical(logits=logits).sample(),
            v=totals.sum(-2) / counts.sum(-2))


class WinnerLoser(arrdict.namedarrtuple('WinnerLoser', fields=('seats',))):
    """Two-move toy game: the first seat wins (+1) when the second seat has moved (boardlaw/validation.py:113-143).  Pure
    tensor arithmetic on whatever device ``seats`` lives on; used by the arena's known-answer test."""

    @classmethod
    def initial(cls, n_envs=1, device='cuda'):
        return cls(seats=torch.zeros(n_envs, device=device, dtype=torch.int))

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if not isinstance(self.seats, torch.Tensor):
            return
        self.device = self.seats.device
        self.n_envs = len(self.seats)
        self.n_seats = 2
        self.boardsize = 0
        self.valid = torch.ones((self.n_envs, 1), dtype=torch.bool, device=self.device)
        self.logits = uniform_logits(self.valid)
        self.v = torch.stack([torch.ones_like(self.seats), -torch.ones_like(self.seats)], -1).float()

    def step(self, actions):
        over = self.seats == 1
        trans = arrdict.arrdict(terminal=over, rewards=torch.stack([over.float(), -over.float()], -1))
        return type(self)(seats=1 - self.seats), trans
