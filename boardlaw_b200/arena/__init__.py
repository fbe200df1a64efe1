"""Head-to-head evaluation of two agents over a batch of games — ``boardlaw/arena/common.py:50-106`` on the B200 env.

Every env plays one game; which agent sits in which seat follows the seat permutations in turn (env e uses pattern
e mod n_seats!), so both agents play both colours equally often.  Each round, every agent moves in the envs where it
is to play and the game is still on; finished games are frozen (the env auto-resets on a win, the ``terminal`` mask
keeps the reset board out of play).

Differences from the reference's loop that do not change its results: the sub-batch is selected by an index tensor
computed once per agent per round (one device->host sync) instead of a boolean mask used five times (each use a
``nonzero`` + sync), and an ``MCTSAgent`` searching a shrinking sub-batch re-uses one workspace
(``SearchEngine._search_partial``) instead of allocating a tree per call.
"""
import math
import time
from itertools import permutations

import torch

from .. import arrdict


def matchup_patterns(n_seats):
    """Every assignment of agents to seats: row p, column s = the agent that plays seat s (boardlaw/arena/common.py:50-51)."""
    return torch.as_tensor(list(permutations(range(n_seats))))


def matchup_indices(n_envs, n_seats):
    """(n_envs, n_seats): the patterns repeated in order down the envs (boardlaw/arena/common.py:53-55)."""
    patterns = matchup_patterns(n_seats)
    return patterns.repeat((n_envs // len(patterns), 1))


def gather(wins, moves, times, matchup_idxs, agents, boardsize):
    """Per seat pattern: names in seat order, wins per seat, moves, games, seconds (boardlaw/arena/common.py:57-73)."""
    names = [name for name, _ in agents]
    results = []
    for p in matchup_patterns(matchup_idxs.shape[1]):
        rows = (matchup_idxs == p).all(-1)
        w = wins[rows].sum(0)
        results.append(arrdict.arrdict(
            names=tuple(names[int(i)] for i in p),
            wins=tuple(float(x) for x in w),
            moves=float(moves[rows].sum()),
            games=float(w.sum()),
            times=float(times[rows].sum()),
            boardsize=boardsize))
    return results


def evaluate(worlds, agents):
    """Plays every env of ``worlds`` to the end of one game; ``agents`` is a dict or list of (name, agent), one per seat.
    Returns ``gather``'s list (boardlaw/arena/common.py:75-106)."""
    if isinstance(agents, dict):
        agents = list(agents.items())
    if worlds.n_seats != 2:
        raise AssertionError('Only support 2 seats for now')
    if worlds.n_envs % math.factorial(worlds.n_seats) != 0:
        raise AssertionError('Number of envs needs to be divisible by the number of permutations of seats')
    if len(agents) != worlds.n_seats:
        raise AssertionError('Need to pass one agent per seat')

    dev, B = worlds.device, worlds.n_envs
    envs = torch.arange(B, device=dev)
    terminal = torch.zeros((B,), dtype=torch.bool, device=dev)
    wins = torch.zeros((B, worlds.n_seats), dtype=torch.int, device=dev)
    moves = torch.zeros((B,), dtype=torch.int, device=dev)
    times = torch.zeros((B,), dtype=torch.float, device=dev)
    matchup_idxs = matchup_indices(B, worlds.n_seats).to(dev)
    errors = torch.zeros((), dtype=torch.int32, device=dev)      # rule violations of every step of the evaluation (Hex.step)
    while True:
        for i, (_, agent) in enumerate(agents):
            idx = ((matchup_idxs[envs, worlds.seats.long()] == i) & ~terminal).nonzero().squeeze(-1)
            if idx.numel() == 0:
                continue
            start = time.time()
            sub = worlds[idx]
            sub.errors = errors
            decisions = agent(sub, eval=True)
            stepped, transitions = sub.step(decisions.actions)
            worlds[idx] = stepped
            terminal[idx] = transitions.terminal
            if dev.type == 'cuda':
                torch.cuda.synchronize(dev)                          # the clock below measures the move, not its launch
            end = time.time()

            wins[idx] += (transitions.rewards == 1).int()
            moves[idx] += 1
            times[idx] += (end - start) / idx.numel()
        if bool(terminal.all()):
            break
    if int(errors) != 0:
        raise AssertionError(f'an agent played an invalid action during the evaluation (error bits {int(errors):#x})')
    return gather(wins.cpu(), moves.cpu(), times.cpu(), matchup_idxs.cpu(), agents, worlds.boardsize)


from . import neural  # noqa: E402,F401  (boardlaw.arena.neural: the all-pairs ChunkEvaluator)
