"""Generates tests/golden/arena_S5.npz by running the REFERENCE's own ``boardlaw.arena.common.evaluate`` (build container
only: needs /root/reference) on seeded Hex positions with the deterministic agents of tests/golden_checks.py.

    python tests/golden/make_golden_arena.py
"""
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests'))

from oracle import refpy  # noqa: E402
import golden_checks as gc  # noqa: E402


def load_reference_arena():
    """``boardlaw/arena/common.py`` imports the storage / SQL stack at module level (pavlov, sqlalchemy: absent here and
    unused by ``evaluate``); those modules are stubbed and the package ``__init__`` (which pulls in the live arena) skipped."""
    ref = refpy.load()
    for name in ('pavlov', 'pavlov.storage', 'pavlov.runs', 'pavlov.stats', 'boardlaw.sql', 'boardlaw.backup'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['pavlov'].storage, sys.modules['pavlov'].runs = sys.modules['pavlov.storage'], sys.modules['pavlov.runs']
    import boardlaw
    boardlaw.sql, boardlaw.backup = sys.modules['boardlaw.sql'], sys.modules['boardlaw.backup']
    pkg = types.ModuleType('boardlaw.arena')
    pkg.__path__ = [str(refpy.REFERENCE / 'boardlaw' / 'arena')]
    sys.modules['boardlaw.arena'] = pkg
    if not hasattr(np, 'math'):
        import math
        np.math = math                   # common.py:79 uses np.math.factorial, gone from numpy 2
    import boardlaw.arena.common as common
    return ref, common


def main():
    ref, common = load_reference_arena()
    out = {}
    for S, B, n_random, seed in [(5, 64, 6, 0), (7, 32, 15, 1)]:
        g = torch.Generator().manual_seed(seed)
        worlds = ref.Hex.initial(B, S, device='cpu')
        for _ in range(n_random):
            a = torch.multinomial(worlds.valid.float(), 1, generator=g).squeeze(-1)
            worlds, _ = worlds.step(a)
        out[f'S{S}_board'], out[f'S{S}_seats'] = worlds.board.numpy().copy(), worlds.seats.numpy().copy()
        results = common.evaluate(worlds, [('kth3', gc.KthValid(3)), ('kth7', gc.KthValid(7))])
        out[f'S{S}_names'] = np.array([list(r.names) for r in results])
        out[f'S{S}_wins'] = np.array([r.wins for r in results])
        out[f'S{S}_moves'] = np.array([r.moves for r in results])
        out[f'S{S}_games'] = np.array([r.games for r in results])
        print(S, [dict(names=r.names, wins=r.wins, moves=r.moves, games=r.games) for r in results])
    # The reference's own known-answer test (boardlaw/arena/common.py:108-116: WinnerLoser + RandomAgent, each pattern's wins
    # == (2, 0)) cannot be run as shipped: evaluate passes eval=True, which validation.RandomAgent does not accept, and reads
    # worlds.boardsize, which WinnerLoser does not have.  tests/test_arena.py asserts its intended result on our implementation.
    np.savez_compressed(ROOT / 'tests' / 'golden' / 'arena.npz', **out)


if __name__ == '__main__':
    main()
