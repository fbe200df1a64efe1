"""Generates tests/golden/chunk.npz by running the REFERENCE's own ``boardlaw.arena.neural.ChunkEvaluator`` / ``Tracker`` and
``boardlaw.hex.Lazy`` (build container only: needs /root/reference) on CPU with the deterministic agents of
tests/golden_checks.py.  The reference's own test for the evaluator (boardlaw/arena/neural.py:417-439) needs stored training runs
(pavlov) that are not shipped, so the fixture drives the same classes with position-determined agents instead.

    python tests/golden/make_golden_chunk.py
"""
import sys
import types
from pathlib import Path

import numpy as np
import pandas as pd
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests'))

from oracle import refpy  # noqa: E402
import golden_checks as gc  # noqa: E402
from make_golden_arena import load_reference_arena  # noqa: E402


def main():
    ref, common = load_reference_arena()
    # modules neural.py imports at the top and ChunkEvaluator / Tracker never touch: the Elo fitter, the process-pool helper (loky)
    for name in ('boardlaw.elos', 'rebar.parallel'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    import rebar
    rebar.parallel = sys.modules['rebar.parallel']
    import boardlaw
    boardlaw.elos = sys.modules['boardlaw.elos']
    import boardlaw.arena.neural as neural
    # pandas >= 2 hands out read-only `.values` (copy-on-write); Tracker.__init__ (neural.py:56-57, written for pandas 1) assigns into it
    _values = pd.DataFrame.values
    pd.DataFrame.values = property(lambda self: np.array(_values.fget(self)))
    out = {}
    # ---- ChunkEvaluator: every agent against every other, n_envs_per games per ordered pair
    for S, n_envs_per, mults in [(5, 6, (3, 7, 11)), (7, 4, (2, 5, 9, 13))]:
        agents = {f'kth{m}': gc.KthValid(m) for m in mults}

        def worldfunc(n):
            # diverse start positions: a few seeded random moves each (stored, so that the GPU run starts from the same boards)
            g = torch.Generator().manual_seed(100 + S)
            w = ref.Hex.initial(n, S, device='cpu')
            for _ in range(S):
                w, _ = w.step(torch.multinomial(w.valid.float(), 1, generator=g).squeeze(-1))
            out[f'S{S}_board'], out[f'S{S}_seats'] = w.board.numpy().copy(), w.seats.numpy().copy()
            return w
        ev = neural.ChunkEvaluator(worldfunc, agents, n_envs_per=n_envs_per, device='cpu')
        results, steps, masks = [], 0, []
        while not ev.finished():
            # (record what the tracker suggests at every step: the order of play is part of the contract)
            name, mask, live = ev.tracker.suggest(ev.worlds.seats)
            masks.append((list(agents).index(name), int(mask.sum())))
            results.extend(ev.step())
            steps += 1
        results.sort(key=lambda r: r.names)
        out[f'S{S}_names'] = np.array([list(r.names) for r in results])
        out[f'S{S}_wins'] = np.array([r.wins for r in results])
        out[f'S{S}_moves'] = np.array([r.moves for r in results])
        out[f'S{S}_games'] = np.array([r.games for r in results])
        out[f'S{S}_steps'] = np.array(steps)
        out[f'S{S}_suggest'] = np.array(masks)
        out[f'S{S}_n_envs'] = np.array(ev.tracker.n_envs)
        print(S, steps, [(r.names, r.wins, r.moves) for r in results][:4])
    # ---- Tracker alone, with games already played
    games = pd.DataFrame([[0, 2, 0], [1, 0, 3], [0, 0, 0]], ['a', 'b', 'c'], ['a', 'b', 'c'])
    tr = neural.Tracker(3, games, device='cpu')
    out['tracker_live'] = tr.live.numpy().copy()
    # ---- Lazy: one-player Hex against the first-legal-move opponent
    for S, B in [(5, 16), (9, 8)]:
        w = ref.hex.Lazy.initial(B, S, device='cpu')
        g = torch.Generator().manual_seed(200 + S)
        for _ in range(3):                                # desynchronise the envs: a few random moves of the player (the opponent stays lazy)
            w, _ = w.step(torch.multinomial(w.valid.float(), 1, generator=g).squeeze(-1))
        out[f'lazy_S{S}_board0'], out[f'lazy_S{S}_seats0'] = w.board.numpy().copy(), w.seats.numpy().copy()
        agent = gc.KthValid(5)
        boards, rewards, terminal, actions = [], [], [], []
        for _ in range(S * S // 2 + 8):
            a = agent(w).actions
            w, t = w.step(a)
            actions.append(a.numpy().copy()); boards.append(w.board.numpy().copy()); rewards.append(t.rewards.numpy().copy()); terminal.append(t.terminal.numpy().copy())
        out[f'lazy_S{S}_actions'], out[f'lazy_S{S}_boards'] = np.stack(actions), np.stack(boards)
        out[f'lazy_S{S}_rewards'], out[f'lazy_S{S}_terminal'] = np.stack(rewards), np.stack(terminal)
        print('lazy', S, float(np.stack(rewards).sum()), int(np.stack(terminal).sum()))
    np.savez_compressed(ROOT / 'tests' / 'golden' / 'chunk.npz', **out)


if __name__ == '__main__':
    main()
