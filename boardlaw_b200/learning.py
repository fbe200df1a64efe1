"""The random-playout utility of ``boardlaw/learning.py:6-10`` on the fused kernel."""
import torch


def mix(worlds, T=2500, generator=None):
    """``T`` uniformly random legal moves on every env (finished games auto-reset), one kernel per move —
    ``boardlaw.learning.mix`` (boardlaw/learning.py:6-10), which the reference uses to decorrelate fresh worlds."""
    for _ in range(T):
        worlds, _ = worlds.step_random(generator=generator)
    return worlds
