"""boardlaw_b200 — a B200-native (sm_100a) vectorised MCTS self-play engine for Hex.

Keeps the ``boardlaw.hex.Hex`` env API, the ``boardlaw.mcts.MCTS`` / ``MCTSAgent`` surface and the
``boardlaw.networks.FCModel`` state-dict layout of andyljones/boardlaw; the hot loops run in hand-written
CUDA behind the C ABI of ``include/boardlaw_b200.h``.  There is no CPU path.
"""
from . import arrdict  # noqa: F401

__all__ = ['arrdict', 'hex', 'mcts', 'networks', 'heads', 'cuda']
