"""The drop-in boundary at op level — the counterpart of ``boardlaw/cuda.py:48-63``.

The reference JIT-builds two pybind11 modules (``hexcuda``, ``mctscuda``) and caches them in
``boardlaw.hex.cuda._cache`` / ``boardlaw.mcts.cuda._cache``.  ``load`` returns objects exposing the same
names, backed by ``libboardlaw_b200.so``; ``install`` plants them in a boardlaw checkout's caches so the
reference's unchanged ``Hex`` / ``MCTS`` Python runs on these kernels.
"""
from . import _lib


def load(pkg, files=None):
    """``pkg`` is 'boardlaw.hex' / 'boardlaw.mcts' (or this package's equivalents)."""
    _lib.lib()
    leaf = pkg.split('.')[-1]
    if leaf == 'hex':
        from .hex import cuda as m
        return m.hexcuda
    if leaf == 'mcts':
        from .mcts import cuda as m
        return m.mctscuda
    raise ValueError(f'No kernels for package {pkg}')


def install():
    """Pre-seeds the reference's loader caches (boardlaw/hex/cuda.py:4-9, boardlaw/mcts/cuda.py:6-11)."""
    import boardlaw.hex.cuda as rhex
    import boardlaw.mcts.cuda as rmcts
    rhex._cache = load('boardlaw.hex')
    rmcts._cache = load('boardlaw.mcts')


def assert_shape(x, s):
    assert (x.ndim == len(s)) and x.shape == s, f'Expected {s}, got {x.shape}'
    assert x.device.type == 'cuda', f'Expected CUDA tensor, got {x.device.type}'
