"""TEST INFRASTRUCTURE — import the *Python* reference from ``/root/reference``.

Only works in the build container (the GPU box has no ``/root/reference``); it is used to
validate the restatements in ``oracle/`` and to generate the golden vectors committed under
``tests/golden/`` (``tests/golden/make_golden.py``).

The reference imports two third-party modules that are not installed here (``aljpy`` via
``rebar/profiling.py:2`` and ``matplotlib`` via ``boardlaw/hex/__init__.py:3-4``); neither touches
the arithmetic, so they are stubbed.  The native ops are the reference's own CPU extension
built from its unmodified sources by ``oracle/build_ref.py`` and planted in the loader caches
(``boardlaw/hex/cuda.py:4-9``, ``boardlaw/mcts/cuda.py:6-11``), which is exactly what the reference's
``boardlaw.cuda.load`` would build on a machine without CUDA.
"""
import sys
import types
import logging
import contextlib
from pathlib import Path

from . import build_ref

REFERENCE = build_ref.REFERENCE


def present():
    return (REFERENCE / 'boardlaw' / 'mcts' / '__init__.py').exists()


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_ref = None


def load(variant='O0'):
    """Returns a namespace with the reference's ``hex``, ``mcts``, ``networks``, ``heads``,
    ``arrdict`` modules, running on the reference's own CPU kernels (``variant`` O0/O2)."""
    global _ref
    if _ref is not None and _ref.variant == variant:
        return _ref
    if not present():
        raise FileNotFoundError('/root/reference is not present on this machine')
    build_ref.build()

    if 'aljpy' not in sys.modules:
        @contextlib.contextmanager
        def timer():
            yield None
        _stub('aljpy', logger=lambda *a, **k: logging.getLogger('ref'), timer=timer)
    if 'matplotlib' not in sys.modules:
        mpl = _stub('matplotlib')
        mpl.pyplot = _stub('matplotlib.pyplot')
        mpl.colors = _stub('matplotlib.colors')
        mpl.patches = _stub('matplotlib.patches')
        mpl.collections = _stub('matplotlib.collections')
    if 'pkg_resources' not in sys.modules:
        _stub('pkg_resources', resource_filename=lambda pkg, fn: str(REFERENCE / pkg.replace('.', '/') / fn))

    if str(REFERENCE) not in sys.path:
        sys.path.insert(0, str(REFERENCE))
    import boardlaw.hex as rhex
    import boardlaw.hex.cuda as rhexcuda
    import boardlaw.mcts as rmcts
    import boardlaw.mcts.cuda as rmctscuda
    import boardlaw.networks as rnetworks
    import boardlaw.heads as rheads
    from rebar import arrdict as rarrdict

    rhexcuda._cache = build_ref.load('hexcuda', variant)
    rmctscuda._cache = build_ref.load('mctscuda', variant)

    _ref = types.SimpleNamespace(
        variant=variant, hex=rhex, mcts=rmcts, networks=rnetworks, heads=rheads, arrdict=rarrdict,
        hexcuda=rhexcuda._cache, mctscuda=rmctscuda._cache,
        Hex=rhex.Hex, MCTS=rmcts.MCTS, MCTSAgent=rmcts.MCTSAgent, FCModel=rnetworks.FCModel)
    return _ref
