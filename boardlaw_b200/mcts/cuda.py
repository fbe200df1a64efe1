"""``mctscuda``-shaped ops on libboardlaw_b200.so — mirrors ``boardlaw/mcts/cuda.py`` and the pybind module of
``boardlaw/mcts/cpp/wrappers.cpp:50-74`` (classes ``MCTS``, ``Backup``, ``Descent``; functions ``descend``,
``root``, ``backup``)."""
import types

import torch

from .. import _lib
from .._lib import proxy, ptr, check
from ..cuda import assert_shape


class MCTS:
    """Argument bundle, as ``mctscuda.MCTS`` (boardlaw/mcts/cpp/common.h:25-55): dtype/rank/contiguity checked
    at construction, tensors exposed read-only."""

    def __init__(self, logits, w, n, c_puct, seats, terminal, children):
        self.logits = proxy(logits, torch.float16, 3, 'logits')
        self.w = proxy(w, torch.float16, 3, 'w')
        self.n = proxy(n, torch.int16, 2, 'n')
        self.c_puct = proxy(c_puct, torch.float16, 1, 'c_puct')
        self.seats = proxy(seats, torch.int16, 2, 'seats')
        self.terminal = proxy(terminal, torch.bool, 2, 'terminal')
        self.children = proxy(children, torch.int16, 3, 'children')
        self.device = _lib.require_cuda(logits, w, n, c_puct, seats, terminal, children)


class Backup:
    """As ``mctscuda.Backup`` (boardlaw/mcts/cpp/common.h:85-130)."""

    def __init__(self, v, w, n, rewards, parents, terminal):
        self.v = proxy(v, torch.float16, 3, 'v')
        self.w = proxy(w, torch.float16, 3, 'w')
        self.n = proxy(n, torch.int16, 2, 'n')
        self.rewards = proxy(rewards, torch.float16, 3, 'rewards')
        self.parents = proxy(parents, torch.int16, 2, 'parents')
        self.terminal = proxy(terminal, torch.bool, 2, 'terminal')
        self.device = _lib.require_cuda(v, w, n, rewards, parents, terminal)


class Descent:

    def __init__(self, parents, actions):
        self.parents, self.actions = parents, actions


_counters = {}


def counters(device):
    """Device counters (policy evaluations, existing children seen, Newton iterations, descents) the descend
    kernels keep for the roofline accounting of SURVEY.md §8(d)."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _counters:
        _counters[key] = torch.zeros(8, dtype=torch.int64, device=device)
    return _counters[key]


def mcts(logits, w, n, c_puct, seats, terminal, children):
    """boardlaw/mcts/cuda.py:13-26 — shape checks, then the bundle.  ``c_puct > 0`` is asserted there with a
    host sync on every call; here it is the caller's contract (checked when hex.STRICT is set)."""
    B, T, A = logits.shape
    S = w.shape[-1]
    assert_shape(w, (B, T, S))
    assert_shape(n, (B, T))
    assert_shape(c_puct, (B,))
    assert_shape(seats, (B, T))
    assert_shape(terminal, (B, T))
    assert_shape(children, (B, T, A))
    from ..hex import STRICT
    if STRICT:
        assert (c_puct > 0.).all(), 'Zero c_puct not supported'
    return MCTS(logits, w, n, c_puct, seats.short() if seats.dtype != torch.int16 else seats, terminal, children)


_next_rands = []


def inject_rands(rands):
    """The next ``descend`` call without a ``rands`` argument uses this (B,T) half tensor instead of drawing one: lets a caller
    that cannot pass the argument (the reference's ``MCTS.descend`` calls ``descend(m)``) run a seeded comparison across devices."""
    _next_rands.append(rands)


def descend(m, rands=None):
    """``mctscuda.descend`` (boardlaw/mcts/cpp/cuda.cu:184-203).  ``rands`` (B,T) half is drawn here with
    ``torch.rand_like`` exactly where the reference draws it, unless injected."""
    B, T, A = m.logits.shape
    Sn = m.w.shape[2]
    if rands is None and _next_rands:
        rands = _next_rands.pop(0)
    if rands is None:
        rands = torch.rand_like(m.logits[:, :, 0])
    rands = proxy(rands.contiguous(), torch.float16, 2, 'rands')
    parents = m.seats.new_empty((B,))
    actions = m.seats.new_empty((B,))
    qrange = m.logits.new_empty((2,), dtype=torch.float32)
    check(_lib.lib().bl_mcts_descend(
        ptr(m.logits), ptr(m.w), ptr(m.n), ptr(m.c_puct), ptr(m.seats), ptr(m.terminal), ptr(m.children),
        ptr(rands), ptr(_lib.exp_lut(m.device)), ptr(qrange), ptr(parents), ptr(actions), ptr(counters(m.device)),
        B, T, A, Sn, _lib.stream_for(m.device)), 'bl_mcts_descend')
    return Descent(parents, actions)


def root(m):
    """``mctscuda.root`` (boardlaw/mcts/cpp/cuda.cu:120-136): (B,A) half probabilities at node 0."""
    B, T, A = m.logits.shape
    probs = m.logits.new_empty((B, A))
    qrange = m.logits.new_empty((2,), dtype=torch.float32)
    check(_lib.lib().bl_mcts_root(
        ptr(m.logits), ptr(m.w), ptr(m.n), ptr(m.c_puct), ptr(m.seats), ptr(m.terminal), ptr(m.children),
        ptr(_lib.exp_lut(m.device)), ptr(qrange), ptr(probs), B, T, A, m.w.shape[2], _lib.stream_for(m.device)),
        'bl_mcts_root')
    return probs


def backup(bk, leaves):
    """``mctscuda.backup`` (boardlaw/mcts/cpp/cuda.cu:238-248): in place on ``bk.n`` and ``bk.w``."""
    proxy(leaves, torch.int16, 1, 'leaves')
    _lib.require_cuda(leaves, bk.v)
    B, T, Sn = bk.v.shape
    check(_lib.lib().bl_mcts_backup(ptr(bk.v), ptr(bk.w), ptr(bk.n), ptr(bk.rewards), ptr(bk.parents), ptr(bk.terminal),
                                    ptr(leaves), B, T, Sn, _lib.stream_for(bk.device)), 'bl_mcts_backup')


def transition_q(m):
    """transition_q (boardlaw/mcts/cpp/cuda.cu:101-105) on its own: (q (B,T,Sn) half, (lo, hi) f32)."""
    B, T, Sn = m.w.shape
    q = torch.empty_like(m.w)
    qrange = m.logits.new_empty((2,), dtype=torch.float32)
    check(_lib.lib().bl_mcts_transition_q(ptr(m.w), ptr(m.n), ptr(q), ptr(qrange), B, T, Sn,
                                          _lib.stream_for(m.device)), 'bl_mcts_transition_q')
    ordered = qrange.view(torch.int32)
    bits = torch.where(ordered >= 0, ordered, ordered ^ 0x7FFFFFFF)
    return q, bits.view(torch.float32)


mctscuda = types.SimpleNamespace(MCTS=MCTS, Backup=Backup, Descent=Descent, descend=descend, root=root, backup=backup)


def module():
    return mctscuda
