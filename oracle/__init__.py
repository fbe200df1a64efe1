"""TEST INFRASTRUCTURE — CPU oracle for the boardlaw hot path.

``oracle/boardlaw_oracle.c`` restates the reference's CPU algorithms in plain C; this module
builds it with gcc, loads it with ctypes and wraps it in functions that take CPU torch tensors
with the reference's op signatures (``boardlaw/hex/cpp/wrappers.cpp:39-40``,
``boardlaw/mcts/cpp/wrappers.cpp:52-73``).

Nothing under ``boardlaw_b200/`` imports this package.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()``, and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.
"""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
SRC = HERE / 'boardlaw_oracle.c'
LIB = HERE / '_build' / 'libboardlaw_oracle.so'

_lib = None


def build(force=False):
    if LIB.exists() and not force and LIB.stat().st_mtime >= SRC.stat().st_mtime:
        return LIB
    LIB.parent.mkdir(exist_ok=True)
    cmd = ['gcc', '-O2', '-ffp-contract=off', '-fno-fast-math', '-fPIC', '-shared', '-std=c11',
           '-o', str(LIB), str(SRC), '-lm']
    subprocess.run(cmd, check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(LIB))
        _lib.orc_h2f.restype = ctypes.c_float
        _lib.orc_h2f.argtypes = [ctypes.c_uint16]
        _lib.orc_f2h.restype = ctypes.c_uint16
        _lib.orc_f2h.argtypes = [ctypes.c_float]
    return _lib


def _p(t):
    """Raw pointer of a contiguous CPU tensor (or None)."""
    if t is None:
        return ctypes.c_void_p(0)
    assert t.device.type == 'cpu' and t.is_contiguous(), 'oracle takes contiguous CPU tensors'
    return ctypes.c_void_p(t.data_ptr())


def _check(t, dtype, ndim, name):
    if t.dtype != dtype:
        raise TypeError(f'{name}: expected {dtype} got {t.dtype}')
    if t.ndim != ndim:
        raise RuntimeError(f'{name}: expected {ndim} dims got {t.ndim}')
    if not t.is_contiguous():
        raise RuntimeError(f'{name} must be contiguous')


# --------------------------------------------------------------------------------------------
# hexcuda-shaped ops
# --------------------------------------------------------------------------------------------

def step(board, seats, actions):
    """``hexcuda.step`` (boardlaw/hex/cpp/cpu.cpp:132-143): in-place on board, returns rewards (B,2) f32."""
    _check(board, torch.uint8, 3, 'board'); _check(seats, torch.int32, 1, 'seats'); _check(actions, torch.int32, 1, 'actions')
    B, S, _ = board.shape
    rewards = torch.zeros((B, 2), dtype=torch.float32)
    lib().orc_hex_step(_p(board), _p(seats), _p(actions), _p(rewards), B, S)
    return rewards


def observe(board, seats):
    """``hexcuda.observe`` (boardlaw/hex/cpp/cpu.cpp:183-203): any leading batch shape."""
    S = board.shape[-1]
    flat = board.contiguous().reshape(-1, S, S)
    fseats = seats.contiguous().reshape(-1).to(torch.int32)
    obs = torch.empty((flat.shape[0], S, S, 2), dtype=torch.float32)
    lib().orc_hex_observe(_p(flat), _p(fseats), _p(obs), flat.shape[0], S)
    return obs.reshape(*board.shape, 2)


# --------------------------------------------------------------------------------------------
# mctscuda-shaped ops
# --------------------------------------------------------------------------------------------

class MCTS:
    """Argument bundle, as ``mctscuda.MCTS`` (boardlaw/mcts/cpp/common.h:25-55)."""

    def __init__(self, logits, w, n, c_puct, seats, terminal, children):
        _check(logits, torch.float16, 3, 'logits'); _check(w, torch.float16, 3, 'w')
        _check(n, torch.int16, 2, 'n'); _check(c_puct, torch.float16, 1, 'c_puct')
        _check(seats, torch.int16, 2, 'seats'); _check(terminal, torch.bool, 2, 'terminal')
        _check(children, torch.int16, 3, 'children')
        self.logits, self.w, self.n, self.c_puct = logits, w, n, c_puct
        self.seats, self.terminal, self.children = seats, terminal, children


class Backup:
    """As ``mctscuda.Backup`` (boardlaw/mcts/cpp/common.h:103-130)."""

    def __init__(self, v, w, n, rewards, parents, terminal):
        _check(v, torch.float16, 3, 'v'); _check(w, torch.float16, 3, 'w'); _check(n, torch.int16, 2, 'n')
        _check(rewards, torch.float16, 3, 'rewards'); _check(parents, torch.int16, 2, 'parents')
        _check(terminal, torch.bool, 2, 'terminal')
        self.v, self.w, self.n, self.rewards, self.parents, self.terminal = v, w, n, rewards, parents, terminal


class Descent:
    def __init__(self, parents, actions):
        self.parents, self.actions = parents, actions


def transition_q(m, return_range=False):
    B, T, Sn = m.w.shape
    q = torch.empty_like(m.w)
    mm = torch.empty(2, dtype=torch.float32)
    lib().orc_transition_q(_p(m.w), _p(m.n), _p(q), _p(mm), B, T, Sn)
    return (q, mm) if return_range else q


def descend(m, rands=None, pow_mode=0, counters=None):
    """``mctscuda.descend``.  ``rands`` (B,T) half may be injected; otherwise drawn from torch's
    global CPU generator exactly where the reference draws it (boardlaw/mcts/cpp/cpu.cpp:187)."""
    B, T, A = m.logits.shape
    q = transition_q(m)
    if rands is None:
        rands = torch.rand_like(m.logits[:, :, 0])
    rands = rands.contiguous()
    _check(rands, torch.float16, 2, 'rands')
    parents = torch.empty(B, dtype=torch.int16)
    actions = torch.empty(B, dtype=torch.int16)
    cnt = None
    if counters is not None:
        _check(counters, torch.int64, 1, 'counters')
        cnt = counters
    lib().orc_descend(_p(m.logits), _p(q), _p(m.n), _p(m.c_puct), _p(m.seats), _p(m.terminal),
                      _p(m.children), _p(rands), _p(parents), _p(actions), B, T, A, m.w.shape[2],
                      int(pow_mode), _p(cnt))
    return Descent(parents, actions)


def root(m, pow_mode=0):
    B, T, A = m.logits.shape
    q = transition_q(m)
    probs = torch.empty((B, A), dtype=torch.float16)
    lib().orc_root(_p(m.logits), _p(q), _p(m.n), _p(m.c_puct), _p(m.seats), _p(m.terminal),
                   _p(m.children), _p(probs), B, T, A, m.w.shape[2], int(pow_mode))
    return probs


def backup(bk, leaves):
    _check(leaves, torch.int16, 1, 'leaves')
    B, T, Sn = bk.v.shape
    lib().orc_backup(_p(bk.v), _p(bk.w), _p(bk.n), _p(bk.rewards), _p(bk.parents), _p(bk.terminal),
                     _p(leaves), B, T, Sn)


def exp_table():
    """expf of every fp16 bit pattern, evaluated by this host's libm."""
    out = np.empty(65536, dtype=np.float32)
    lib().orc_exp_table(out.ctypes.data_as(ctypes.c_void_p))
    return out
