"""``hexcuda``-shaped ops on libboardlaw_b200.so — mirrors ``boardlaw/hex/cuda.py`` and the pybind module
of ``boardlaw/hex/cpp/wrappers.cpp:37-41``."""
import types

import torch

from .. import _lib
from .._lib import proxy, ptr, check


def step(board, seats, actions):
    """In place on ``board`` (B,S,S) u8; returns rewards (B,2) f32.  boardlaw/hex/cpp/cuda.cu:139-152."""
    proxy(board, torch.uint8, 3, 'board'); proxy(seats, torch.int32, 1, 'seats'); proxy(actions, torch.int32, 1, 'actions')
    dev = _lib.require_cuda(board, seats, actions)
    B, S, _ = board.shape
    rewards = board.new_empty((B, 2), dtype=torch.float32)
    check(_lib.lib().bl_hex_step(ptr(board), ptr(seats), ptr(actions), ptr(rewards), B, S, _lib.stream_for(dev)),
          'bl_hex_step')
    return rewards


def observe(board, seats):
    """obs (...,S,S,2) f32 for any leading batch shape.  boardlaw/hex/cpp/cuda.cu:197-217."""
    dev = _lib.require_cuda(board, seats)
    S = board.shape[-1]
    flat = proxy(board.contiguous().reshape(-1, S, S), torch.uint8, 3, 'board')
    fseats = seats.contiguous().reshape(-1)
    if fseats.dtype != torch.int32:
        fseats = fseats.to(torch.int32)
    if fseats.shape[0] != flat.shape[0]:
        raise RuntimeError('board and seats batch shapes differ')
    obs = board.new_empty((flat.shape[0], S, S, 2), dtype=torch.float32)
    check(_lib.lib().bl_hex_observe(ptr(flat), ptr(fseats), ptr(obs), flat.shape[0], S, _lib.stream_for(dev)),
          'bl_hex_observe')
    return obs.reshape(*board.shape, 2)


def valid(board, seats):
    """(...,A) bool mask of empty cells in the mover's frame, without materialising obs."""
    dev = _lib.require_cuda(board, seats)
    S = board.shape[-1]
    flat = proxy(board.contiguous().reshape(-1, S, S), torch.uint8, 3, 'board')
    fseats = seats.contiguous().reshape(-1).to(torch.int32)
    out = board.new_empty((flat.shape[0], S * S), dtype=torch.bool)
    check(_lib.lib().bl_hex_valid(ptr(flat), ptr(fseats), ptr(out), flat.shape[0], S, _lib.stream_for(dev)),
          'bl_hex_valid')
    return out.reshape(*board.shape[:-2], S * S)


def transition(board, seats, actions, reset=True, error_word=None):
    """Fused ``Hex.step`` (boardlaw/hex/__init__.py:161-195): returns (new_board, new_seats, rewards, terminal).
    Rule violations are OR-ed into ``error_word`` (int32 scalar tensor) instead of asserted on the host."""
    proxy(board, torch.uint8, 3, 'board'); proxy(seats, torch.int32, 1, 'seats'); proxy(actions, torch.int64, 1, 'actions')
    dev = _lib.require_cuda(board, seats, actions, error_word)
    B, S, _ = board.shape
    new_board = torch.empty_like(board)
    new_seats = torch.empty_like(seats)
    rewards = board.new_empty((B, 2), dtype=torch.float32)
    terminal = board.new_empty((B,), dtype=torch.bool)
    if error_word is None:
        error_word = board.new_zeros((), dtype=torch.int32)
    check(_lib.lib().bl_hex_transition(ptr(board), ptr(seats), ptr(actions), ptr(new_board), ptr(new_seats), ptr(rewards),
                                       ptr(terminal), ptr(error_word), int(bool(reset)), B, S, _lib.stream_for(dev)),
          'bl_hex_transition')
    return new_board, new_seats, rewards, terminal


def random_transition(board, seats, uniforms, reset=True, error_word=None):
    """One fused step of a uniformly random playout (the loop body of ``learning.mix``, boardlaw/learning.py:6-10): env b
    plays its k-th legal move, k = floor(uniforms[b] * n_legal).  Returns (new_board, new_seats, actions, rewards, terminal)."""
    proxy(board, torch.uint8, 3, 'board'); proxy(seats, torch.int32, 1, 'seats'); proxy(uniforms, torch.float32, 1, 'uniforms')
    dev = _lib.require_cuda(board, seats, uniforms, error_word)
    B, S, _ = board.shape
    if uniforms.shape[0] != B:
        raise RuntimeError('uniforms must have one entry per env')
    new_board = torch.empty_like(board)
    new_seats = torch.empty_like(seats)
    actions = board.new_empty((B,), dtype=torch.int64)
    rewards = board.new_empty((B, 2), dtype=torch.float32)
    terminal = board.new_empty((B,), dtype=torch.bool)
    if error_word is None:
        error_word = board.new_zeros((), dtype=torch.int32)
    check(_lib.lib().bl_hex_random_transition(ptr(board), ptr(seats), ptr(uniforms), ptr(new_board), ptr(new_seats), ptr(actions),
                                              ptr(rewards), ptr(terminal), ptr(error_word), int(bool(reset)), B, S,
                                              _lib.stream_for(dev)), 'bl_hex_random_transition')
    return new_board, new_seats, actions, rewards, terminal


# the object planted in ``boardlaw.hex.cuda._cache``
hexcuda = types.SimpleNamespace(step=step, observe=observe)


def module():
    return hexcuda
