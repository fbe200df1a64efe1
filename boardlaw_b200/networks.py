"""``FCModel`` — the policy/value network of ``boardlaw/networks.py:20-41`` with the heads of
``boardlaw/heads.py`` (TensorIntake :41-52, MaskedOutput :93-104, ValueOutput :128-142), evaluated by the
fused forward kernels of libboardlaw_b200.so straight from ``worlds.board`` / ``worlds.seats``.

The module tree reproduces the reference's ``state_dict`` keys exactly
(``body.0.{weight,bias}``, ``body.k.{weight,bias,α}``, ``policy.core.*``, ``value.core.*``), so snapshots
load either way.  Initialisation follows the reference: orthogonal gain sqrt(2) residual weights with
α = 0 (boardlaw/networks.py:12-15), torch defaults elsewhere.
"""
import ctypes

import numpy as np
import torch
from torch import nn

from . import _lib, arrdict
from ._lib import ptr, check


class ReZeroResidual(nn.Linear):
    """x + α·Linear(relu(x)) (boardlaw/networks.py:10-18); holds parameters only, the math is in the kernel."""

    def __init__(self, width):
        super().__init__(width, width)
        nn.init.orthogonal_(self.weight, gain=2**.5)
        self.register_parameter('α', nn.Parameter(torch.zeros(())))


class _Head(nn.Module):

    def __init__(self, width, n_out):
        super().__init__()
        self.core = nn.Linear(width, n_out)


class FCModel(nn.Module):

    def __init__(self, obs_space, action_space, width=256, depth=64, precision='fp32'):
        super().__init__()
        self.obs_space = obs_space
        self.action_space = action_space
        self.width, self.depth = width, depth
        self.precision = precision

        n_actions = int(np.prod(action_space.dim)) if hasattr(action_space, 'dim') else int(np.prod(action_space))
        n_obs = int(np.prod(obs_space.dim))
        assert n_obs == 2 * n_actions, 'FCModel here is specialised to the Hex observation (S,S,2)'
        self.boardsize = int(round(n_actions ** .5))

        self.policy = _Head(width, n_actions)
        blocks = [nn.Linear(n_obs, width)]
        for _ in range(depth):
            blocks.append(ReZeroResidual(width))
        self.body = nn.Sequential(*blocks)
        self.value = _Head(width, 1)

        self._pack_key = None
        self._pack = None

    def sampler(self, logits, test=False):
        if test:
            return logits.argmax(-1)
        return torch.distributions.Categorical(logits=logits).sample()

    # ---- weight staging -----------------------------------------------------------------------------
    def packed(self):
        """Contiguous fp32 device tensors in the order bl_fc_params wants, rebuilt when a parameter changes."""
        params = list(self.parameters())
        key = tuple((p.data_ptr(), p._version, p.device) for p in params)
        if key != self._pack_key:
            f = lambda t: t.detach().float().contiguous()
            res = list(self.body)[1:]
            dev = self.body[0].weight.device
            W = self.width
            pack = dict(
                w_in=f(self.body[0].weight), b_in=f(self.body[0].bias),
                w_res=torch.stack([f(r.weight) for r in res]) if res else torch.zeros((0, W, W), device=dev),
                b_res=torch.stack([f(r.bias) for r in res]) if res else torch.zeros((0, W), device=dev),
                alpha=torch.stack([f(getattr(r, 'α')) for r in res]) if res else torch.zeros((0,), device=dev),
                w_pol=f(self.policy.core.weight), b_pol=f(self.policy.core.bias),
                w_val=f(self.value.core.weight).reshape(-1), b_val=f(self.value.core.bias).reshape(-1))
            cp = _lib.FCParams(
                S=self.boardsize, W=W, D=len(res), precision=0 if self.precision == 'fp32' else 1,
                **{k: t.data_ptr() for k, t in pack.items()}, packed=None)
            self._pack, self._pack_key, self._cparams = pack, key, cp
        return self._cparams

    # ---- forward --------------------------------------------------------------------------------------
    def evaluate(self, board, seats):
        """board (B,S,S) u8, seats (B,) i32 -> (logits (B,A) f32, v (B,2) f32)."""
        dev = _lib.require_cuda(board, seats, self.body[0].weight)
        board = _lib.proxy(board.contiguous(), torch.uint8, 3, 'board')
        seats = _lib.proxy(seats.contiguous().int(), torch.int32, 1, 'seats')
        B, S, _ = board.shape
        assert S == self.boardsize
        cp = self.packed()
        l = _lib.lib()
        nbytes = l.bl_fc_scratch_bytes(ctypes.byref(cp), B)
        scratch = torch.empty((max(int(nbytes), 1),), dtype=torch.uint8, device=dev)
        logits = torch.empty((B, S * S), dtype=torch.float32, device=dev)
        v = torch.empty((B, 2), dtype=torch.float32, device=dev)
        check(l.bl_fc_forward(ctypes.byref(cp), ptr(board), ptr(seats), ptr(logits), ptr(v), ptr(scratch), B,
                              _lib.stream_for(dev)), 'bl_fc_forward')
        return logits, v

    def forward(self, worlds):
        """``network(worlds) -> arrdict(logits, v)`` (boardlaw/networks.py:37-41).  Inference only."""
        with torch.no_grad():
            logits, v = self.evaluate(worlds.board, worlds.seats)
        return arrdict.arrdict(logits=logits, v=v)


def synthetic_state_dict(boardsize, width, depth, seed=0):
    """Random-init FCModel weights for benchmarks (no checkpoints exist offline): torch-default-style uniform fan-in
    init for the plain layers, gaussian sqrt(2/W) residual weights, and α ~ U(0.1, 0.5) in every ReZero block —
    at the reference's own init α = 0 (boardlaw/networks.py:15) the residual weights would not contribute at all."""
    g = torch.Generator().manual_seed(seed)
    A = boardsize * boardsize

    def lin(o, i):
        bound = 1 / i ** .5
        return ((torch.rand((o, i), generator=g) * 2 - 1) * bound, (torch.rand((o,), generator=g) * 2 - 1) * bound)

    sd = {}
    sd['body.0.weight'], sd['body.0.bias'] = lin(width, 2 * A)
    for k in range(1, depth + 1):
        sd[f'body.{k}.weight'] = torch.randn((width, width), generator=g) * (2 / width) ** .5
        sd[f'body.{k}.bias'] = (torch.rand((width,), generator=g) * 2 - 1) / width ** .5
        sd[f'body.{k}.α'] = torch.rand((), generator=g) * .4 + .1
    sd['policy.core.weight'], sd['policy.core.bias'] = lin(A, width)
    sd['value.core.weight'], sd['value.core.bias'] = lin(1, width)
    return sd
