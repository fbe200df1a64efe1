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
        self._pack_gen = 0          # bumped whenever the staged operands move to new device addresses (captured graphs key on it)
        # False routes every shape through the CUDA-core fp32 kernels (net.cu) — used as the on-device cross-check
        self.tensor_cores = True
        self.token = next(_TOKENS)
        # tile order of the packed operands (bl_fc_params.tc_nsplit); BL_TC_NSPLIT overrides for experiments
        import os
        self.tc_nsplit = 1      # whole-N tiles: one tcgen05.mma issue costs ~120 cycles, so N = W/2 tiles are issue-bound

    def sampler(self, logits, test=False):
        if test:
            return logits.argmax(-1)
        return torch.distributions.Categorical(logits=logits).sample()

    # ---- weight staging -----------------------------------------------------------------------------
    def packed(self):
        """Contiguous fp32 device tensors in the order bl_fc_params wants, rebuilt when a parameter changes."""
        params = list(self.parameters())
        key = (self.precision, self.tensor_cores, self.tc_nsplit) + tuple((p.data_ptr(), p._version, p.device) for p in params)
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
            tc = {}
            if self.tensor_cores and W in (32, 64, 128, 256, 512) and self.boardsize ** 2 + 1 <= 256:
                # W = 512 runs net_tc_wide.cu: 16-wide K chunks, whole-N tiles
                blob, b_head = pack_tensor_core_operands(pack, self.boardsize, self.tc_nsplit, kc=16 if W == 512 else KC, wide=W == 512)
                pack['packed'], pack['b_head'] = blob, b_head
                tc = dict(packed=blob.data_ptr(), b_head=b_head.data_ptr(), tc_nsplit=self.tc_nsplit)
            if self._pack is not None and getattr(self, '_pack_cfg', None) == key[:3] and self._pack.keys() == pack.keys() and all(
                    self._pack[k].shape == v.shape and self._pack[k].device == v.device for k, v in pack.items()):
                # same shapes as the staged set (a weight update): refresh IN PLACE, so that the addresses baked into bl_fc_params —
                # and into any CUDA graph captured with them — stay valid and serve the new weights
                for k, v in pack.items():
                    self._pack[k].copy_(v)
                self._pack_key = key
            else:
                cp = _lib.FCParams(
                    S=self.boardsize, W=W, D=len(res), precision=0 if self.precision == 'fp32' else 1,
                    **{k: t.data_ptr() for k, t in pack.items() if k not in ('packed', 'b_head')}, **tc)
                self._pack, self._pack_key, self._cparams = pack, key, cp
                self._pack_cfg = key[:3]
                self._pack_gen += 1
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


import itertools
_TOKENS = itertools.count(1)     # identity of an FCModel for caches that outlive it (id() values are reused by CPython)

KC = 32     # K elements per operand tile (net_tc.cu)


def _split_tiles(w, n_pad, k_pad, KC=KC):
    """(N,K) fp32 -> (k_pad/KC, 2, n_pad/8, KC/8, 8, 8) fp16: per K-chunk a hi block then a lo block, each in the UMMA
    canonical K-major no-swizzle layout: element (n, k) at ((n/8)*(KC/8) + k/8)*64 + (n%8)*8 + k%8 halves."""
    N, K = w.shape
    full = w.new_zeros((n_pad, k_pad))
    full[:N, :K] = w
    hi = full.half()
    lo = (full - hi.float()).half()
    t = torch.stack([hi, lo])                                            # (2, n_pad, k_pad)
    t = t.reshape(2, n_pad // 8, 8, k_pad // KC, KC // 8, 8)             # (2, n/8, n%8, chunk, k/8, k%8)
    return t.permute(3, 0, 1, 4, 2, 5).contiguous()                      # (chunk, 2, n/8, k/8, n%8, k%8)


def pack_tensor_core_operands(pack, boardsize, nsplit=1, kc=KC, wide=False):
    """The weight blob fc_tc_kernel streams, in the kernel's consumption order (net_tc.cu): per layer the four blocks
    (N half 0, K half 0), (N half 1, K half 0), (N half 0, K half 1), (N half 1, K half 1) — one block when W < 64 — then the
    fused head [policy ; value]; every weight split as hi = fp16(w), lo = fp16(w - hi).  The ReZero gate is folded into the
    residual weights (alpha_k W_k), and the biases are returned as the cumulative vectors c_0 = b_in, c_k = c_{k-1} +
    alpha_k b_k the kernel adds on the way out of the accumulator, followed by the head bias.
    ``wide`` (net_tc_wide.cu, W = 512): 16-wide K chunks, and per layer all the chunks of N half 0 (rows 0..255), then all of half 1."""
    A = boardsize * boardsize
    W = pack['w_in'].shape[0]
    k0p = (2 * A + kc - 1) // kc * kc
    n_p = (A + 1 + 31) // 32 * 32
    nh = 2 if (nsplit == 2 and W >= 64) else 1
    wh = W // nh
    _split = lambda w, n_pad, k_pad: _split_tiles(w, n_pad, k_pad, kc)

    def body(w, k_pad):
        if wide:
            return [_split(w[h * 256:(h + 1) * 256], 256, k_pad).reshape(-1) for h in range(W // 256)]
        nk = k_pad // kc
        ks = ((nk + 1) // 2 if k_pad == k0p and w is pack['w_in'] else nk // 2) if nh == 2 else nk
        tiles = [_split(w[h * wh:(h + 1) * wh], wh, k_pad) for h in range(nh)]      # (nk, 2, wh/8, KC/8, 8, 8) each
        if nh == 1:
            return [tiles[0].reshape(-1)]
        return [tiles[0][:ks].reshape(-1), tiles[1][:ks].reshape(-1), tiles[0][ks:].reshape(-1), tiles[1][ks:].reshape(-1)]

    parts = body(pack['w_in'], k0p)
    cb = [pack['b_in']]
    for k in range(pack['w_res'].shape[0]):
        alpha = pack['alpha'][k]
        parts += body(alpha * pack['w_res'][k], W)
        cb.append(cb[-1] + alpha * pack['b_res'][k])
    head = torch.cat([pack['w_pol'], pack['w_val'][None]], 0)            # (A+1, W)
    parts.append(_split(head, n_p, W).reshape(-1))
    b_head = pack['b_pol'].new_zeros((n_p,))
    b_head[:A] = pack['b_pol']
    b_head[A] = pack['b_val'][0]
    return torch.cat(parts).contiguous(), torch.cat([torch.stack(cb).reshape(-1), b_head]).contiguous()


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
