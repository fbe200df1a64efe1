"""The learner step on the self-play trajectories — ``main.as_chunk`` / ``learning.reward_to_go`` / ``main.optimize``
(boardlaw/main.py:61-101, boardlaw/learning.py:57-76) — SURVEY.md 8 f2.

Hand-written kernels (csrc/learner.cu, through the C ABI): the reward-to-go scan, the fused policy/value loss with its
gradient, and Adam over one flat parameter buffer.  The learner's dense contractions are plain GEMMs (forward, dgrad,
wgrad of ``FCModel``: ~3x one network launch of the self-play path per 64 moves) and go to cuBLAS through
``torch.addmm`` / ``torch.mm``; there is no autograd graph — the backward below is the network's written out.
Arithmetic: fp32 throughout (the reference runs this step under fp16 autocast with a GradScaler, boardlaw/main.py:88,103-106;
fp32 is the higher-precision form of the same update).
"""
import torch

from . import _lib, arrdict
from ._lib import ptr, check


def reward_to_go(reward, value, terminal, gamma=1., half=False):
    """``learning.reward_to_go`` (boardlaw/learning.py:70-76): reward, value (T, ...) f32; terminal bool of the same shape, or
    without the trailing seat axis.  Returns (T, ...) f32 (half when ``half``, as as_chunk stores it).  ``value`` is left
    untouched (the reference overwrites its terminal entries with the reward, on a temporary)."""
    dev = _lib.require_cuda(reward, value, terminal)
    shape = reward.shape
    T = shape[0]
    if terminal.shape == shape and reward.ndim >= 3:
        terminal = terminal[..., 0]                               # as_chunk stacks one copy per seat (main.py:63)
    Sn = shape[-1] if (reward.ndim >= 3 and terminal.ndim == reward.ndim - 1) else 1
    B = reward[0].numel() // Sn
    assert value.shape == shape and terminal.numel() == T * B
    reward = reward.float().contiguous()
    value = value.float().contiguous()
    term = terminal.contiguous().view(torch.uint8) if terminal.dtype == torch.bool else terminal.to(torch.uint8).contiguous()
    out = torch.empty(shape, dtype=torch.float16 if half else torch.float32, device=dev)
    check(_lib.lib().bl_reward_to_go(ptr(reward), ptr(value), ptr(term), ptr(out), int(half), T, B, Sn, float(gamma),
                                     _lib.stream_for(dev)), 'bl_reward_to_go')
    return out


def as_chunk(buffer, batch_size):
    """``main.as_chunk`` (boardlaw/main.py:61-73) without its statistics side channel: stacks the buffered moves into a
    (T, B, ...) chunk, attaches ``reward_to_go`` (half) and drops the oldest ``batch_size // B`` moves from the buffer."""
    chunk = arrdict.stack(buffer)
    chunk['reward_to_go'] = reward_to_go(chunk.transitions.rewards.float(), chunk.decisions.v.float(),
                                         chunk.transitions.terminal, half=True)
    n_new = batch_size // chunk.transitions.terminal.size(1)
    return chunk, buffer[n_new:]


class Learner:
    """``opt = Adam(network.parameters(), lr)`` + ``optimize(network, scaler, opt, batch)`` (boardlaw/main.py:75-101,154).
    The network's parameters are re-pointed into one flat fp32 buffer so that the update is a single kernel."""

    def __init__(self, network, lr=1e-3, betas=(.9, .999), eps=1e-8):
        self.network = network
        self.lr, self.betas, self.eps = lr, betas, eps
        named = list(network.named_parameters())
        dev = named[0][1].device
        if dev.type != 'cuda':
            raise RuntimeError('the learner runs on a CUDA device only (there is no CPU path)')
        n = sum(p.numel() for _, p in named)
        self.flat = torch.empty((n,), dtype=torch.float32, device=dev)
        self.grad = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.g, off = {}, 0
        for name, p in named:
            k = p.numel()
            self.flat[off:off + k].copy_(p.detach().float().reshape(-1))
            p.data = self.flat[off:off + k].view(p.shape)
            self.g[name] = self.grad[off:off + k].view(p.shape)
            off += k
        self.step = 0
        self.device = dev

    # ---- forward + loss + backward: fills self.grad, returns (policy_loss, value_loss) as 0-dim device tensors --------------
    def forward_backward(self, batch):
        net, g = self.network, self.g
        worlds, d0 = batch.worlds, batch.decisions
        N = worlds.board.shape[0]
        A = worlds.board.shape[-1] ** 2
        obs = worlds.obs.reshape(N, -1)
        valid = worlds.valid
        body = list(net.body)
        with torch.no_grad():
            x = torch.addmm(body[0].bias, obs, body[0].weight.t())                     # heads.py:47-52
            xs, us = [x], []
            for blk in body[1:]:                                                       # networks.py:17-18
                u = torch.addmm(blk.bias, torch.relu(x), blk.weight.t())
                x = x + getattr(blk, 'α') * u
                us.append(u)
                xs.append(x)
            scores = torch.addmm(net.policy.core.bias, x, net.policy.core.weight.t())  # heads.py:101-104
            logp = torch.log_softmax(scores.masked_fill(~valid, float('-inf')), -1)
            t = torch.tanh(torch.addmv(net.value.core.bias, x, net.value.core.weight[0]))   # heads.py:136-142
            seats = worlds.seats.int().contiguous()
            v = torch.where(seats[:, None] == 0, torch.stack([t, -t], -1), torch.stack([-t, t], -1)).contiguous()

            dscores = torch.empty_like(logp)
            dz = torch.empty((N,), dtype=torch.float32, device=self.device)
            sums = torch.zeros((2,), dtype=torch.float32, device=self.device)
            tl = _lib.proxy(d0.logits.contiguous(), torch.float16, 2, 'decisions.logits')
            tv = _lib.proxy(batch.reward_to_go.contiguous(), torch.float16, 2, 'reward_to_go')
            check(_lib.lib().bl_policy_value_loss(ptr(logp), ptr(v), ptr(tl), ptr(tv), ptr(seats), ptr(dscores), ptr(dz), ptr(sums),
                                                  N, A, _lib.stream_for(self.device)), 'bl_policy_value_loss')
            # backward of the network, written out
            torch.mm(dscores.t(), x, out=g['policy.core.weight'])
            g['policy.core.bias'].copy_(dscores.sum(0))
            g['value.core.weight'].copy_((dz[None, :] @ x))
            g['value.core.bias'].copy_(dz.sum(0, keepdim=True))
            dx = torch.addmm(dz[:, None] * net.value.core.weight, dscores, net.policy.core.weight)
            for k in range(len(body) - 1, 0, -1):
                blk, u, xin = body[k], us[k - 1], xs[k - 1]
                alpha = getattr(blk, 'α')
                g[f'body.{k}.α'].copy_((dx * u).sum())
                du = alpha * dx
                torch.mm(du.t(), torch.relu(xin), out=g[f'body.{k}.weight'])
                g[f'body.{k}.bias'].copy_(du.sum(0))
                dx = dx + (du @ blk.weight) * (xin > 0)
            torch.mm(dx.t(), obs, out=g['body.0.weight'])
            g['body.0.bias'].copy_(dx.sum(0))
        return -sums[0] / N, sums[1] / (2 * N)

    def apply(self):
        """One Adam step on the flat buffer with the gradients currently in ``self.grad``."""
        self.step += 1
        check(_lib.lib().bl_adam_step(ptr(self.flat), ptr(self.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq), self.flat.numel(),
                                      float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps), self.step,
                                      _lib.stream_for(self.device)), 'bl_adam_step')
        self.network._pack_key = None                 # the kernel wrote through raw pointers: restage the inference operands

    def optimize(self, batch):
        """``main.optimize``: returns arrdict(policy_loss, value_loss)."""
        pl, vl = self.forward_backward(batch)
        self.apply()
        return arrdict.arrdict(policy_loss=pl, value_loss=vl)


def chunk_from_records(records, boardsize, batch_size=None):
    """The learner's chunk from the all-gathered trajectory records (``selfplay.TrajectoryPool.wait()``: one (world, B, R)
    uint8 tensor per buffered move) — the same structure ``as_chunk`` builds from the actor's buffer (boardlaw/main.py:61-73,
    171-181), with the ranks' shards side by side on the env axis.  Returns (chunk, remaining records)."""
    from . import selfplay
    from .hex import Hex
    rec = torch.stack([r.reshape(-1, r.shape[-1]) for r in records])             # (T, world*B, R)
    u = selfplay.unpack_records(rec, boardsize)
    chunk = arrdict.arrdict(
        worlds=Hex(board=u.board, seats=u.seats),
        decisions=arrdict.arrdict(logits=u.logits, prior=u.prior, v=u.v, actions=u.actions),
        transitions=arrdict.arrdict(rewards=u.rewards, terminal=u.terminal))
    chunk['reward_to_go'] = reward_to_go(u.rewards.float(), u.v.float(), u.terminal, half=True)
    n_new = (batch_size or rec.shape[1]) // rec.shape[1]
    return chunk, records[n_new:]
