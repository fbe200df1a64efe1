"""The learner step on the self-play trajectories — ``main.as_chunk`` / ``learning.reward_to_go`` / ``main.optimize``
(boardlaw/main.py:61-101, boardlaw/learning.py:57-76) — SURVEY.md 8 f2.

Hand-written kernels (through the C ABI): the reward-to-go scan, the fused policy/value loss with its gradient, Adam over
one flat parameter buffer (csrc/learner.cu), and the dense contractions — forward, dgrad, wgrad of ``FCModel`` — on the tcgen05
tensor cores with fp32 accuracy (csrc/gemm_tc.cu, ``gemm`` below: split-fp16 products, fp32 accumulation in tensor memory,
transposed operands as stride swaps, wgrad split over the sample axis).  There is no autograd graph — the backward below is the
network's written out; the remaining torch calls are elementwise / reductions.
Arithmetic: fp32 throughout (the reference runs this step under fp16 autocast with a GradScaler, boardlaw/main.py:88,103-106;
fp32 is the higher-precision form of the same update).
"""
import torch

from . import _lib, arrdict
from ._lib import ptr, check


def amax(t):
    """max |t| as a 0-dim device tensor (one reduction, no host sync): the scale ``gemm`` brings an operand into fp16's range with."""
    return torch.linalg.vector_norm(t, float('inf'))


def gemm(a, b, bias=None, a_relu=False, b_relu=False, out=None, a_amax=None, b_amax=None):
    """``op(a) @ op(b).T (+ bias)`` on the tensor cores with fp32 accuracy (csrc/gemm_tc.cu): a (M, K) and b (N, K) are 2-D fp32 CUDA
    tensors with ANY strides (pass ``x.t()`` for a transposed operand: no copy is made); ``a_relu`` / ``b_relu`` apply relu to the
    operand on the way in.  The kernel scales each operand into fp16's range by a power of two taken from its max |x| — ``a_amax`` /
    ``b_amax`` (0-dim device tensors; any upper bound within a few binades works), computed here on the device when not given (no
    host sync).  Returns (M, N) fp32, or fills ``out`` (row-major rows, any row pitch)."""
    dev = _lib.require_cuda(a, b)
    assert a.dtype == torch.float32 and b.dtype == torch.float32 and a.ndim == 2 and b.ndim == 2 and a.shape[1] == b.shape[1]
    M, K = a.shape
    N = b.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=dev)
    assert out.shape == (M, N) and out.dtype == torch.float32 and (N == 1 or out.stride(1) == 1)
    if M == 0 or N == 0:
        return out
    if a_amax is None:
        a_amax = amax(a) if K else a.new_zeros(())
    if b_amax is None:
        b_amax = amax(b) if K else b.new_zeros(())
    lib = _lib.lib()
    wsb = lib.bl_gemm_f32_workspace_bytes(M, N, K)
    ws = torch.empty((wsb,), dtype=torch.uint8, device=dev) if wsb else None
    if bias is not None:
        bias = bias.float().contiguous()
    check(lib.bl_gemm_f32(ptr(a), a.stride(0), a.stride(1), int(a_relu), ptr(a_amax), ptr(b), b.stride(0), b.stride(1), int(b_relu),
                          ptr(b_amax), ptr(bias) if bias is not None else None, ptr(out), out.stride(0), M, N, K,
                          ptr(ws) if ws is not None else None, wsb, _lib.stream_for(dev)), 'bl_gemm_f32')
    return out


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
            obs = obs.float()
            one = obs.new_ones(())                                                      # max |obs| (0/1 planes)
            wv, wp = net.value.core.weight, net.policy.core.weight                      # (1, W), (A, W)
            wmax = {id(m): amax(m.weight) for m in body}
            wp_max, wv_max = amax(wp), amax(wv)
            x = gemm(obs, body[0].weight, bias=body[0].bias, a_amax=one, b_amax=wmax[id(body[0])])   # heads.py:47-52
            xs, xmax, us = [x], [amax(x)], []
            for blk in body[1:]:                                                       # networks.py:17-18
                u = gemm(x, blk.weight, bias=blk.bias, a_relu=True, a_amax=xmax[-1], b_amax=wmax[id(blk)])
                x = x + getattr(blk, 'α') * u
                us.append(u)
                xs.append(x)
                xmax.append(amax(x))
            scores = gemm(x, wp, bias=net.policy.core.bias, a_amax=xmax[-1], b_amax=wp_max)          # heads.py:101-104
            logp = torch.log_softmax(scores.masked_fill(~valid, float('-inf')), -1)
            t = torch.tanh(gemm(x, wv, bias=net.value.core.bias, a_amax=xmax[-1], b_amax=wv_max)[:, 0])   # heads.py:136-142
            seats = worlds.seats.int().contiguous()
            v = torch.where(seats[:, None] == 0, torch.stack([t, -t], -1), torch.stack([-t, t], -1)).contiguous()

            dscores = torch.empty_like(logp)
            dz = torch.empty((N,), dtype=torch.float32, device=self.device)
            sums = torch.zeros((2,), dtype=torch.float32, device=self.device)
            tl = _lib.proxy(d0.logits.contiguous(), torch.float16, 2, 'decisions.logits')
            tv = _lib.proxy(batch.reward_to_go.contiguous(), torch.float16, 2, 'reward_to_go')
            check(_lib.lib().bl_policy_value_loss(ptr(logp), ptr(v), ptr(tl), ptr(tv), ptr(seats), ptr(dscores), ptr(dz), ptr(sums),
                                                  N, A, _lib.stream_for(self.device)), 'bl_policy_value_loss')
            # backward of the network, written out: wgrad = dY^T . X (both operands transposed views, split over the sample axis),
            # dgrad = dY . W (W as a transposed view); the ReZero gate multiplies the small results, not the (N, W) operands
            ds_max, dz_max = amax(dscores), amax(dz)
            gemm(dscores.t(), x.t(), out=g['policy.core.weight'], a_amax=ds_max, b_amax=xmax[-1])
            g['policy.core.bias'].copy_(dscores.sum(0))
            gemm(dz[None, :], x.t(), out=g['value.core.weight'], a_amax=dz_max, b_amax=xmax[-1])
            g['value.core.bias'].copy_(dz.sum(0, keepdim=True))
            dx = gemm(dscores, wp.t(), a_amax=ds_max, b_amax=wp_max).addcmul_(dz[:, None], wv)
            for k in range(len(body) - 1, 0, -1):
                blk, u, xin = body[k], us[k - 1], xs[k - 1]
                alpha = getattr(blk, 'α')
                dmax = amax(dx)
                g[f'body.{k}.α'].copy_((dx * u).sum())
                gw = g[f'body.{k}.weight']
                gemm(dx.t(), xin.t(), b_relu=True, out=gw, a_amax=dmax, b_amax=xmax[k - 1])
                gw.mul_(alpha)
                g[f'body.{k}.bias'].copy_(alpha * dx.sum(0))
                dx = dx + (alpha * gemm(dx, blk.weight.t(), a_amax=dmax, b_amax=wmax[id(blk)])) * (xin > 0)
            gemm(dx.t(), obs.t(), out=g['body.0.weight'], a_amax=amax(dx), b_amax=one)
            g['body.0.bias'].copy_(dx.sum(0))
        return -sums[0] / N, sums[1] / (2 * N)

    def apply(self):
        """One Adam step on the flat buffer with the gradients currently in ``self.grad``."""
        self.step += 1
        check(_lib.lib().bl_adam_step(ptr(self.flat), ptr(self.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq), self.flat.numel(),
                                      float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps), self.step,
                                      _lib.stream_for(self.device)), 'bl_adam_step')
        self.network._pack_key = None                 # the kernel wrote through raw pointers: restage the inference operands

    def optimize(self, batch, group=None, world=1):
        """``main.optimize``: returns arrdict(policy_loss, value_loss).  With ``world`` > 1 the batch is this rank's SHARD of the
        step's samples (equal shards): the gradient (and the reported losses) are averaged over the ranks with one NCCL all-reduce
        of the flat gradient buffer, so every replica applies the update of the whole batch — the data-parallel form of the
        reference's single learner."""
        pl, vl = self.forward_backward(batch)
        if world > 1:
            import torch.distributed as dist
            losses = torch.stack([pl, vl])
            dist.all_reduce(self.grad, group=group)
            dist.all_reduce(losses, group=group)
            self.grad.div_(world)
            pl, vl = losses[0] / world, losses[1] / world
        self.apply()
        return arrdict.arrdict(policy_loss=pl, value_loss=vl)


def chunk_from_records(records, boardsize, batch_size=None, shard=None):
    """The learner's chunk from the all-gathered trajectory records (``selfplay.TrajectoryPool.wait()``: one (world, B, R)
    uint8 tensor per buffered move) — the same structure ``as_chunk`` builds from the actor's buffer (boardlaw/main.py:61-73,
    171-181), with the ranks' shards side by side on the env axis.  Returns (chunk, remaining records)."""
    from . import selfplay
    from .hex import Hex
    if shard is not None:
        records_view = [r[shard:shard + 1] for r in records]                       # one rank's envs only (the data-parallel learner)
    else:
        records_view = records
    rec = torch.stack([r.reshape(-1, r.shape[-1]) for r in records_view])        # (T, world*B, R)
    u = selfplay.unpack_records(rec, boardsize)
    chunk = arrdict.arrdict(
        worlds=Hex(board=u.board, seats=u.seats),
        decisions=arrdict.arrdict(logits=u.logits, prior=u.prior, v=u.v, actions=u.actions),
        transitions=arrdict.arrdict(rewards=u.rewards, terminal=u.terminal))
    chunk['reward_to_go'] = reward_to_go(u.rewards.float(), u.v.float(), u.terminal, half=True)
    n_new = (batch_size or rec.shape[1]) // rec.shape[1]
    return chunk, records[n_new:]
