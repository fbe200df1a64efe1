"""Env-sharded self-play across the GPUs of one box.

The reference has no multi-GPU code: a run is one process on one GPU.  Here each rank owns an
independent shard of envs — exactly one reference process per GPU, no communication inside a move — and the only
collective is one all-gather per move that pools fixed-size per-env trajectory records for the learner
(actor loop of boardlaw/main.py:171-181: ``buffer.append(arrdict(worlds, decisions.half(), transitions))``).
The gather runs on a side stream so it overlaps the next move's search.
"""
import torch
import torch.distributed as dist

from . import arrdict


def record_width(n_actions):
    """board A u8 | seat u8 | terminal u8 | action i16 | rewards 2 x f16 | v 2 x f16 | logits A x f16 | prior A x f16,
    padded to a multiple of 16 bytes."""
    return (5 * n_actions + 12 + 15) // 16 * 16


def pack_records(worlds, decisions, transitions):
    """One (B, record_width) uint8 row per env for the move just played."""
    B = worlds.board.shape[0]
    A = worlds.board.shape[-1] ** 2
    b = lambda t: t.contiguous().view(torch.uint8).reshape(B, -1)
    parts = [
        worlds.board.reshape(B, A),
        worlds.seats.to(torch.uint8)[:, None],
        transitions.terminal.to(torch.uint8)[:, None],
        b(decisions.actions.to(torch.int16)),
        b(transitions.rewards.half()),
        b(decisions.v.half()),
        b(decisions.logits.half()),
        b(decisions.prior.half())]
    rec = torch.cat(parts, -1)
    pad = record_width(A) - rec.shape[1]
    if pad:
        rec = torch.cat([rec, rec.new_zeros((B, pad))], -1)
    return rec


def unpack_records(rec, boardsize):
    """Inverse of pack_records (any leading shape)."""
    A = boardsize * boardsize
    lead = rec.shape[:-1]
    o = 0

    def take(nbytes, dtype, shape):
        nonlocal o
        out = rec[..., o:o + nbytes].contiguous().view(dtype).reshape(*lead, *shape)
        o += nbytes
        return out

    return arrdict.arrdict(
        board=take(A, torch.uint8, (boardsize, boardsize)),
        seats=take(1, torch.uint8, ()).int(),
        terminal=take(1, torch.uint8, ()).bool(),
        actions=take(2, torch.int16, ()).long(),
        rewards=take(4, torch.float16, (2,)),
        v=take(4, torch.float16, (2,)),
        logits=take(2 * A, torch.float16, (A,)),
        prior=take(2 * A, torch.float16, (A,)))


class TrajectoryPool:
    """All-gathers each move's records over the process group (NCCL over NVLink on GPUs, gloo in CPU tests)."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._side = None
        self._pending = None

    def gather(self, rec):
        """Starts the gather of this rank's (B, R) records; returns immediately on CUDA (side stream)."""
        if self.world == 1:
            self._pending = (rec[None], None)
            return
        out = rec.new_empty((self.world,) + tuple(rec.shape))
        if rec.is_cuda:
            if self._side is None:
                self._side = torch.cuda.Stream(device=rec.device)
            self._side.wait_stream(torch.cuda.current_stream(rec.device))
            with torch.cuda.stream(self._side):
                dist.all_gather_into_tensor(out, rec, group=self.group)
                rec.record_stream(self._side)
                out.record_stream(self._side)     # allocated on the caller's stream, written on this one
            self._pending = (out, self._side)
        else:
            try:
                dist.all_gather_into_tensor(out, rec.contiguous(), group=self.group)
            except Exception:
                parts = [torch.empty_like(rec) for _ in range(self.world)]
                dist.all_gather(parts, rec.contiguous(), group=self.group)
                out = torch.stack(parts)
            self._pending = (out, None)

    def sync_before_overwrite(self):
        """Orders the current stream after the gather in flight: call before re-using the buffer that was handed to ``gather`` (the
        engine's static record buffer is rewritten by the next move)."""
        if self._pending is not None and self._pending[1] is not None:
            torch.cuda.current_stream(self._pending[0].device).wait_stream(self._pending[1])

    def wait(self):
        """(world, B, R) records of the last gathered move, safe to read on the current stream."""
        out, side = self._pending
        if side is not None:
            torch.cuda.current_stream(out.device).wait_stream(side)
        return out


class SelfPlay:
    """worlds -> agent -> step, one shard per rank; ``step()`` is the unit the benchmark times."""

    def __init__(self, worlds, agent, pool=None):
        self.worlds = worlds
        self.agent = agent
        self.pool = pool

    def step(self):
        """One move of every env.  With an agent that has ``play`` (MCTSAgent on the fused engine) the whole move — search, action,
        env transition, trajectory record — is one captured CUDA graph; what comes back are views of the engine's static buffers,
        valid until the next step."""
        if hasattr(self.agent, 'play'):
            if self.pool is not None:
                self.pool.sync_before_overwrite()          # the previous move's record is still being gathered on the side stream
                decisions, new_worlds, transitions, rec = self.agent.play(self.worlds, record=True)
                self.pool.gather(rec if rec is not None else pack_records(self.worlds, decisions, transitions))
            else:
                decisions, new_worlds, transitions = self.agent.play(self.worlds)
        else:
            decisions = self.agent(self.worlds, value=True)
            new_worlds, transitions = self.worlds.step(decisions.actions)
            if self.pool is not None:
                self.pool.gather(pack_records(self.worlds, decisions, transitions))
        self.worlds = new_worlds
        return decisions, transitions


def weight_checksum(network):
    """A float64 fingerprint of every parameter (sum and sum of squares), cheap enough to compare after every optimiser step."""
    ps = [p.detach().double().reshape(-1) for p in network.parameters()]
    flat = torch.cat(ps) if ps else torch.zeros(0, dtype=torch.double)
    return torch.stack([flat.sum(), (flat * flat).sum()])


def check_replicas(network, pool):
    """Raises if the ranks' replicas of ``network`` have drifted apart: the env-sharded loop relies on every rank applying the same
    deterministic update to identical weights (no weight broadcast).  One 16-byte all-gather."""
    if pool is None or pool.world == 1:
        return
    mine = weight_checksum(network)
    out = [torch.empty_like(mine) for _ in range(pool.world)]
    dist.all_gather(out, mine, group=pool.group)
    for r, other in enumerate(out):
        if not torch.equal(other, out[0]):
            raise RuntimeError(f'rank {r} holds different network weights than rank 0 (checksums {other.tolist()} vs {out[0].tolist()})')


def shard_bounds(n_envs, world, rank):
    """rank r owns envs [r*n/world, (r+1)*n/world)."""
    return rank * n_envs // world, (rank + 1) * n_envs // world
