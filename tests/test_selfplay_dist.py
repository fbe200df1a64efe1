"""CPU, world_size 2 over gloo: the multi-rank host logic of boardlaw_b200/selfplay.py — shard bounds, the packed
trajectory record (pack -> all-gather -> unpack round trip) — without touching a GPU kernel."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from boardlaw_b200 import arrdict
from boardlaw_b200.selfplay import TrajectoryPool, check_replicas, pack_records, record_width, shard_bounds, unpack_records, weight_checksum


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _fake_move(B, S, seed):
    """A move's worth of records with every field pseudo-random (incl. -inf logits), as plain CPU tensors."""
    g = torch.Generator().manual_seed(seed)
    A = S * S
    logits = torch.randn((B, A), generator=g).half()
    logits[torch.rand((B, A), generator=g) < .3] = -float('inf')
    worlds = arrdict.arrdict(board=torch.randint(0, 7, (B, S, S), generator=g, dtype=torch.uint8),
                             seats=torch.randint(0, 2, (B,), generator=g, dtype=torch.int32))
    decisions = arrdict.arrdict(actions=torch.randint(0, A, (B,), generator=g), logits=logits,
                                prior=torch.randn((B, A), generator=g).half(), v=torch.rand((B, 2), generator=g).half())
    transitions = arrdict.arrdict(terminal=torch.rand((B,), generator=g) < .2,
                                  rewards=torch.randint(-1, 2, (B, 2), generator=g).float())
    return worlds, decisions, transitions


def _check_roundtrip(rec, S, worlds, decisions, transitions):
    u = unpack_records(rec, S)
    assert torch.equal(u.board, worlds.board) and torch.equal(u.seats, worlds.seats)
    assert torch.equal(u.actions, decisions.actions) and torch.equal(u.terminal, transitions.terminal)
    assert torch.equal(u.rewards, transitions.rewards.half())
    for a, b in ((u.v, decisions.v), (u.logits, decisions.logits), (u.prior, decisions.prior)):
        assert torch.equal(a.view(torch.int16), b.view(torch.int16))


def test_record_roundtrip_single():
    for S in (3, 5, 9, 11, 13):
        w, d, t = _fake_move(17, S, S)
        rec = pack_records(w, d, t)
        assert rec.shape == (17, record_width(S * S)) and rec.shape[1] % 16 == 0
        _check_roundtrip(rec, S, w, d, t)


def test_shard_bounds_partition():
    for n, world in ((262144, 8), (1000, 3), (7, 8), (0, 2)):
        spans = [shard_bounds(n, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def _worker(rank, world, port, S, B, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        pool = TrajectoryPool()
        assert pool.world == world and pool.rank == rank
        for move in range(3):
            w, d, t = _fake_move(B, S, seed=100 * move + rank)
            pool.gather(pack_records(w, d, t))
            got = pool.wait()
            assert got.shape == (world, B, record_width(S * S))
            for r in range(world):                      # every rank sees every shard's records, in rank order
                _check_roundtrip(got[r], S, *_fake_move(B, S, seed=100 * move + r))
        # replica check of the env-sharded actor/learner loop (main.run): identical weights pass, a drifted replica raises everywhere
        torch.manual_seed(0)
        net = torch.nn.Linear(7, 5)
        check_replicas(net, pool)
        with torch.no_grad():
            net.weight[0, 0] += float(rank)
        try:
            check_replicas(net, pool)
            raise AssertionError('drifted replicas were not detected')
        except RuntimeError as e:
            assert 'different network weights' in str(e)
        out.put((rank, 'ok'))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_trajectory_allgather_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 5, 12, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert results == {0: 'ok', 1: 'ok'}, results


def test_weight_checksum_detects_a_changed_parameter():
    torch.manual_seed(1)
    a, b = torch.nn.Linear(9, 4), torch.nn.Linear(9, 4)
    b.load_state_dict(a.state_dict())
    assert torch.equal(weight_checksum(a), weight_checksum(b))
    with torch.no_grad():
        b.bias[2] += 1e-6
    assert not torch.equal(weight_checksum(a), weight_checksum(b))
