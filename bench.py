#!/usr/bin/env python
"""MCTS simulations/sec for vectorised Hex self-play (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2] [--impl ours|reference]

A "step" is one move of every env: ``decisions = agent(worlds)`` (n_nodes simulations per env) followed by
``worlds.step(decisions.actions)``; value = n_envs * n_nodes * steps / time.  Prints ONE JSON line.
For N > 1 launch under torch.distributed.run: each rank owns its own shard of envs (weak scaling) and the only
collective is the per-move trajectory all-gather over NCCL.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# name -> (boardsize, envs per GPU, n_nodes, width, depth)
CONFIGS = {
    'c1': (5, 256, 16, 32, 2),
    'c2': (9, 32768, 64, 256, 4),
    'c3': (11, 65536, 256, 512, 8),
    'c5-5': (5, 32768, 64, 256, 4), 'c5-7': (7, 32768, 64, 256, 4), 'c5-9': (9, 32768, 64, 256, 4),
    'c5-11': (11, 32768, 64, 256, 4), 'c5-13': (13, 32768, 64, 256, 4),
}
CPU_SAMPLE_ENVS = {'c1': 256, 'c2': 2048, 'c3': 64}      # 2 048 envs at c2 (cost is linear in envs: serial per-env loops), ~8 s per move


def describe(config):
    S, B, T, W, D = CONFIGS[config]
    return f'Hex {S}x{S}, {B} envs/GPU, {T} sims/move, FCModel width {W} depth {D}'


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.FIELDS}',
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')] + [time.time()])

    def window(self, t0, t1):
        """Keeps the samples taken inside the timed region [t0, t1]; a region shorter than nvidia-smi's period keeps the samples
        nearest to it instead (taken under the same load: the sampler runs from the warm-up on)."""
        inside = [r for r in self.rows if t0 <= r[-1] <= t1]
        self.in_window = bool(inside)
        if not inside:
            inside = sorted(self.rows, key=lambda r: min(abs(r[-1] - t0), abs(r[-1] - t1)))[:3]
        self.rows = inside

    def __exit__(self, *a):
        if self.proc:
            time.sleep(.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k] == 'Active' for r in self.rows)]
        smax = max((int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()), default=None)
        out = {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': smax, 'reasons': reasons, 'samples': len(sm)}
        if not getattr(self, 'in_window', True):
            out['note'] = 'timed region shorter than the sampling period: nearest samples under the same load (warm-up / e2e pass)'
        return out


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def make_worlds(S, B, device, seed):
    """Hex.initial advanced by 2*A uniformly random valid moves (auto-resets spread envs over all game phases)."""
    import torch
    from boardlaw_b200.hex import Hex
    g = torch.Generator(device=device).manual_seed(seed)
    worlds = Hex.initial(B, S, device=device)
    for _ in range(2 * S * S):
        actions = torch.multinomial(worlds.valid.float(), 1, generator=g).squeeze(-1)
        worlds, _ = worlds.step(actions)
    return worlds


def algorithmic_bytes(S, T, counters, n_desc):
    """SURVEY.md 8(d) one-touch bytes under the reference's tensor layout, from the engine's exact counters."""
    A, Sn = S * S, 2
    evals, children, _, _, backup_nodes = counters[:5]
    descend = evals * (4 * A + 5) + 4 * children          # children row 2A + logits row 2A + seat/terminal/rand, +4 per child
    expand_step = n_desc * (6 + 2 * A + 8 + 2 * Sn + 1)
    net_io = n_desc * (2 * A + 2 * Sn) + n_desc * 2 * A    # logits + v out, child-row init
    backup = backup_nodes * (7 + 6 * Sn) + n_desc * T * (2 * Sn + 2)
    return {'descend_expand': descend + expand_step, 'net': net_io, 'backup': backup}


def source_sha():
    import hashlib
    h = hashlib.sha1()
    for f in sorted((ROOT / 'boardlaw_b200' / 'csrc').glob('*.cu*')):
        h.update(f.name.encode()); h.update(f.read_bytes())
    return h.hexdigest()[:16]


def ncu_traffic(kind, config):
    """dram read + write bytes per launch of a kernel class, mean over every launch of one move, from the capture taken by
    tools/capture_traffic.py (profiles/r02_traffic_<config>.json).  None when there is no capture for this config or the kernel
    sources have changed since it was taken (the file records their sha1)."""
    f = ROOT / 'profiles' / f'r02_traffic_{config.split("-")[0] if config.startswith("c5-9") else config}.json'
    if not f.exists():
        return None, 'no capture for this config'
    d = json.loads(f.read_text())
    if d.get('source_sha') != source_sha():
        return None, f'capture {f.name} is of other kernel sources (sha {d.get("source_sha")})'
    c = d['classes'].get(kind)
    if not c:
        return None, 'kernel class not in the capture'
    return c['dram_bytes_per_launch'], f'{f.name}: {c["kernel"]}, {c["launches"]} launches, {c["registers"]} registers'


def net_flops(S, W, D):
    A = S * S
    return 2 * (2 * A * W + D * W * W + W * A + W)


def kernel_split(agent, worlds, T, n_moves=2):
    """Eager (un-graphed) moves with CUDA events around every C-ABI launch group: ms per move by kernel class."""
    import torch
    from boardlaw_b200.mcts import engine_for, dirichlet_mix
    eng = engine_for(worlds, T)
    net = agent.network
    cp = net.packed()
    acc = {'descend_expand': 0., 'net': 0., 'backup': 0., 'other': 0.}
    launches = {'descend_expand': 0, 'net': 0, 'backup': 0}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    eng.ws.counters[:6].zero_()
    for _ in range(n_moves):
        pairs = []

        def timed(kind, fn):
            a, b = ev(), ev()
            a.record(); fn(); b.record()
            pairs.append((kind, a, b))
        timed('other', lambda: (eng.reset(worlds.board, worlds.seats, 1 / 16), eng.eval_root(cp)))
        timed('other', lambda: eng.set_root_prior(.25, 10))
        for sim in range(1, T):
            timed('descend_expand', lambda: eng.descend_expand(sim))
            timed('net', lambda: eng.eval_leaves(cp, sim))
            timed('backup', lambda: eng.backup(sim))
            for k in launches:
                launches[k] += 1
        timed('other', lambda: (eng.root_act(T), eng.step_and_record(True)))
        torch.cuda.synchronize()
        for kind, a, b in pairs:
            acc[kind] += a.elapsed_time(b)
    counters = eng.ws.counters.cpu().tolist()
    return {k: v / n_moves for k, v in acc.items()}, [c / n_moves for c in counters], {k: v // n_moves for k, v in launches.items()}


def time_hex_step(worlds, hbm_peak, hbm_src, reps=20):
    """The env transition of the real worlds (bl_hex_transition: copy + move + auto-reset, one launch per move) on its own:
    algorithmic bytes 2A + 8 + 2*Sn*4 + 1 per env (board in and out, seats, rewards, terminal)."""
    import torch
    A = worlds.boardsize ** 2
    actions = torch.multinomial(worlds.valid.float(), 1).squeeze(-1)
    worlds.step(actions)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        worlds.step(actions)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    nbytes = worlds.n_envs * (2 * A + 8 + 16 + 1 + 8)
    gbs = nbytes / (us * 1e-6) / 1e9
    return {'bound': 'hbm', 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gbs / hbm_peak, 'peak_source': hbm_src,
            'us_per_launch': us, 'algorithmic_bytes_per_launch': nbytes, 'traffic': None,
            'note': 'includes the torch allocations of Hex.step\'s outputs; at 5.7 MB per launch the kernel is launch/latency-bound, not HBM-bound'}


def run_leg(leg, config, timeout=900):
    """Runs one baseline leg (``--leg``) of this script in a fresh process and returns the JSON object it prints."""
    try:
        r = subprocess.run([sys.executable, str(ROOT / 'bench.py'), '--leg', leg, '--config', config], capture_output=True, text=True, timeout=timeout)
        for line in reversed(r.stdout.splitlines()):
            if line.startswith('{'):
                return json.loads(line)
        return {'unavailable': (r.stderr or r.stdout)[-300:]}
    except Exception as e:                      # informational legs never take the bench line down
        return {'unavailable': f'{type(e).__name__}: {e}'[:300]}


def measure_config(config, precision='fp32', steps=2, warmup=3):
    """value (sims/s) and ms per move of one more config, through SelfPlay.step as the headline (device-resident worlds)."""
    import torch
    from boardlaw_b200 import heads, mcts
    from boardlaw_b200.mcts import MCTSAgent
    from boardlaw_b200.networks import FCModel, synthetic_state_dict
    from boardlaw_b200.selfplay import SelfPlay
    S, B, T, W, D = CONFIGS[config]
    device = torch.device('cuda', torch.cuda.current_device())
    net = FCModel(heads.Tensor((S, S, 2)), heads.Masked(S * S), width=W, depth=D, precision=precision)
    net.load_state_dict(synthetic_state_dict(S, W, D, seed=0))
    agent = MCTSAgent(net.to(device), n_nodes=T, c_puct=1 / 16)
    torch.manual_seed(0)
    play = SelfPlay(make_worlds(S, B, device, seed=0), agent, None)
    for _ in range(warmup):
        play.step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        play.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del play, agent, net
    for eng in list(mcts._engines.values()):
        eng.release()
    mcts._engines.clear()
    torch.cuda.empty_cache()
    return {'value': B * T / (ms / 1e3), 'unit': 'sims/s', 'ms_per_step': ms, 'steps': steps, 'workload': describe(config), 'net_precision': precision}


def measure_learner(config='c2', steps=5):
    """One optimiser step (Learner.optimize: forward, loss, backward on the tcgen05 GEMMs, Adam) on a chunk of the config's batch size —
    the learner's share of the actor/learner loop (boardlaw/main.py:147-205: one step per move once the buffer is full)."""
    import torch
    from boardlaw_b200 import heads, arrdict
    from boardlaw_b200.learner import Learner
    from boardlaw_b200.networks import FCModel, synthetic_state_dict
    S, B, T, W, D = CONFIGS[config]
    A = S * S
    device = torch.device('cuda', torch.cuda.current_device())
    net = FCModel(heads.Tensor((S, S, 2)), heads.Masked(A), width=W, depth=D)
    net.load_state_dict(synthetic_state_dict(S, W, D, seed=0))
    net = net.to(device)
    worlds = make_worlds(S, B, device, seed=1)
    g = torch.Generator(device=device).manual_seed(0)
    logits = torch.log_softmax(torch.randn((B, A), device=device, generator=g).masked_fill(~worlds.valid, float('-inf')), -1).half()
    batch = arrdict.arrdict(worlds=worlds, decisions=arrdict.arrdict(logits=logits),
                            reward_to_go=(torch.rand((B, 2), device=device, generator=g) * 2 - 1).half())
    L = Learner(net, lr=1e-4)
    for _ in range(3):
        L.optimize(batch)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        L.optimize(batch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    flops = 3 * 2 * B * (2 * A * W + D * W * W + W * (A + 1))
    del L, net, batch, worlds
    torch.cuda.empty_cache()
    return {'ms_per_step': ms, 'samples': B, 'useful_tflops': flops / ms / 1e9, 'workload': f'Learner.optimize on {B} samples, ' + describe(config),
            'note': 'forward, dgrad and wgrad contractions on bl_gemm_f32 (tcgen05, split-fp16, fp32-accurate); not part of `value`'}


def run_extras(args):
    """The other BASELINE.json configs on this GPU, a few moves each (device-timed as `value`): c3, the c5 board-size sweep, and c2 with the
    reference's own GPU precision class (fp16 autocast, boardlaw/mcts/__init__.py:131-133)."""
    out = {}
    for name, precision in [('c2', 'amp'), ('c5-5', 'fp32'), ('c5-7', 'fp32'), ('c5-11', 'fp32'), ('c5-13', 'fp32'), ('c3', 'fp32')]:
        try:
            out[name + ('-amp' if precision == 'amp' else '')] = measure_config(name, precision)
        except Exception as e:
            out[name] = {'unavailable': f'{type(e).__name__}: {e}'[:200]}
    try:
        out['learner-c2'] = measure_learner('c2')
    except Exception as e:
        out['learner-c2'] = {'unavailable': f'{type(e).__name__}: {e}'[:200]}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from boardlaw_b200 import heads
    from boardlaw_b200.hex import Hex
    from boardlaw_b200.mcts import MCTSAgent, engine_for
    from boardlaw_b200.networks import FCModel, synthetic_state_dict
    from boardlaw_b200.selfplay import SelfPlay, TrajectoryPool, record_width

    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world}: launch N>1 under torch.distributed.run'
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)

    S, B, T, W, D = CONFIGS[args.config]
    if args.envs:
        B = args.envs
    A = S * S
    sd = synthetic_state_dict(S, W, D, seed=0)
    net = FCModel(heads.Tensor((S, S, 2)), heads.Masked(A), width=W, depth=D, precision=args.precision)
    net.load_state_dict(sd)
    net = net.to(device)
    agent = MCTSAgent(net, n_nodes=T, c_puct=1 / 16)
    torch.manual_seed(rank)
    worlds = make_worlds(S, B, device, seed=rank)
    pool = TrajectoryPool() if world > 1 else None
    play = SelfPlay(worlds, agent, pool)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    clocks.__enter__()                                         # nvidia-smi takes a while to start: it runs from the warm-up on
    for _ in range(max(args.warmup, 3)):
        play.step()
    eng = engine_for(play.worlds, T)

    # ---- device-resident timing -------------------------------------------------------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = eng.launches
    t_begin = time.time()
    e0.record()
    for _ in range(args.steps):
        play.step()
    if pool is not None:
        pool.wait()
    e1.record()
    barrier()
    t_end = time.time()
    clocks.__exit__()
    clocks.window(t_begin, t_end)
    ms = e0.elapsed_time(e1)
    gpu_launches = eng.launches - launches0                   # every kernel of the move, the env transition and the record packing included, is the engine's
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    value = world * B * T * args.steps / (ms / 1e3)

    # ---- end to end through the public API with host buffers ---------------------------------------------------
    hb = play.worlds.board.cpu().pin_memory()
    hs = play.worlds.seats.cpu().pin_memory()
    out = {'actions': torch.empty((B,), dtype=torch.int64).pin_memory(),
           'logits': torch.empty((B, A), dtype=torch.float16).pin_memory(),
           'v': torch.empty((B, 2), dtype=torch.float16).pin_memory(),
           'board': torch.empty((B, S, S), dtype=torch.uint8).pin_memory(),
           'seats': torch.empty((B,), dtype=torch.int32).pin_memory(),
           'rewards': torch.empty((B, 2), dtype=torch.float32).pin_memory(),
           'terminal': torch.empty((B,), dtype=torch.bool).pin_memory()}
    h2d = hb.numel() * hb.element_size() + hs.numel() * hs.element_size()
    d2h = sum(v.numel() * v.element_size() for v in out.values())

    io = {'hb': hb, 'hs': hs}

    def e2e_step():
        w = Hex(board=io['hb'].to(device, non_blocking=True), seats=io['hs'].to(device, non_blocking=True))
        d, w2, tr = agent.play(w)                             # the whole move (search, action, env transition): one captured graph
        for k, src in (('actions', d.actions), ('logits', d.logits), ('v', d.v), ('board', w2.board), ('seats', w2.seats),
                       ('rewards', tr.rewards), ('terminal', tr.terminal)):
            out[k].copy_(src, non_blocking=True)
        torch.cuda.current_stream().synchronize()            # the caller reads the result on the host
        # next step's inputs come from the host again: the pinned result buffers become the next inputs (no host-side memcpy)
        io['hb'], out['board'] = out['board'], io['hb']
        io['hs'], out['seats'] = out['seats'], io['hs']

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = t.item()
    e2e = world * B * T * args.steps / (ms_e2e / 1e3)

    # ---- per-kernel split + roofline (rank 0) -------------------------------------------------------------------
    result = None
    if rank == 0:
        split, counters, per_move_launches = kernel_split(agent, play.worlds, T)
        peaks = {}
        pk = ROOT / 'MEASURED_PEAKS.json'
        if pk.exists():
            peaks = json.loads(pk.read_text())
        hbm_peak, hbm_src = (peaks['hbm_gbs'], 'measured') if 'hbm_gbs' in peaks else (6650., 'fallback')
        tc_peak = peaks.get('bf16_tflops_sustained', 1400.)
        n_desc = B * (T - 1)
        bytes_by = algorithmic_bytes(S, T, counters, n_desc)
        dominant = max(('descend_expand', 'net', 'backup'), key=lambda k: split[k])
        flops = net_flops(S, W, D) * n_desc
        by_kernel = {}
        for kind in ('descend_expand', 'net', 'backup'):
            traffic, tnote = ncu_traffic(kind, args.config)
            ms_k, n_k = split[kind], per_move_launches[kind]
            hbm = bytes_by[kind] / (ms_k / 1e3) / 1e9
            entry = {'bound': 'hbm', 'achieved': hbm, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': hbm / hbm_peak, 'peak_source': hbm_src,
                     'ms_per_move': ms_k, 'us_per_launch': ms_k * 1e3 / max(n_k, 1), 'algorithmic_bytes_per_launch': bytes_by[kind] / max(n_k, 1),
                     'traffic': traffic, 'traffic_source': tnote}
            if kind == 'net':
                tf = flops / (ms_k / 1e3) / 1e12
                entry.update({'bound': 'tensor', 'achieved': tf, 'peak': tc_peak, 'unit': 'TFLOP/s', 'frac': tf / tc_peak,
                              'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained' if peaks else 'fallback',
                              'hbm_gbs': hbm, 'note': 'useful fp32-equivalent FLOPs (SURVEY 8d); the kernel issues 3 fp16 MMAs per product (hi*hi + hi*lo + lo*hi) '
                                                      'to reach fp32 accuracy, so the tensor pipe does 3x this work'})
            by_kernel[kind] = entry
        # the env transition (Hex.step) of the real worlds: one launch per move, timed on its own
        by_kernel['hex_step'] = time_hex_step(play.worlds, hbm_peak, hbm_src)
        roofline = dict(by_kernel[dominant])
        roofline['kernel'] = {'descend_expand': 'descend + expand + env step (one launch per simulation)', 'net': 'fc_forward (leaf evaluation)',
                              'backup': 'backup + q-range'}[dominant]
        if dominant != 'net':
            roofline['note'] = ('bytes are SURVEY 8(d) algorithmic bytes; the descent is a chain of dependent fp32 additions in the reference order '
                                '(bit-exact parity), bounded by fp32-pipe issue and latency, not by HBM (DESIGN.md 5.1)')
        roofline['by_kernel'] = by_kernel
        roofline['ms_per_move_by_kernel'] = {k: round(v, 3) for k, v in split.items()}
        roofline['launches_per_move_by_kernel'] = per_move_launches
        roofline['algorithmic_bytes_per_move'] = {k: int(v) for k, v in bytes_by.items()}
        roofline['net_tflops'] = flops / (split['net'] / 1e3) / 1e12
        roofline['tree_shape'] = {'policy_evals_per_descent': counters[0] / max(counters[3], 1),
                                  'children_per_eval': counters[1] / max(counters[0], 1),
                                  'newton_iters_per_eval': counters[2] / max(counters[0], 1)}

        # the CPU and reference-CUDA legs load oracle/ (and oracle/_ref): they run in their own processes, after the timed regions, so that
        # this process — whose loaded libraries the driver records — holds the product library only
        cpu = run_leg('cpu_baseline', args.config) if (world == 1 and not args.no_cpu) else None
        refcuda = run_leg('reference_cuda', args.config) if (world == 1 and not args.no_cpu) else None
        extra = run_extras(args) if (world == 1 and not args.no_extra) else None
        result = {
            'metric': 'MCTS sims/sec', 'value': value, 'unit': 'sims/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32' if args.precision == 'fp32' else 'f16',
            'data': 'synthetic (random-playout positions, random-init FCModel weights with alpha~U(.1,.5))',
            'config': {'workload': f'{args.config}: {describe(args.config)}', 'envs_per_gpu': B, 'global_envs': B * world,
                       'n_nodes': T, 'boardsize': S, 'width': W, 'depth': D, 'net_precision': args.precision,
                       'parallelism': f'env-sharded x{world}' + (', per-move NCCL all-gather of trajectory records' if world > 1 else ''),
                       'l2_policy': 'tree workspace per GPU exceeds L2 (inputs larger than L2), no flush'
                       if B * T * (4 * A + 2 * A) > 126e6 else 'working set fits L2; not flushed'},
            'e2e': {'value': e2e, 'unit': 'sims/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': int(gpu_launches),
            'clocks': clocks.summary(),
            'roofline': roofline,
            'cpu_baseline': cpu,
            'reference_cuda': refcuda,
            'extra': extra,
        }
        if world > 1:
            result['config']['allgather_bytes_per_rank_per_move'] = B * record_width(A)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if result is not None:
        print(json.dumps(result))


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's CPU implementation of the path (oracle/_ref kernels when built, else the C port)
# ---------------------------------------------------------------------------------------------------------------
def cpu_baseline(config, sd, steps, warmup):
    import torch
    from oracle import build_ref, pyref, refpy
    # torchrun exports OMP_NUM_THREADS=1: the reference arm uses every host core it can, as a stand-alone reference process would
    torch.set_num_threads(os.cpu_count() or 1)
    S, B, T, W, D = CONFIGS[config]
    Bc = min(CPU_SAMPLE_ENVS.get(config, 512 if S <= 9 else 128), B)
    python = 'restated'
    if build_ref.available('O0') and refpy.present():
        # the reference's OWN Python (Hex, MCTSAgent, FCModel) on its own CPU kernels — only where /root/reference exists (not on the GPU box)
        r = refpy.load('O0')
        kind, python = 'reference', 'reference'
        torch.manual_seed(0)
        g = torch.Generator().manual_seed(0)
        w = r.Hex.initial(Bc, S, device='cpu')
        for _ in range(2 * S * S):
            w, _ = w.step(torch.multinomial(w.valid.float(), 1, generator=g).squeeze(-1))
        rnet = r.FCModel(w.obs_space, w.action_space, width=W, depth=D)
        rnet.load_state_dict(sd)
        ragent = r.MCTSAgent(rnet, n_nodes=T, c_puct=1 / 16)

        def move(w):
            d = ragent(w)
            w2, _ = w.step(d.actions)
            return w2
    else:
        if build_ref.available('O0'):
            ops, kind = pyref.RefOps('O0'), 'reference'
        else:
            ops, kind = pyref.COps(), 'port'
        torch.manual_seed(0)
        g = torch.Generator().manual_seed(0)
        w = pyref.HexWorld.initial(Bc, S, ops)
        for _ in range(2 * S * S):
            w, _ = w.step(torch.multinomial(w.valid.float(), 1, generator=g).squeeze(-1))
        net = pyref.FCNet(sd)

        def move(w):
            d = pyref.agent_call(w, net, n_nodes=T, c_puct=1 / 16)
            w2, _ = w.step(d.actions)
            return w2
    for _ in range(warmup):
        w = move(w)
    t0 = time.perf_counter()
    for _ in range(steps):
        w = move(w)
    dt = time.perf_counter() - t0
    return {'value': Bc * T * steps / dt, 'unit': 'sims/s', 'cores': torch.get_num_threads(), 'kind': kind, 'python': python,
            'host_cpus': os.cpu_count(), 'seconds': dt, 'sample_envs': Bc, 'warmup': warmup,
            'sample': f'{steps} moves of {Bc} envs (of {B}) at {describe(config)}; '
                      + ('reference CPU kernels built from its unmodified sources at the reference loader\'s flags (-O0), '
                         'single-threaded per-env loops as in the reference, torch ops on all threads; '
                         + ('the reference\'s own Python (Hex, MCTSAgent, FCModel) drives them'
                            if python == 'reference' else 'Python orchestration restated (oracle/pyref.py): /root/reference is not on this machine')
                         if kind == 'reference' else 'C restatement of the reference kernels (oracle/boardlaw_oracle.c, -O2)')}


def reference_cuda_baseline(config, sd, steps=2, warmup=1):
    """The reference's own CUDA kernels (oracle/_ref/CUDA: its unmodified sources compiled for sm_100a by its loader's recipe,
    oracle/build_ref.py) under the restated Python orchestration, on this GPU, on a bounded sample — SURVEY.md 8(c)/(d)'s
    "second baseline".  An extra, informational figure: the contract's reference arm stays the CPU one.  None when the build is
    absent."""
    import torch
    from oracle import build_ref, pyref
    if not build_ref.available('CUDA'):
        return None
    try:
        S, B, T, W, D = CONFIGS[config]
        Bc = min({'c3': 16384}.get(config, 32768), B)      # the whole batch where it fits: the reference's path is launch-bound, a small sample would flatter us
        dev = torch.device('cuda', torch.cuda.current_device())
        ops = pyref.RefCudaOps()
        torch.manual_seed(0)
        w = pyref.HexWorld.initial(Bc, S, ops, device=dev)
        for _ in range(2 * S * S):
            w, _ = w.step(torch.multinomial(w.valid.float(), 1).squeeze(-1))
        net = pyref.FCNet(sd, device=dev)

        def move(w):
            d = pyref.agent_call(w, net, n_nodes=T, c_puct=1 / 16)
            w2, _ = w.step(d.actions)
            return w2
        for _ in range(warmup):
            w = move(w)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            w = move(w)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return {'value': Bc * T * steps / dt, 'unit': 'sims/s', 'seconds': dt,
                'sample': f'{steps} moves of {Bc} envs (of {B}) at {describe(config)}: the reference\'s CUDA kernels (unmodified sources, '
                          'its loader\'s flags with -std=c++17) + torch ops under autocast as in its MCTS.simulate, Python orchestration '
                          'restated (oracle/pyref.py); wall clock with synchronize on both sides'}
    except Exception as e:                      # informational leg: never takes the bench line down
        return {'unavailable': f'{type(e).__name__}: {e}'[:300]}


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    import torch
    from boardlaw_b200.networks import synthetic_state_dict
    S, B, T, W, D = CONFIGS[args.config]
    sd = synthetic_state_dict(S, W, D, seed=0)
    t0 = time.perf_counter()
    # a bounded sample: at most 8 timed moves (~8 s each at c2) however many steps the GPU arm was asked for, same warm-up rule as ours
    steps = min(args.steps, 8)
    cpu = cpu_baseline(args.config, sd, steps=steps, warmup=min(max(args.warmup, 1), 3))
    Bc = CPU_SAMPLE_ENVS.get(args.config, 512)
    print(json.dumps({
        'impl': 'reference', 'metric': 'MCTS sims/sec', 'value': cpu['value'], 'unit': 'sims/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': cpu['warmup'], 'ms_per_step': cpu['seconds'] / steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'{args.config}: {describe(args.config)}', 'sample_envs': min(Bc, B), 'device': 'host CPU'},
        'cpu_baseline': {k: cpu[k] for k in ('kind', 'python', 'cores', 'sample', 'value', 'unit', 'host_cpus')},
        'e2e': {'value': cpu['value'], 'unit': 'sims/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--config', default='c2', choices=sorted(CONFIGS))
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default='fp32', choices=['fp32', 'amp'])
    ap.add_argument('--envs', type=int, default=0, help='override envs per GPU (debugging)')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline / reference_cuda legs')
    ap.add_argument('--no-extra', action='store_true', help='skip the other configs (c3, c5 sweep, c2 amp) reported under "extra"')
    ap.add_argument('--leg', default=None, choices=['cpu_baseline', 'reference_cuda'], help='(internal) run one baseline leg and print its JSON')
    args = ap.parse_args()
    if args.leg:
        from boardlaw_b200.networks import synthetic_state_dict
        S, B, T, W, D = CONFIGS[args.config]
        sd = synthetic_state_dict(S, W, D, seed=0)
        print(json.dumps(cpu_baseline(args.config, sd, steps=2, warmup=1) if args.leg == 'cpu_baseline' else reference_cuda_baseline(args.config, sd)))
        return
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
