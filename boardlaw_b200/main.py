"""The actor/learner loop of ``boardlaw/main.py:147-205`` (``run``) on the B200 path, without the reference's storage,
statistics and live-arena side channels: decorrelate fresh worlds with random playouts, collect ``buffer_len`` moves of
self-play, build the chunk (reward-to-go), take one optimiser step on one (t, env) sample per env, repeat.

Multi-GPU (one process per GPU under ``torchrun``): every rank plays its own shard of envs and each move's trajectory records
are all-gathered (``selfplay.TrajectoryPool``), so every rank holds the whole chunk.  The optimiser step is data-parallel by default
(``learner='sharded'``): a rank runs forward / backward on ITS shard of the step's samples and the flat gradient buffer is averaged
with one all-reduce (1.7 MB at W256 D4), so the step costs the same at any number of GPUs; ``learner='replicated'`` runs the whole
batch on every rank instead (no gradient exchange; n_gpus times the work).  Either way every replica applies the same update to the
same weights — no weight broadcast is needed, and ``selfplay.check_replicas`` verifies it.
"""
import torch

from . import arrdict, learner, learning, selfplay
from .hex import Hex
from .mcts import MCTSAgent
from .networks import FCModel


def setup(boardsize, width, depth, nodes=64, c_puct=1 / 16, n_envs=32 * 1024, mix_steps=None, device='cuda', seed=0, rank=0):
    """(worlds, network, agent) of one rank.  What must be IDENTICAL on every rank is drawn from ``seed``: the network's initial
    weights (and, in ``run``, the chunk sampler's indices); what must DIFFER is drawn from ``seed + 1 + rank``: the decorrelating
    playouts, the Dirichlet noise, the sampled actions and the engine's in-kernel random stream — otherwise every rank would play
    the same games and the gathered chunk would hold world_size copies of one shard."""
    torch.manual_seed(seed)
    probe = Hex.initial(1, boardsize, device=device)
    network = FCModel(probe.obs_space, probe.action_space, width=width, depth=depth).to(probe.device)
    torch.manual_seed(seed + 1 + rank)
    worlds = learning.mix(Hex.initial(n_envs, boardsize, device=device), T=2500 if mix_steps is None else mix_steps)   # main.py:150
    agent = MCTSAgent(network, n_nodes=nodes, c_puct=c_puct, engine_seed=seed + 1 + rank)
    return worlds, network, agent


def run(boardsize, width, depth, nodes=64, c_puct=1 / 16, lr=1e-3, n_envs=32 * 1024, buffer_len=64, mix_steps=None, max_steps=1,
        device='cuda', pool=None, seed=0, on_step=None, learner_mode='sharded'):
    """Returns (agent, list of arrdict(policy_loss, value_loss) per optimiser step).  ``n_envs`` is per rank; ``pool`` a
    ``selfplay.TrajectoryPool`` when running under torch.distributed (None: single process)."""
    pool = pool or selfplay.TrajectoryPool()
    world_size, rank = pool.world, pool.rank
    worlds, network, agent = setup(boardsize, width, depth, nodes, c_puct, n_envs, mix_steps, device, seed, rank)
    L = learner.Learner(network, lr=lr)
    g = torch.Generator(device=worlds.device).manual_seed(seed)                   # same draw on every rank
    n_all = n_envs * world_size
    idxs = (torch.randint(buffer_len, (n_all,), device=worlds.device, generator=g), torch.arange(n_all, device=worlds.device))   # main.py:170
    records, losses = [], []
    for step in range(max_steps):
        while len(records) < buffer_len:                                            # main.py:173-185
            pool.sync_before_overwrite()
            decisions, new_worlds, transition, rec = agent.play(worlds, record=True)      # the whole move: one captured graph
            pool.gather(rec if rec is not None else selfplay.pack_records(worlds, decisions, transition))
            records.append(pool.wait().clone())
            worlds = new_worlds
        worlds.check()                                                              # rule violations recorded on the device by Hex.step
        if world_size > 1 and learner_mode == 'sharded':
            lo = rank * n_envs                                                      # this rank's envs inside the gathered chunk
            chunk, records = learner.chunk_from_records(records, boardsize, n_envs, shard=rank)
            out = L.optimize(chunk[(idxs[0][lo:lo + n_envs], idxs[1][:n_envs])], group=pool.group, world=world_size)
        else:
            chunk, records = learner.chunk_from_records(records, boardsize, n_all)    # main.py:188
            out = L.optimize(chunk[idxs])                                              # main.py:189
        selfplay.check_replicas(network, pool)                                      # every rank applied the same update to the same weights
        losses.append(out)
        if on_step is not None:
            on_step(step, agent, out)
    return agent, losses
