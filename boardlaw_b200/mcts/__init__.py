"""Monte-Carlo tree search over a batch of envs — the surface of ``boardlaw.mcts``
(boardlaw/mcts/__init__.py:13-241): ``dirichlet_noise``, ``MCTS``, ``mcts()``, ``MCTSAgent``.

Two execution paths sit behind the same surface:

  * ``MCTS`` keeps the reference's tensors (``tree``, ``worlds``, ``transitions``, ``decisions``, ``stats``) and
    drives the op-level kernels (``mcts.cuda``) one simulation at a time with any ``network(worlds)`` callable
    and any world type — this is what arena-style callers and the toy-world tests use.
  * ``engine.SearchEngine`` (picked by ``mcts()`` / ``MCTSAgent`` for a Hex world with an ``FCModel``) runs the
    whole move on a persistent private workspace with O(1) host involvement.
"""
import numpy as np
import torch
import torch.distributions

from .. import arrdict
from . import cuda


def dirichlet_noise(logits, valid, eps, alpha_scale=10):
    """Mixes a Dirichlet(alpha_scale/A) draw, restricted to the valid actions and renormalised, into the prior
    in probability space (boardlaw/mcts/__init__.py:13-24)."""
    conc = torch.full((valid.shape[-1],), alpha_scale / logits.size(-1), dtype=torch.float, device=logits.device)
    draw = torch.distributions.Dirichlet(conc).sample(logits.shape[:-1])
    draw[~valid] = 0.
    draw = draw / draw.sum(-1, keepdims=True)
    return (logits.exp() * (1 - eps) + draw * eps).log()


def dirichlet_mix(logits, valid, eps, alpha_scale=10, draw=None):
    """``dirichlet_noise`` with the Dirichlet draw optionally injected (``draw`` (B,A) f32, the raw sample before
    masking) so that seeded runs can be compared across devices."""
    if draw is None:
        conc = torch.full((valid.shape[-1],), alpha_scale / logits.size(-1), dtype=torch.float, device=logits.device)
        draw = torch.distributions.Dirichlet(conc).sample(logits.shape[:-1])
    else:
        draw = draw.to(logits.device, torch.float).clone()
    draw[~valid] = 0.
    draw = draw / draw.sum(-1, keepdims=True)
    return (logits.exp() * (1 - eps) + draw * eps).log()


class MCTS:

    def __init__(self, world, n_nodes=64, c_puct=1 / 16, noise_eps=.25, alpha_scale=10):
        """c_puct high: concentrates on the prior; low: concentrates on value."""
        assert n_nodes > 0, 'MCTS requires at least one node'
        self.device = world.device
        self.n_envs = world.n_envs
        self.n_nodes = n_nodes
        self.n_seats = world.n_seats
        self.n_actions = int(np.prod(world.action_space))
        B, T, A, Sn, dev = self.n_envs, n_nodes, self.n_actions, self.n_seats, self.device

        self.envs = torch.arange(B, device=dev)
        self.tree = arrdict.arrdict(
            children=torch.full((B, T, A), -1, device=dev, dtype=torch.short),
            parents=torch.full((B, T), -1, device=dev, dtype=torch.short),
            relation=torch.full((B, T), -1, device=dev, dtype=torch.short))
        self.worlds = arrdict.stack([world for _ in range(T)], 1)
        self.transitions = arrdict.arrdict(
            rewards=torch.zeros((B, T, Sn), device=dev, dtype=torch.half),
            terminal=torch.zeros((B, T), device=dev, dtype=torch.bool))
        self.decisions = arrdict.arrdict(
            logits=torch.full((B, T, A), np.nan, device=dev, dtype=torch.half),
            v=torch.full((B, T, Sn), np.nan, device=dev, dtype=torch.half))
        self.stats = arrdict.arrdict(
            n=torch.zeros((B, T), device=dev, dtype=torch.short),
            w=torch.zeros((B, T, Sn), device=dev, dtype=torch.half))
        self.sim = 0
        self.c_puct = torch.full((B,), c_puct, device=dev, dtype=torch.half)
        self.noise_eps = noise_eps
        self.alpha_scale = alpha_scale

    def initialize(self, network, noise=None):
        world = self.worlds[:, 0]
        with torch.no_grad():
            decisions = network(world)
        self.decisions.logits[:, self.sim] = dirichlet_mix(decisions.logits, world.valid, self.noise_eps, self.alpha_scale, noise)
        self.decisions.v[:, 0] = decisions.v
        self.sim += 1

    def _cuda(self):
        return cuda.mcts(self.decisions.logits, self.stats.w, self.stats.n, self.c_puct, self.worlds.seats,
                         self.transitions.terminal, self.tree.children)

    def descend(self, rands=None):
        result = cuda.descend(self._cuda(), rands)
        return result.parents.long(), result.actions.long()

    def backup(self, leaves):
        bk = cuda.Backup(v=self.decisions.v, w=self.stats.w, n=self.stats.n, rewards=self.transitions.rewards,
                         parents=self.tree.parents, terminal=self.transitions.terminal)
        cuda.backup(bk, leaves.short())

    def simulate(self, network, rands=None):
        if self.sim >= self.n_nodes:
            raise ValueError('Called simulate more times than were declared in the constructor')
        envs = self.envs
        parents, actions = self.descend(rands)

        # a descent that stopped at a terminal child re-uses that child's slot (boardlaw/mcts/__init__.py:115-118)
        leaves = self.tree.children[envs, parents, actions].long()
        leaves[leaves == -1] = self.sim
        self.tree.children[envs, parents, actions] = leaves.short()
        self.tree.parents[envs, leaves] = parents.short()
        self.tree.relation[envs, leaves] = actions.short()

        world, transition = self.worlds[envs, parents].step(actions)
        self.worlds[envs, leaves] = world
        self.transitions.rewards[envs, leaves] = transition.rewards.half()
        self.transitions.terminal[envs, leaves] = transition.terminal

        with torch.no_grad():
            decisions = network(world)
        self.decisions.logits[envs, leaves] = decisions.logits.half()
        self.decisions.v[envs, leaves] = decisions.v.half()

        self.backup(leaves)
        self.sim += 1

    def root(self):
        r = cuda.root(self._cuda())
        # log of the half probabilities through the host-evaluated table: identical to the reference's CPU path
        # `r.float().log().half()` (boardlaw/mcts/__init__.py:147) for every input, on any device
        from .. import _lib
        return arrdict.arrdict(
            logits=_lib.log_lut(r.device)[r.view(torch.int16).long() & 0xFFFF],
            prior=self.decisions.logits[:, 0],
            v=self.decisions.v[:, 0])

    def n_leaves(self):
        return ((self.tree.children == -1).all(-1) & (self.tree.parents != -1)).sum(-1)


class EngineSearch:
    """What ``mcts()`` returns on the fused path: the ``root()`` / ``n_leaves()`` / ``sim`` / ``envs`` members
    ``MCTSAgent`` uses, plus the reference's tree tensors materialised on demand from the engine workspace."""

    def __init__(self, eng, outputs):
        self.engine = eng
        self.device = eng.device
        self.n_envs, self.n_nodes, self.n_seats, self.n_actions = outputs[0].shape[0], eng.T, eng.Sn, eng.A   # <= the workspace's capacity
        self.envs = torch.arange(self.n_envs, device=eng.device)
        self.sim = eng.T
        self._out = outputs

    def root(self):
        logits, prior, v, _ = self._out
        return arrdict.arrdict(logits=logits, prior=prior, v=v)

    def n_leaves(self):
        return self._out[3]

    @property
    def tree(self):
        ws, n = self.engine.ws, self.n_envs
        return arrdict.arrdict(children=self.engine.children_dense()[:n], parents=ws.parents[:n], relation=ws.relation[:n])

    @property
    def stats(self):
        return arrdict.arrdict(n=self.engine.ws.n[:self.n_envs], w=self.engine.ws.w[:self.n_envs])

    @property
    def transitions(self):
        return arrdict.arrdict(rewards=self.engine.ws.rewards[:self.n_envs], terminal=self.engine.ws.terminal[:self.n_envs].bool())


_engines = {}


def engine_for(world, n_nodes, seed=0):
    """Workspaces are persistent: one per (device, capacity, boardsize, n_nodes), reused move after move.  A batch smaller than
    an existing workspace of the same shape runs inside it (``SearchEngine._search_partial``): arena-style callers, whose
    sub-batch shrinks every move as games end, allocate once."""
    from ..engine import SearchEngine
    key = (str(world.device), world.n_envs, world.boardsize, n_nodes, seed)
    if key in _engines:
        return _engines[key]
    fits = [k for k in _engines if k[0] == key[0] and k[2:] == key[2:] and world.n_envs < k[1] <= 4 * max(world.n_envs, 256)]
    if fits:
        return _engines[min(fits, key=lambda k: k[1])]
    if len(_engines) >= 4:
        _engines.pop(next(iter(_engines))).release()      # its CUDA graphs would keep the whole workspace alive
    _engines[key] = SearchEngine(world.n_envs, world.boardsize, n_nodes, world.device, seed=seed)
    return _engines[key]


def _fusable(worlds, network):
    from ..hex import Hex
    from ..networks import FCModel
    # the engine's node records hold actions in a byte and node ids in 16 bits (check_tree, csrc/engine.cu)
    return (isinstance(worlds, Hex) and isinstance(network, FCModel) and worlds.board.ndim == 3 and worlds.device.type == 'cuda'
            and worlds.boardsize ** 2 <= 255)


def mcts(worlds, network, n_nodes=64, c_puct=1 / 16, noise_eps=.25, alpha_scale=10, fused=None, engine_seed=0, **kwargs):
    """boardlaw/mcts/__init__.py:200-207.  Hex worlds with an FCModel run on the fused engine; anything else (toy
    worlds, arbitrary network callables, boards of 16x16 and more) runs the op-level ``MCTS`` loop.  ``engine_seed`` keys the
    engine's in-kernel random stream (one per rank in multi-process runs)."""
    fused = _fusable(worlds, network) if fused is None else fused
    if fused and 1 < n_nodes <= 32767:
        eng = engine_for(worlds, n_nodes, engine_seed)
        out = eng.search(worlds.board, worlds.seats, network, c_puct=c_puct, noise_eps=noise_eps, alpha_scale=alpha_scale, **kwargs)
        return EngineSearch(eng, out)
    m = MCTS(worlds, n_nodes=n_nodes, c_puct=c_puct, noise_eps=noise_eps, alpha_scale=alpha_scale)
    m.initialize(network)
    for _ in range(m.n_nodes - 1):
        m.simulate(network)
    return m


class MCTSAgent:

    def __init__(self, network, **kwargs):
        self.network = network
        self.kwargs = kwargs

    def __call__(self, world, value=True, eval=False, **kwargs):
        m = mcts(world, self.network, **{**self.kwargs, **kwargs})
        r = m.root()
        # sampling wants floats (boardlaw/mcts/__init__.py:220-221)
        actions = r.logits.argmax(-1) if eval else torch.distributions.Categorical(logits=r.logits.float()).sample()
        return arrdict.arrdict(
            logits=r.logits,
            prior=r.prior,
            n_sims=torch.full_like(m.envs, m.sim + 1),
            n_leaves=m.n_leaves(),
            v=r.v,
            actions=actions).clone()

    def play(self, world, eval=False, record=False, **kwargs):
        """One whole move — ``decisions = agent(world)`` and ``world.step(decisions.actions)`` — as ONE captured CUDA graph on the
        fused engine (``SearchEngine.play``): in-kernel Dirichlet noise, search, action, env transition and, with ``record``, the packed
        trajectory record.  Returns (decisions, new_world, transitions[, record]) as VIEWS of the engine's static buffers, valid
        until the next call (clone what must outlive it).  Falls back to the two separate calls for worlds / networks the engine does
        not take."""
        from ..hex import Hex
        kw = {**self.kwargs, **kwargs}
        n_nodes = kw.get('n_nodes', 64)
        if not (_fusable(world, self.network) and type(world) is Hex and 1 < n_nodes <= 32767):
            d = self(world, eval=eval, **kwargs)
            new_world, transitions = world.step(d.actions)
            return (d, new_world, transitions, None) if record else (d, new_world, transitions)
        eng = engine_for(world, n_nodes, kw.get('engine_seed', 0))
        logits, prior, v, n_leaves, actions, nb, ns, rewards, terminal, rec = eng.play(
            world.board, world.seats, self.network, c_puct=kw.get('c_puct', 1 / 16), noise_eps=kw.get('noise_eps', .25),
            alpha_scale=kw.get('alpha_scale', 10), greedy=eval, record=record)
        decisions = arrdict.arrdict(logits=logits, prior=prior, n_sims=eng.n_sims, n_leaves=n_leaves, v=v, actions=actions)
        new_world = Hex(board=nb, seats=ns)
        new_world.errors = eng.errors
        transitions = arrdict.arrdict(terminal=terminal, rewards=rewards)
        return (decisions, new_world, transitions, rec) if record else (decisions, new_world, transitions)

    def load_state_dict(self, sd):
        network = {k[8:]: v for k, v in sd.items() if k.startswith('network.')}
        kwargs = {k[7:]: v for k, v in sd.items() if k.startswith('kwargs.')}
        self.network.load_state_dict(network)
        self.kwargs.update(kwargs)

    def state_dict(self):
        network = {f'network.{k}': v for k, v in self.network.state_dict().items()}
        kwargs = {f'kwargs.{k}': v for k, v in self.kwargs.items()}
        return {**network, **kwargs}


class DummyAgent:
    """Acts straight from the network's policy (boardlaw/mcts/__init__.py:243-258)."""

    def __init__(self, network):
        self.network = network

    def __call__(self, world, eval=False):
        r = self.network(world)
        actions = r.logits.argmax(-1) if eval else torch.distributions.Categorical(logits=r.logits.float()).sample()
        return arrdict.arrdict(
            logits=r.logits, prior=r.logits,
            n_sims=torch.full((world.n_envs,), 0, device=world.device),
            n_leaves=torch.full((world.n_envs,), 1, device=world.device),
            v=r.v, actions=actions).clone()
