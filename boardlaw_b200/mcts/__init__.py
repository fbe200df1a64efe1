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

    def initialize(self, network):
        world = self.worlds[:, 0]
        with torch.no_grad():
            decisions = network(world)
        self.decisions.logits[:, self.sim] = dirichlet_noise(decisions.logits, world.valid, self.noise_eps, self.alpha_scale)
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
        return arrdict.arrdict(
            logits=r.float().log().half(),
            prior=self.decisions.logits[:, 0],
            v=self.decisions.v[:, 0])

    def n_leaves(self):
        return ((self.tree.children == -1).all(-1) & (self.tree.parents != -1)).sum(-1)


def mcts(worlds, network, **kwargs):
    m = MCTS(worlds, **kwargs)
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
