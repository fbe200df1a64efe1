"""TEST INFRASTRUCTURE — CPU restatement of the reference's Python-level orchestration.

Restates, with plain tensors and no ``arrdict``:
  * ``Hex`` env record: initial / obs / valid / step with auto-reset  (boardlaw/hex/__init__.py:120-195)
  * ``FCModel.forward`` + heads                                       (boardlaw/networks.py:10-41, boardlaw/heads.py:41-52,93-104,128-142)
  * ``dirichlet_noise``                                               (boardlaw/mcts/__init__.py:13-24)
  * ``MCTS`` initialize / simulate / root / n_leaves, ``mcts()``      (boardlaw/mcts/__init__.py:27-152,200-207)
  * ``MCTSAgent.__call__``                                            (boardlaw/mcts/__init__.py:216-229)

The native ops come from a pluggable ``ops`` namespace: the C restatement (``oracle``), or the
reference's own CPU extension (``oracle/_ref`` via ``RefOps``).  Random numbers are consumed
from torch's global CPU generator in exactly the reference's order (Dirichlet draw at
``initialize``; one ``(B,T)`` half ``rand`` per ``simulate``; one Categorical draw per move) so
that seeded runs of this file and of the imported reference agree bit for bit
(tests/test_oracle_vs_ref.py, runs wherever /root/reference exists).

Third-party arithmetic: ATen CPU kernels of the installed torch (linear, log_softmax, tanh,
_sample_dirichlet, multinomial) — the same library calls the reference makes.
"""
import types

import numpy as np
import torch
import torch.distributions
from torch.nn import functional as F

import oracle as _c


# ------------------------------------------------------------------------------------------
# native-op backends
# ------------------------------------------------------------------------------------------

class COps:
    """The C restatement (oracle/boardlaw_oracle.c)."""
    name = 'port'

    def __init__(self, pow_mode=0):
        self.pow_mode = pow_mode

    step = staticmethod(_c.step)
    observe = staticmethod(_c.observe)
    MCTS = staticmethod(_c.MCTS)
    Backup = staticmethod(_c.Backup)
    backup = staticmethod(_c.backup)

    def descend(self, m, rands=None):
        return _c.descend(m, rands, pow_mode=self.pow_mode)

    def root(self, m):
        return _c.root(m, pow_mode=self.pow_mode)


class RefCudaOps:
    """The reference's own CUDA extension built for sm_100a from its unmodified sources (oracle/build_ref.py, variant CUDA: its
    loader's flags with -std=c++14 -> c++17, which torch 2.x headers require).  Used only by bench.py's
    ``reference_cuda`` leg to time the reference's kernels on the same GPU."""
    name = 'reference-cuda'

    def __init__(self):
        from . import build_ref
        self.hex = build_ref.load('hexcuda', 'CUDA')
        self.mcts = build_ref.load('mctscuda', 'CUDA')
        self.step, self.observe = self.hex.step, self.hex.observe
        self.MCTS, self.Backup, self.backup = self.mcts.MCTS, self.mcts.Backup, self.mcts.backup
        self.root = self.mcts.root

    def descend(self, m, rands=None):
        return self.mcts.descend(m)


class RefOps:
    """The reference's own CPU extension (oracle/_ref), unmodified sources."""
    name = 'reference'

    def __init__(self, variant='O0'):
        from . import build_ref
        self.hex = build_ref.load('hexcuda', variant)
        self.mcts = build_ref.load('mctscuda', variant)
        self.step, self.observe = self.hex.step, self.hex.observe
        self.MCTS, self.Backup, self.backup = self.mcts.MCTS, self.mcts.Backup, self.mcts.backup
        self.root = self.mcts.root

    def descend(self, m, rands=None):
        if rands is not None:
            raise ValueError('the reference draws its own rands (boardlaw/mcts/cpp/cpu.cpp:187)')
        return self.mcts.descend(m)


# ------------------------------------------------------------------------------------------
# Hex
# ------------------------------------------------------------------------------------------

class HexWorld:
    """board (B,S,S) u8 absolute coords; seats (B,) i32.  boardlaw/hex/__init__.py:120-195."""
    n_seats = 2

    def __init__(self, board, seats, ops=None):
        self.board, self.seats = board, seats
        self.ops = ops or COps()
        self.n_envs, self.boardsize = board.shape[0], board.shape[-1]
        self._obs = self._valid = None

    @classmethod
    def initial(cls, n_envs, boardsize=11, ops=None, device='cpu'):
        return cls(torch.zeros((n_envs, boardsize, boardsize), dtype=torch.uint8, device=device),
                   torch.zeros((n_envs,), dtype=torch.int32, device=device), ops)

    @property
    def obs(self):
        if self._obs is None:
            self._obs = self.ops.observe(self.board, self.seats)
        return self._obs

    @property
    def valid(self):
        if self._valid is None:                                   # hex/__init__.py:154-159
            self._valid = (self.obs == 0).all(-1).reshape(*self.board.shape[:-2], -1)
        return self._valid

    def step(self, actions, reset=True):                          # hex/__init__.py:161-195
        assert (0 <= actions).all(), 'You passed a negative action'
        assert actions.shape == (self.n_envs,)
        assert self.valid.gather(1, actions[:, None].long()).squeeze(-1).all()
        new_board = self.board.clone()
        rewards = self.ops.step(new_board, self.seats.int(), actions.int())
        if reset:
            terminal = (rewards > 0).any(-1)
        else:
            terminal = torch.zeros((self.n_envs,), dtype=torch.bool, device=self.board.device)
        new_board[terminal] = 0
        new_seat = 1 - self.seats
        new_seat[terminal] = 0
        return HexWorld(new_board, new_seat, self.ops), types.SimpleNamespace(terminal=terminal, rewards=rewards)

    def index(self, idx):
        return HexWorld(self.board[idx], self.seats[idx], self.ops)


class LazyWorld(HexWorld):
    """``boardlaw.hex.Lazy`` (boardlaw/hex/__init__.py:224-263) restated: one-player Hex, the opponent answers with the first legal
    move; a single reward column (the player's)."""
    n_seats = 1

    def step(self, actions):
        new, trans = HexWorld.step(self, actions)
        board, seats = new.board.clone(), new.seats.clone()
        rewards, terminal = trans.rewards.clone(), trans.terminal.clone()
        while True:
            mask = seats != self.seats
            if not mask.any():
                break
            sub = HexWorld(board[mask], seats[mask], self.ops)
            v = sub.valid
            n = v.shape[1]
            first = torch.where(v, torch.arange(n)[None].expand_as(v), torch.full_like(v, n, dtype=torch.long)).min(-1).values
            replied, other = HexWorld.step(sub, first)
            board[mask], seats[mask] = replied.board, replied.seats
            rewards[mask] += other.rewards
            terminal[mask] |= other.terminal
        envs = torch.arange(self.n_envs)
        return LazyWorld(board, seats, self.ops), types.SimpleNamespace(terminal=terminal, rewards=rewards[envs, self.seats.long()][:, None])


def random_actions(valid, uniforms):
    """The draw of bl_hex_random_transition restated: env b takes its k-th legal move (mover's frame order),
    k = min(floor(fp32(u_b) * fp32(n_legal)), n_legal - 1) — a uniform draw over the legal moves, i.e. the distribution of
    ``Categorical(probs=worlds.valid.float()).sample()`` (boardlaw/learning.py:8)."""
    import numpy as np
    v = valid.numpy().astype(bool)
    u = uniforms.numpy().astype(np.float32)
    n = v.sum(-1)
    k = np.minimum((u * n.astype(np.float32)).astype(np.int64), n - 1)
    k = np.maximum(k, 0)
    order = np.cumsum(v, -1) - 1                                   # rank of each legal move
    hit = v & (order == k[:, None])
    actions = np.where(hit.any(-1), hit.argmax(-1), -1)
    return torch.from_numpy(actions.astype(np.int64))


def random_playout(world, n_steps):
    """``learning.mix``-style decorrelation (boardlaw/learning.py:6-10): uniformly random valid moves."""
    for _ in range(n_steps):
        actions = torch.distributions.Categorical(probs=world.valid.float()).sample()
        world, _ = world.step(actions)
    return world


# ------------------------------------------------------------------------------------------
# network (fp32, as the reference runs it on CPU)
# ------------------------------------------------------------------------------------------

def fc_forward(sd, obs, valid, seats):
    """FCModel.forward on a state dict with the reference's keys:
    body.0.{weight,bias}; body.k.{weight,bias,α} k=1..D; policy.core.*; value.core.*"""
    B = obs.shape[0]
    x = F.linear(obs.reshape(B, -1), sd['body.0.weight'], sd['body.0.bias'])     # heads.py:47-52
    k = 1
    while f'body.{k}.weight' in sd:                                               # networks.py:17-18
        x = x + sd[f'body.{k}.α'] * F.linear(F.relu(x), sd[f'body.{k}.weight'], sd[f'body.{k}.bias'])
        k += 1
    y = F.linear(x, sd['policy.core.weight'], sd['policy.core.bias'])             # heads.py:101-104
    y = y.where(valid, torch.full_like(y, -np.inf))
    logits = F.log_softmax(y, -1)
    t = torch.tanh(F.linear(x, sd['value.core.weight'], sd['value.core.bias']).squeeze(-1))   # heads.py:136-142
    st = torch.stack([seats, 1 - seats], -1)
    vs = torch.stack([t, -t], -1)
    v = torch.full_like(vs, np.nan).scatter(-1, st.long(), vs)
    return types.SimpleNamespace(logits=logits, v=v)


class FCNet:
    """Callable network over HexWorld, holding a reference-format state dict."""

    def __init__(self, sd, device='cpu'):
        self.sd = {k: v.detach().float().to(device) for k, v in sd.items()}

    def __call__(self, world):
        # on a CUDA device (bench.py's reference-kernels-on-this-GPU leg only) under autocast, as mcts/__init__.py:131-132
        with torch.no_grad(), torch.autocast('cuda', enabled=world.board.device.type == 'cuda'):
            return fc_forward(self.sd, world.obs, world.valid, world.seats)


# ------------------------------------------------------------------------------------------
# MCTS
# ------------------------------------------------------------------------------------------

def dirichlet_noise(logits, valid, eps, alpha_scale=10):          # mcts/__init__.py:13-24
    alpha = alpha_scale / logits.size(-1)
    alpha = torch.full((valid.shape[-1],), alpha, dtype=torch.float, device=logits.device)
    draw = torch.distributions.Dirichlet(alpha).sample(logits.shape[:-1])
    draw[~valid] = 0.
    draw = draw / draw.sum(-1, keepdims=True)
    return (logits.exp() * (1 - eps) + draw * eps).log()


class Tree:
    """One search over all envs; tensor names and dtypes as boardlaw/mcts/__init__.py:29-70."""

    def __init__(self, world, n_nodes=64, c_puct=1 / 16, noise_eps=.25, alpha_scale=10, hooks=None):
        B, T, S = world.n_envs, n_nodes, world.boardsize
        A, Sn = S * S, world.n_seats
        self.ops = world.ops
        self.B, self.T, self.A, self.Sn = B, T, A, Sn
        dev = world.board.device
        self.envs = torch.arange(B, device=dev)
        self.children = torch.full((B, T, A), -1, dtype=torch.int16, device=dev)
        self.parents = torch.full((B, T), -1, dtype=torch.int16, device=dev)
        self.relation = torch.full((B, T), -1, dtype=torch.int16, device=dev)
        self.board = world.board[:, None].repeat(1, T, 1, 1).contiguous()
        self.seats = world.seats[:, None].repeat(1, T).contiguous()
        self.rewards = torch.zeros((B, T, Sn), dtype=torch.float16, device=dev)
        self.terminal = torch.zeros((B, T), dtype=torch.bool, device=dev)
        self.logits = torch.full((B, T, A), np.nan, dtype=torch.float16, device=dev)
        self.v = torch.full((B, T, Sn), np.nan, dtype=torch.float16, device=dev)
        self.n = torch.zeros((B, T), dtype=torch.int16, device=dev)
        self.w = torch.zeros((B, T, Sn), dtype=torch.float16, device=dev)
        self.c_puct = torch.full((B,), c_puct, dtype=torch.float16, device=dev)
        self.noise_eps, self.alpha_scale = noise_eps, alpha_scale
        self.sim = 0
        self.hooks = hooks

    def _m(self):
        return self.ops.MCTS(self.logits, self.w, self.n, self.c_puct, self.seats.short(), self.terminal, self.children)

    def initialize(self, network):                                # mcts/__init__.py:72-80
        world = HexWorld(self.board[:, 0], self.seats[:, 0], self.ops)
        d = network(world)
        self.logits[:, 0] = dirichlet_noise(d.logits, world.valid, self.noise_eps, self.alpha_scale)
        self.v[:, 0] = d.v
        self.sim += 1

    def simulate(self, network):                                  # mcts/__init__.py:108-140
        envs = self.envs
        if self.hooks and 'pre_descend' in self.hooks:
            self.hooks['pre_descend'](self)
        r = self.ops.descend(self._m())
        parents, actions = r.parents.long(), r.actions.long()
        if self.hooks and 'post_descend' in self.hooks:
            self.hooks['post_descend'](self, r)
        leaves = self.children[envs, parents, actions].long()
        leaves[leaves == -1] = self.sim
        self.children[envs, parents, actions] = leaves.short()
        self.parents[envs, leaves] = parents.short()
        self.relation[envs, leaves] = actions.short()

        old = HexWorld(self.board[envs, parents], self.seats[envs, parents], self.ops)
        world, trans = old.step(actions)
        self.board[envs, leaves] = world.board
        self.seats[envs, leaves] = world.seats
        self.rewards[envs, leaves] = trans.rewards.half()
        self.terminal[envs, leaves] = trans.terminal

        d = network(world)
        self.logits[envs, leaves] = d.logits.half()
        self.v[envs, leaves] = d.v.half()

        bk = self.ops.Backup(v=self.v, w=self.w, n=self.n, rewards=self.rewards, parents=self.parents,
                             terminal=self.terminal)
        self.ops.backup(bk, leaves.short())
        if self.hooks and 'post_backup' in self.hooks:
            self.hooks['post_backup'](self, leaves)
        self.sim += 1

    def root(self):                                               # mcts/__init__.py:142-149
        r = self.ops.root(self._m())
        return types.SimpleNamespace(logits=r.float().log().half(), prior=self.logits[:, 0], v=self.v[:, 0])

    def n_leaves(self):                                           # mcts/__init__.py:151-152
        return ((self.children == -1).all(-1) & (self.parents != -1)).sum(-1)


def mcts(world, network, **kwargs):                               # mcts/__init__.py:200-207
    m = Tree(world, **kwargs)
    m.initialize(network)
    for _ in range(m.T - 1):
        m.simulate(network)
    return m


def agent_call(world, network, eval=False, **kwargs):             # mcts/__init__.py:216-229
    m = mcts(world, network, **kwargs)
    r = m.root()
    if eval:
        actions = r.logits.argmax(-1)
    else:
        actions = torch.distributions.Categorical(logits=r.logits.float()).sample()
    return types.SimpleNamespace(
        logits=r.logits.clone(), prior=r.prior.clone(), n_sims=torch.full_like(m.envs, m.sim + 1),
        n_leaves=m.n_leaves(), v=r.v.clone(), actions=actions.clone(), tree=m)


# ------------------------------------------------------------------------------------------
# deterministic synthetic network weights (shared by tests, smoke and bench; no RNG state touched)
# ------------------------------------------------------------------------------------------

def synth_state_dict(S, W, D, seed=0):
    """FCModel-shaped random weights: torch-default-like uniform fan-in init for the plain layers,
    scaled-orthogonal-like gaussian for residual layers and α ~ U(0.1, 0.5) so that residual
    blocks are not the identity."""
    g = torch.Generator().manual_seed(seed)
    A = S * S

    def lin(o, i):
        bound = 1 / i ** .5
        return ((torch.rand((o, i), generator=g) * 2 - 1) * bound, (torch.rand((o,), generator=g) * 2 - 1) * bound)

    sd = {}
    sd['body.0.weight'], sd['body.0.bias'] = lin(W, 2 * A)
    for k in range(1, D + 1):
        sd[f'body.{k}.weight'] = torch.randn((W, W), generator=g) * (2 / W) ** .5
        sd[f'body.{k}.bias'] = (torch.rand((W,), generator=g) * 2 - 1) / W ** .5
        sd[f'body.{k}.α'] = torch.rand((), generator=g) * .4 + .1
    sd['policy.core.weight'], sd['policy.core.bias'] = lin(A, W)
    sd['value.core.weight'], sd['value.core.bias'] = lin(1, W)
    return sd


# ---------------------------------------------------------------------------------------------------------------
# arena (boardlaw/arena/common.py:50-106) — restated on HexWorld; pinned by tests/golden/arena.npz, which holds the results of
# the reference's own evaluate() on the same positions and agents (tests/golden/make_golden_arena.py)
# ---------------------------------------------------------------------------------------------------------------
def matchup_indices(n_envs, n_seats):                             # common.py:50-55
    from itertools import permutations
    patterns = torch.as_tensor(list(permutations(range(n_seats))))
    return patterns, patterns.repeat((n_envs // len(patterns), 1))


def evaluate(world, agents):                                      # common.py:75-106, 57-73
    """agents: list of (name, agent(world, eval=True) -> .actions).  Returns per seat pattern (names, wins per seat, moves,
    games)."""
    B = world.n_envs
    patterns, matchup = matchup_indices(B, world.n_seats)
    envs = torch.arange(B)
    terminal = torch.zeros(B, dtype=torch.bool)
    wins = torch.zeros((B, world.n_seats), dtype=torch.int)
    moves = torch.zeros(B, dtype=torch.int)
    board, seats = world.board.clone(), world.seats.clone()
    while not terminal.all():
        for i, (_, agent) in enumerate(agents):
            mask = (matchup[envs, seats.long()] == i) & ~terminal                       # common.py:90
            if mask.any():
                sub = HexWorld(board[mask], seats[mask], world.ops)
                new, trans = sub.step(agent(sub, eval=True).actions)                     # common.py:93-94
                board[mask], seats[mask] = new.board, new.seats
                terminal[mask] = trans.terminal
                wins[mask] += (trans.rewards == 1).int()                                # common.py:98-99
                moves[mask] += 1
    out = []
    for p in patterns:                                                                   # common.py:61-72
        rows = (matchup == p).all(-1)
        w = wins[rows].sum(0)
        out.append(types.SimpleNamespace(names=tuple(agents[int(i)][0] for i in p), wins=tuple(float(x) for x in w),
                                         moves=float(moves[rows].sum()), games=float(w.sum())))
    return out


# ---------------------------------------------------------------------------------------------------------------
# learner (boardlaw/learning.py:57-76, boardlaw/main.py:75-101) — restated on CPU tensors; reward_to_go is pinned against the
# reference's own function (tests/test_learner.py, where /root/reference exists) and its KATs (learning.py:83-94)
# ---------------------------------------------------------------------------------------------------------------
def present_value(deltas, fallback, terminal, alpha):            # learning.py:57-68
    result = torch.full_like(fallback, np.nan)
    result[-1] = fallback[-1]
    for t in range(deltas.size(0) - 1, -1, -1):
        result[t] = torch.where(terminal[t], fallback[t], deltas[t] + alpha * result[t + 1])
    return result


def reward_to_go(reward, value, terminal, gamma=1.):             # learning.py:70-76
    fallback = value.clone()                                      # (the reference writes into `value` itself)
    fallback[terminal] = reward[terminal]
    return present_value(reward[:-1], fallback, terminal, gamma)


def learner_loss(sd, world, target_logits, target_v):            # main.py:86-99, fp32
    """(policy_loss, value_loss) of the network ``sd`` on ``world`` with autograd enabled on ``sd``'s tensors."""
    d = fc_forward(sd, world.obs, world.valid, world.seats)
    zeros = torch.zeros_like(d.logits)
    l = d.logits.where(d.logits > -np.inf, zeros)
    l0 = target_logits.float().where(target_logits > -np.inf, zeros)
    policy_loss = -(l0.exp() * l).sum(-1).mean()
    value_loss = (target_v.float() - d.v).square().mean()
    return policy_loss, value_loss


def learner_step(sd, world, target_logits, target_v, lr=1e-3, steps=1):
    """``steps`` iterations of main.optimize's update with torch autograd + torch.optim.Adam in fp32 on the CPU.  Returns
    (new state dict, gradients of the first step, (policy_loss, value_loss) of the first step)."""
    params = {k: v.detach().clone().float().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.Adam(list(params.values()), lr=lr)
    first = None
    for _ in range(steps):
        pl, vl = learner_loss(params, world, target_logits, target_v)
        opt.zero_grad()
        (pl + vl).backward()
        if first is None:
            first = ({k: p.grad.detach().clone() for k, p in params.items()}, (pl.detach(), vl.detach()))
        opt.step()
    return {k: p.detach() for k, p in params.items()}, first[0], first[1]
