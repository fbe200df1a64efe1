"""The fused search engine: a persistent, privately laid out tree workspace driven through the ``bl_tree_*``
entry points of libboardlaw_b200.so (include/boardlaw_b200.h), replacing the per-move allocation and the
~120 kernel launches + host syncs per simulation of ``MCTS.simulate`` (boardlaw/mcts/__init__.py:108-140).

One ``SearchEngine`` owns the workspace for a fixed (n_envs, boardsize, n_nodes) on one device; the per-move
work is three launches per simulation (descend+expand+step, network, backup+q-range), captured in a CUDA
graph after the first move.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib, arrdict
from ._lib import ptr, check


def _round_up(x, m):
    return (x + m - 1) // m * m


class SearchEngine:

    def __init__(self, n_envs, boardsize, n_nodes, device, mirror_logits=False, seed=0, fast_descent=None):
        self.B, self.S, self.T = n_envs, boardsize, n_nodes
        self.A, self.Sn = boardsize * boardsize, 2
        self.AP, self.BP = _round_up(self.A, 4), _round_up(self.A, 16)
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('SearchEngine runs on a CUDA device only (there is no CPU path)')
        self.seed = seed
        B, T, A, Sn, dev = self.B, self.T, self.A, self.Sn, self.device
        z = lambda shape, dtype: torch.zeros(shape, dtype=dtype, device=dev)
        # two 16-byte records per (env, node) — bl_node / bl_aux of include/boardlaw_b200.h — so that a node visit is
        # one 128-bit load; the reference-named tensors below are strided views into them
        node = z((B, T, 8), torch.int16)
        aux = z((B, T, 8), torch.int16)
        node_bytes = node.view(torch.uint8)
        self.ws = arrdict.arrdict(
            pi=z((B, T, self.AP), torch.float32),
            board=z((B, T, self.BP), torch.uint8),
            node=node, aux=aux, parent_of=z((B, _round_up(T, 8)), torch.int16), kids=z((B, T, (T + 63) // 64), torch.int64),
            c_puct=z((B,), torch.float16),
            leaf=z((B,), torch.int16), leaf_parent=z((B,), torch.int16), leaf_action=z((B,), torch.int16),
            leaf_v=z((B, Sn), torch.float16),
            prior=z((B, A), torch.float16), qrange=z((T + 1, 2), torch.float32), counters=z((16,), torch.int64))
        # the certified fast descent (variant 5) samples from the prefix sums of the pi rows: three more arrays
        if fast_descent is None:
            fast_descent = os.environ.get('BL_DESCEND_VARIANT') in ('5', '6')
        if fast_descent:
            self.ws.update(cpi=z((B, T, self.AP), torch.float32), psum=z((B, T), torch.float32), cprior=z((B, T), torch.float32))
        if mirror_logits:
            self.ws['logits'] = torch.full((B, T, A), np.nan, dtype=torch.float16, device=dev)
        views = arrdict.arrdict(
            parents=node[:, :, 0], relation=node[:, :, 1], first_child=node[:, :, 2], next_sib=node[:, :, 3],
            n=node[:, :, 4], w=node[:, :, 5:7].view(torch.float16),
            seats=node_bytes[:, :, 14], terminal=node_bytes[:, :, 15],
            rewards=aux[:, :, 0:2].view(torch.float16), v=aux[:, :, 2:4].view(torch.float16))
        self.exp_lut = _lib.exp_lut(dev)
        self.log_lut = _lib.log_lut(dev)
        fields = {k: self.ws[k].data_ptr() for k in self.ws}
        for k in ('logits', 'cpi', 'psum', 'cprior'):
            fields.setdefault(k, None)
        self.ctree = _lib.Tree(B=B, T=T, S=self.S, A=A, Sn=Sn, AP=self.AP, BP=self.BP, exp_lut=self.exp_lut.data_ptr(),
                                scratch=None, scratch_bytes=0, **fields)
        nscratch = int(_lib.lib().bl_tree_scratch_bytes(ctypes.byref(self.ctree)))
        self.scratch = torch.empty((max(nscratch, 16),), dtype=torch.uint8, device=dev)
        self.ctree.scratch, self.ctree.scratch_bytes = self.scratch.data_ptr(), self.scratch.numel()
        self.ws.update(views)
        self._tp = ctypes.byref(self.ctree)
        # static I/O buffers (graph-replay safe)
        self.in_board = z((B, self.S, self.S), torch.uint8)
        self.in_seats = z((B,), torch.int32)
        self.root_logits = z((B, A), torch.float32)
        self.root_v = z((B, 2), torch.float32)
        self.out_logits = z((B, A), torch.float16)
        self.out_v = z((B, 2), torch.float16)
        self.out_n_leaves = z((B,), torch.int64)
        # whole-move outputs (``play``): the agent's actions, the stepped worlds, the transition, the packed trajectory record
        self.out_actions = z((B,), torch.int64)
        self.nx_board = z((B, self.S, self.S), torch.uint8)
        self.nx_seats = z((B,), torch.int32)
        self.out_rewards = z((B, 2), torch.float32)
        self.out_terminal = z((B,), torch.bool)
        self.record_width = (5 * A + 12 + 15) // 16 * 16
        self.out_record = z((B, self.record_width), torch.uint8)
        self.errors = z((), torch.int32)              # rule violations of every env transition played through ``play``
        self.n_sims = torch.full((B,), n_nodes + 1, dtype=torch.int64, device=dev)
        self._scratch = None
        self._scratch_key = None
        self._graphs = {}
        self._graph_refs = {}      # key -> the staged operand tensors the graphs of that key point into
        self._graph_lru = []       # least recently used first; at most MAX_GRAPH_KEYS (network, c_puct) combinations stay captured
        self.sim = 0
        self.move = 0
        self.launches = 0          # kernels of libboardlaw_b200.so launched (or replayed from a graph) so far
        self._graph_launches = {}

    MAX_GRAPH_KEYS = 4

    def _touch(self, key):
        if key in self._graph_lru:
            self._graph_lru.remove(key)
        self._graph_lru.append(key)
        while len(self._graph_lru) > self.MAX_GRAPH_KEYS:
            old = self._graph_lru.pop(0)
            for k in [k for k in self._graphs if k[-len(old):] == old]:
                self._graphs.pop(k, None)
                self._graph_launches.pop(k, None)
            self._graph_refs.pop(old, None)

    def _move_index(self):
        return self.ws.counters[6:7].clone()

    def _restore_move_index(self, saved):
        """The reset kernel advances the device-side move counter (it keys the in-kernel random streams); warm-up and capture passes
        must not: a search is exactly one step of the counter whether it runs eagerly, captures, or replays."""
        self.ws.counters[6:7].copy_(saved)

    def release(self):
        """Drops the captured graphs (they pin the workspace and the networks' staged operands)."""
        self._graphs.clear(); self._graph_launches.clear(); self._graph_refs.clear(); self._graph_lru.clear()

    # ---- thin wrappers over the C ABI ------------------------------------------------------------------
    def _stream(self):
        return _lib.stream_for(self.device)

    def reset(self, board, seats, c_puct):
        self.in_board.copy_(board)
        self.in_seats.copy_(seats)
        self._reset(c_puct)

    def _reset(self, c_puct):
        check(_lib.lib().bl_tree_reset(self._tp, ptr(self.in_board), ptr(self.in_seats), float(c_puct), self._stream()),
              'bl_tree_reset')
        self.sim = 0
        self.launches += 1

    def scratch_for(self, cparams):
        key = (cparams.W, cparams.D)
        if self._scratch_key != key:
            n = _lib.lib().bl_tree_eval_scratch_bytes(self._tp, ctypes.byref(cparams))
            self._scratch = torch.empty((max(int(n), 1),), dtype=torch.uint8, device=self.device)
            self._scratch_key = key
        return self._scratch

    def eval_root(self, cparams):
        check(_lib.lib().bl_tree_eval_root(self._tp, ctypes.byref(cparams), ptr(self.root_logits), ptr(self.root_v),
                                           ptr(self.scratch_for(cparams)), self._stream()), 'bl_tree_eval_root')
        self.launches += 2 if _lib.lib().bl_fc_uses_tensor_cores(ctypes.byref(cparams)) else cparams.D + 4
        return self.root_logits, self.root_v

    def set_eval(self, node, logits, v):
        half = logits.dtype == torch.float16
        assert logits.dtype in (torch.float16, torch.float32) and v.dtype == logits.dtype
        assert logits.shape == (self.B, self.A) and v.shape == (self.B, 2)
        logits, v = logits.contiguous(), v.contiguous()
        check(_lib.lib().bl_tree_set_eval(self._tp, node, ptr(logits), ptr(v), int(half), self._stream()), 'bl_tree_set_eval')
        self.launches += 1
        if node == 0:
            self.sim = 1

    def descend_expand(self, sim, rands=None):
        if rands is not None:
            rands = _lib.proxy(rands.contiguous(), torch.float16, 2, 'rands')
            assert rands.shape == (self.B, self.T)
        check(_lib.lib().bl_tree_descend_expand(self._tp, sim, ptr(rands), self.seed, self._stream()), 'bl_tree_descend_expand')
        self.launches += 1

    def eval_leaves(self, cparams, sim):
        check(_lib.lib().bl_tree_eval_leaves(self._tp, ctypes.byref(cparams), sim, ptr(self.scratch_for(cparams)),
                                             self._stream()), 'bl_tree_eval_leaves')
        # tensor-core path: one fused kernel; CUDA-core path: gather + network (D+3) + set_eval
        self.launches += 1 if _lib.lib().bl_fc_uses_tensor_cores(ctypes.byref(cparams)) else cparams.D + 5

    def backup(self, sim):
        check(_lib.lib().bl_tree_backup(self._tp, sim, self._stream()), 'bl_tree_backup')
        self.launches += 1
        self.sim = sim + 1

    def root(self, sim=None):
        sim = self.sim if sim is None else sim
        check(_lib.lib().bl_tree_root(self._tp, sim, ptr(self.log_lut), ptr(self.out_logits), ptr(self.out_v),
                                      ptr(self.out_n_leaves), self._stream()), 'bl_tree_root')
        self.launches += 1
        return self.out_logits, self.out_v, self.out_n_leaves

    def set_root_prior(self, noise_eps, alpha_scale, draw=None):
        """Node 0's evaluation = the root evaluation mixed with Dirichlet noise, in one launch (the draw injected, or in-kernel)."""
        if draw is not None:
            draw = draw.to(self.device, torch.float32).contiguous()
            assert draw.shape == (self.B, self.A)
        check(_lib.lib().bl_tree_set_root_prior(self._tp, ptr(self.root_logits), ptr(self.root_v), ptr(self.in_board), ptr(self.in_seats),
                                                ptr(draw), float(noise_eps), float(alpha_scale), self.seed, self._stream()), 'bl_tree_set_root_prior')
        self.launches += 1
        self.sim = 1

    def root_act(self, sim=None, greedy=False, uniforms=None):
        """``root`` plus the agent's action per env (argmax when ``greedy``, else a draw from the root policy)."""
        sim = self.sim if sim is None else sim
        if uniforms is not None:
            uniforms = uniforms.to(self.device, torch.float32).contiguous()
        check(_lib.lib().bl_tree_root_act(self._tp, sim, ptr(self.log_lut), ptr(self.out_logits), ptr(self.out_v), ptr(self.out_n_leaves),
                                          ptr(self.out_actions), ptr(uniforms), int(bool(greedy)), self.seed, self._stream()), 'bl_tree_root_act')
        self.launches += 1
        return self.out_logits, self.out_v, self.out_n_leaves, self.out_actions

    def step_and_record(self, record=True):
        """The env transition of the searched worlds under ``out_actions`` and the packed trajectory record of the move."""
        l = _lib.lib()
        check(l.bl_hex_transition(ptr(self.in_board), ptr(self.in_seats), ptr(self.out_actions), ptr(self.nx_board), ptr(self.nx_seats),
                                  ptr(self.out_rewards), ptr(self.out_terminal), ptr(self.errors), 1, self.B, self.S, self._stream()), 'bl_hex_transition')
        self.launches += 1
        if record:
            check(l.bl_pack_records(ptr(self.in_board), ptr(self.in_seats), ptr(self.out_terminal), ptr(self.out_actions), ptr(self.out_rewards),
                                    ptr(self.out_v), ptr(self.out_logits), ptr(self.ws.prior), ptr(self.out_record), self.B, self.A,
                                    self.record_width, self._stream()), 'bl_pack_records')
            self.launches += 1

    def _move_body(self, cparams, c_puct, noise_eps, alpha_scale, greedy, record):
        self._reset(c_puct)
        self.eval_root(cparams)
        self.set_root_prior(noise_eps, alpha_scale)
        for sim in range(1, self.T):
            self.descend_expand(sim)
            self.eval_leaves(cparams, sim)
            self.backup(sim)
        self.root_act(self.T, greedy)
        self.step_and_record(record)

    def play(self, board, seats, network, c_puct=1 / 16, noise_eps=.25, alpha_scale=10, greedy=False, record=True, use_graph=True):
        """One whole MOVE of every env in ONE CUDA graph: MCTS.__init__ + initialize with in-kernel Dirichlet noise,
        (n_nodes-1) x simulate, root, the agent's action (boardlaw/mcts/__init__.py:216-229), the env transition (boardlaw/main.py:177)
        and the packed trajectory record (boardlaw/main.py:179) — every launch from libboardlaw_b200.so; the host copies the worlds in
        and replays.  Returns views of static buffers, valid until the next call: (logits, prior, v, n_leaves, actions, new_board,
        new_seats, rewards, terminal, record)."""
        if board.shape[0] != self.B:
            raise ValueError(f'{board.shape[0]} envs do not fit a workspace of {self.B} (play() runs whole batches only)')
        cparams = network.packed()
        self.in_board.copy_(board)
        self.in_seats.copy_(seats)
        self.move += 1
        key = (cparams.W, cparams.D, cparams.precision, float(c_puct), network.token, network._pack_gen)
        self._graph_refs[key] = network._pack
        self._touch(key)
        gkey = ('move', float(noise_eps), float(alpha_scale), bool(greedy), bool(record)) + key
        if not use_graph:
            self._move_body(cparams, c_puct, noise_eps, alpha_scale, greedy, record)
        else:
            g = self._graphs.get(gkey)
            if g is None:
                saved = self._move_index()
                self._move_body(cparams, c_puct, noise_eps, alpha_scale, greedy, record)        # warm-up: kernel attributes, allocations
                g = torch.cuda.CUDAGraph()
                before = self.launches
                with torch.cuda.graph(g):
                    self._move_body(cparams, c_puct, noise_eps, alpha_scale, greedy, record)
                self._graph_launches[gkey] = self.launches - before
                self.launches = before
                self._graphs[gkey] = g
                self._restore_move_index(saved)
                self.errors.zero_()                          # (the warm-up pass played the same actions: nothing to keep)
            g.replay()
            self.launches += self._graph_launches[gkey]
        self.sim = self.T
        return (self.out_logits, self.ws.prior[:self.B], self.out_v, self.out_n_leaves, self.out_actions, self.nx_board, self.nx_seats,
                self.out_rewards, self.out_terminal, self.out_record)

    def children_dense(self):
        out = torch.empty((self.B, self.T, self.A), dtype=torch.int16, device=self.device)
        check(_lib.lib().bl_tree_children_dense(self._tp, ptr(out), self._stream()), 'bl_tree_children_dense')
        return out

    def leaf_worlds(self):
        """(board (B,S,S) u8, seats (B,) i32) of the current leaves — what the network is evaluated on."""
        idx = self.ws.leaf.long().clamp(min=0)
        envs = torch.arange(self.B, device=self.device)
        board = self.ws.board[envs, idx][:, :self.A].reshape(self.B, self.S, self.S).contiguous()
        return board, self.ws.seats[envs, idx].int()

    # ---- whole searches -------------------------------------------------------------------------------------
    def _simulate_all(self, cparams, first, last):
        for sim in range(first, last):
            self.descend_expand(sim)
            self.eval_leaves(cparams, sim)
            self.backup(sim)
        self.root(last)

    _STATIC = ('in_board', 'in_seats', 'root_logits', 'root_v', 'out_logits', 'out_v', 'out_n_leaves', 'out_actions', 'nx_board', 'nx_seats',
               'out_rewards', 'out_terminal', 'out_record', 'n_sims')

    def _search_partial(self, board, seats, network, **kwargs):
        """A search over n < capacity envs in the same workspace (arena-style callers, boardlaw/arena/common.py:86-96, hand
        the agent a different sub-batch every move): the arrays are env-major, so the first n envs' slices are a complete
        workspace and only ``bl_tree.B`` changes.  Runs eagerly — a graph per sub-batch size would be captured once and
        never replayed."""
        n, cap = board.shape[0], self.B
        self.scratch_for(network.packed())                        # sized at capacity, before B shrinks
        saved = {k: getattr(self, k) for k in self._STATIC}
        try:
            self.B = self.ctree.B = n
            for k, v in saved.items():
                setattr(self, k, v[:n])
            kwargs['use_graph'] = False
            return tuple(x.clone() for x in self.search(board, seats, network, **kwargs))
        finally:
            self.B = self.ctree.B = cap
            for k, v in saved.items():
                setattr(self, k, v)

    def search(self, board, seats, network, c_puct=1 / 16, noise_eps=.25, alpha_scale=10, noise=None, use_graph=True):
        """One whole search (MCTS.__init__ + initialize + (n_nodes-1) x simulate + root).  Returns
        (logits half (B,A), prior half (B,A), v half (B,2), n_leaves i64 (B,)) as views of static buffers (copies when the
        batch is smaller than the workspace)."""
        from .mcts import dirichlet_mix
        if board.shape[0] < self.B:
            return self._search_partial(board, seats, network, c_puct=c_puct, noise_eps=noise_eps, alpha_scale=alpha_scale, noise=noise)
        if board.shape[0] != self.B:
            raise ValueError(f'{board.shape[0]} envs do not fit a workspace of {self.B}')
        cparams = network.packed()
        self.in_board.copy_(board)
        self.in_seats.copy_(seats)
        self.move += 1                                 # (the device-side move counter, counters[6], is advanced by the reset kernel)
        # graphs bake in the addresses of the network's staged operands: key on the model's own token (CPython reuses id() values once a
        # model is collected) and on the staging generation, and keep the operands alive for as long as the graph is cached
        key = (cparams.W, cparams.D, cparams.precision, float(c_puct), network.token, network._pack_gen)
        self._graph_refs[key] = network._pack
        self._touch(key)
        if not use_graph:
            self._reset(c_puct)
            self.eval_root(cparams)
        else:
            g = self._graphs.get(('head',) + key)
            if g is None:
                saved = self._move_index()
                self._reset(c_puct); self.eval_root(cparams)            # warm-up: sets kernel attributes, allocs
                g = torch.cuda.CUDAGraph()
                before = self.launches
                with torch.cuda.graph(g):
                    self._reset(c_puct)
                    self.eval_root(cparams)
                self._graph_launches[('head',) + key] = self.launches - before
                self.launches = before                                   # captured, not launched
                self._graphs[('head',) + key] = g
                self._restore_move_index(saved)
            g.replay()
            self.launches += self._graph_launches[('head',) + key]
        valid = self.in_board.reshape(self.B, self.A) == 0
        valid = torch.where(self.in_seats[:, None] == 1,
                            (self.in_board.transpose(1, 2).reshape(self.B, self.A) == 0), valid)
        mixed = dirichlet_mix(self.root_logits, valid, noise_eps, alpha_scale, noise)
        self.set_eval(0, mixed, self.root_v)
        if not use_graph:
            self._simulate_all(cparams, 1, self.T)
        else:
            g = self._graphs.get(('sims',) + key)
            if g is None:
                self._simulate_all(cparams, 1, self.T)                       # warm-up on this very tree (discarded)
                # re-establish the state the graph expects, then capture
                saved = self._move_index()
                self._reset(c_puct); self.eval_root(cparams); self.set_eval(0, mixed, self.root_v)
                self._restore_move_index(saved)
                g = torch.cuda.CUDAGraph()
                before = self.launches
                with torch.cuda.graph(g):
                    self._simulate_all(cparams, 1, self.T)
                self._graph_launches[('sims',) + key] = self.launches - before
                self.launches = before
                self._graphs[('sims',) + key] = g
            g.replay()
            self.launches += self._graph_launches[('sims',) + key]
        self.sim = self.T
        return self.out_logits, self.ws.prior[:self.B], self.out_v, self.out_n_leaves
