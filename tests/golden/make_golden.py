"""Generates the golden vectors under tests/golden/ by running the REFERENCE itself:
the Python layer imported from /root/reference on the reference's own CPU extension built from its
unmodified sources (oracle/_ref, -O2 variant; the -O0 variant is run beside it and must agree).

Run in the build container only:   python tests/golden/make_golden.py
The fixtures travel to the GPU box, where neither /root/reference nor this script's inputs exist.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests'))

from oracle import build_ref, pyref, refpy  # noqa: E402
from test_oracle_vs_ref import harvest  # noqa: E402

OUT = Path(__file__).resolve().parent


def npify(d):
    out = {}
    for k, v in d.items():
        v = v.detach()
        if v.dtype == torch.float16:
            out[k] = v.view(torch.int16).numpy().view(np.float16)
        else:
            out[k] = v.numpy()
    return out


def hex_playouts():
    """Random playouts through hexcuda.step / hexcuda.observe: every action, board, reward."""
    ref_hex = build_ref.load('hexcuda', 'O0')
    for S, B in [(3, 64), (5, 64), (9, 32), (11, 32), (13, 16)]:
        g = torch.Generator().manual_seed(100 + S)
        steps = 2 * S * S
        board = torch.zeros((B, S, S), dtype=torch.uint8)
        seats = torch.zeros((B,), dtype=torch.int32)
        rec = dict(actions=[], seats=[], boards=[], rewards=[], obs=[])
        for t in range(steps):
            obs = ref_hex.observe(board, seats)
            valid = (obs == 0).all(-1).reshape(B, -1)
            actions = torch.multinomial(valid.float(), 1, generator=g).squeeze(-1).int()
            rec['seats'].append(seats.clone()); rec['actions'].append(actions.clone())
            if t % 8 == 0:
                rec['obs'].append(obs.to(torch.uint8).clone())
            rewards = ref_hex.step(board, seats, actions)
            rec['boards'].append(board.clone()); rec['rewards'].append(rewards.clone())
            terminal = (rewards > 0).any(-1)
            board[terminal] = 0
            seats = 1 - seats
            seats[terminal] = 0
        np.savez_compressed(OUT / f'hex_playout_S{S}.npz', **{k: torch.stack(v).numpy() for k, v in rec.items()})
        print('hex', S, 'wins', int((torch.stack(rec['rewards']) > 0).sum()))


def mcts_calls():
    """Every descend / backup / root call of real searches: inputs and outputs."""
    for S, B, T, W, D, seed in [(5, 48, 16, 32, 2, 21), (9, 24, 64, 64, 4, 22)]:
        calls, decisions, sd = harvest('O2', S, B, T, W, D, seed)
        calls0, decisions0, _ = harvest('O0', S, B, T, W, D, seed)
        for (k, i, o), (k0, i0, o0) in zip(calls, calls0):
            assert k == k0 and all(torch.equal(o[x], o0[x]) for x in o), 'O0 and O2 reference builds disagree'
        keep = {}
        n_desc = 0
        for idx, (kind, ins, outs) in enumerate(calls):
            if kind == 'descend':
                n_desc += 1
                # full inputs for a subset of the sims, outputs for all
                if n_desc in (1, 2, T // 2, T - 2, T - 1):
                    for k, v in npify(ins).items():
                        keep[f'd{n_desc}_in_{k}'] = v
                for k, v in npify(outs).items():
                    keep[f'd{n_desc}_out_{k}'] = v
            elif kind == 'backup':
                if n_desc in (1, T // 2, T - 1):
                    for k, v in npify(ins).items():
                        keep[f'b{n_desc}_in_{k}'] = v
                    for k, v in npify(outs).items():
                        keep[f'b{n_desc}_out_{k}'] = v
            else:
                for k, v in npify(ins).items():
                    keep[f'r_in_{k}'] = v
                keep['r_out_probs'] = npify(outs)['probs']
        d = decisions[0]
        keep.update({f'agent_{k}': v for k, v in npify(dict(d)).items()})
        keep.update({f'sd_{k}': v.numpy() for k, v in sd.items()})
        np.savez_compressed(OUT / f'mcts_calls_S{S}_T{T}.npz', **keep)
        print('mcts', S, T, len(calls), 'calls')


def network():
    ref = refpy.load('O2')
    for S, W, D in [(5, 32, 2), (9, 256, 4), (11, 64, 3)]:
        sd = pyref.synth_state_dict(S, W, D, seed=S)
        w = ref.Hex.initial(64, S, device='cpu')
        g = torch.Generator().manual_seed(S)
        for _ in range(S * S // 2):
            w, _ = w.step(torch.multinomial(w.valid.float(), 1, generator=g).squeeze(-1))
        net = ref.FCModel(w.obs_space, w.action_space, width=W, depth=D)
        net.load_state_dict(sd)
        with torch.no_grad():
            r = net(w)
        np.savez_compressed(OUT / f'net_S{S}_W{W}_D{D}.npz', board=w.board.numpy(), seats=w.seats.numpy(),
                            logits=r.logits.numpy(), v=r.v.numpy(), **{f'sd_{k}': v.numpy() for k, v in sd.items()})
        print('net', S, W, D)


if __name__ == '__main__':
    assert refpy.present(), 'needs /root/reference'
    build_ref.build()
    hex_playouts()
    mcts_calls()
    network()
