"""ctypes binding of ``libboardlaw_b200.so`` (the C ABI declared in ``include/boardlaw_b200.h``).

This is the only place the shared library is opened.  There is no CPU path: if the library is
missing, or a tensor is not on a CUDA device, the wrappers raise.
"""
import ctypes
from ctypes import c_int, c_int64, c_uint64, c_void_p, c_float, POINTER, Structure
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / 'libboardlaw_b200.so'

ABI_VERSION = 9


class FCParams(Structure):
    """``bl_fc_params`` (include/boardlaw_b200.h)."""
    _fields_ = [('S', c_int), ('W', c_int), ('D', c_int), ('precision', c_int),
                ('w_in', c_void_p), ('b_in', c_void_p), ('w_res', c_void_p), ('b_res', c_void_p),
                ('alpha', c_void_p), ('w_pol', c_void_p), ('b_pol', c_void_p), ('w_val', c_void_p),
                ('b_val', c_void_p), ('packed', c_void_p), ('b_head', c_void_p), ('tc_nsplit', c_int)]


class Tree(Structure):
    """``bl_tree`` (include/boardlaw_b200.h)."""
    _fields_ = [('B', c_int), ('T', c_int), ('S', c_int), ('A', c_int), ('Sn', c_int), ('AP', c_int), ('BP', c_int),
                ('pi', c_void_p), ('cpi', c_void_p), ('psum', c_void_p), ('cprior', c_void_p), ('logits', c_void_p), ('board', c_void_p), ('node', c_void_p), ('aux', c_void_p), ('parent_of', c_void_p), ('kids', c_void_p),
                ('c_puct', c_void_p), ('leaf', c_void_p),
                ('leaf_parent', c_void_p), ('leaf_action', c_void_p), ('leaf_v', c_void_p), ('prior', c_void_p), ('qrange', c_void_p),
                ('counters', c_void_p), ('exp_lut', c_void_p), ('scratch', c_void_p), ('scratch_bytes', c_int64)]


P = c_void_p
# name -> (restype, argtypes).  Must list every symbol include/boardlaw_b200.h declares
# (tests/test_abi.py checks the two against each other).
SIGNATURES = {
    'bl_abi_version': (c_int, []),
    'bl_set_device': (c_int, [c_int]),
    'bl_exp_table_host': (None, [P]),
    'bl_hex_step': (c_int, [P, P, P, P, c_int, c_int, P]),
    'bl_hex_observe': (c_int, [P, P, P, c_int, c_int, P]),
    'bl_hex_transition': (c_int, [P, P, P, P, P, P, P, P, c_int, c_int, c_int, P]),
    'bl_hex_valid': (c_int, [P, P, P, c_int, c_int, P]),
    'bl_hex_random_transition': (c_int, [P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, P]),
    'bl_mcts_descend': (c_int, [P] * 13 + [c_int] * 4 + [P]),
    'bl_mcts_root': (c_int, [P] * 10 + [c_int] * 4 + [P]),
    'bl_mcts_backup': (c_int, [P] * 7 + [c_int] * 3 + [P]),
    'bl_mcts_transition_q': (c_int, [P] * 4 + [c_int] * 3 + [P]),
    'bl_fc_uses_tensor_cores': (c_int, [POINTER(FCParams)]),
    'bl_fc_scratch_bytes': (c_int64, [POINTER(FCParams), c_int]),
    'bl_fc_forward': (c_int, [POINTER(FCParams), P, P, P, P, P, c_int, P]),
    'bl_tree_scratch_bytes': (c_int64, [POINTER(Tree)]),
    'bl_tree_reset': (c_int, [POINTER(Tree), P, P, c_float, P]),
    'bl_tree_set_eval': (c_int, [POINTER(Tree), c_int, P, P, c_int, P]),
    'bl_tree_descend_expand': (c_int, [POINTER(Tree), c_int, P, c_uint64, P]),
    'bl_tree_backup': (c_int, [POINTER(Tree), c_int, P]),
    'bl_debug_set_descend_variant': (c_int, [c_int]),
    'bl_debug_set_phase_profile': (c_int, [P]),
    'bl_debug_set_descend_grid': (c_int, [c_int]),
    'bl_selftest_division': (c_int, [c_uint64, c_int, c_int, P, P]),
    'bl_tree_eval_scratch_bytes': (c_int64, [POINTER(Tree), POINTER(FCParams)]),
    'bl_tree_eval_leaves': (c_int, [POINTER(Tree), POINTER(FCParams), c_int, P, P]),
    'bl_tree_eval_root': (c_int, [POINTER(Tree), POINTER(FCParams), P, P, P, P]),
    'bl_tree_root': (c_int, [POINTER(Tree), c_int, P, P, P, P, P]),
    'bl_tree_root_act': (c_int, [POINTER(Tree), c_int, P, P, P, P, P, P, c_int, c_uint64, P]),
    'bl_tree_set_root_prior': (c_int, [POINTER(Tree), P, P, P, P, P, c_float, c_float, c_uint64, P]),
    'bl_pack_records': (c_int, [P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, P]),
    'bl_tree_children_dense': (c_int, [POINTER(Tree), P, P]),
    'bl_reward_to_go': (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_float, P]),
    'bl_policy_value_loss': (c_int, [P, P, P, P, P, P, P, P, c_int, c_int, P]),
    'bl_adam_step': (c_int, [P, P, P, P, c_int64, c_float, c_float, c_float, c_float, c_int, P]),
    'bl_gemm_f32_workspace_bytes': (c_int64, [c_int, c_int, c_int]),
    'bl_gemm_f32': (c_int, [P, c_int64, c_int64, c_int, P, P, c_int64, c_int64, c_int, P, P, P, c_int64, c_int, c_int, c_int, P, c_int64, P]),
}

_lib = None


def lib():
    """Opens the library (once).  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f'{LIB_PATH} is missing: build it with `python -m boardlaw_b200.build` '
                '(boardlaw_b200 has no CPU or PyTorch fallback path)')
        l = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            if not hasattr(l, name):
                continue   # engine symbols may be absent in partial builds; test_abi checks completeness
            f = getattr(l, name)
            f.restype, f.argtypes = res, args
        v = l.bl_abi_version()
        if v != ABI_VERSION:
            raise RuntimeError(f'libboardlaw_b200.so has ABI version {v}, the Python host expects {ABI_VERSION}')
        _lib = l
    return _lib


def check(err, what):
    if err == 0:
        return
    if err < 0:
        raise ValueError(f'{what}: invalid argument (code {err})')
    raise RuntimeError(f'{what}: CUDA error {err}')


def require_cuda(*tensors):
    dev = None
    for t in tensors:
        if t is None:
            continue
        if t.device.type != 'cuda':
            raise RuntimeError('boardlaw_b200 kernels run on CUDA tensors only (there is no CPU path); '
                               f'got a tensor on {t.device}')
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError('Inputs span multiple devices')
    return dev


def proxy(t, dtype, ndim, name):
    """The checks of the reference's TensorProxy (boardlaw/cpp/common.h:34-38): dtype mismatch is a
    TypeError, wrong rank / non-contiguous a RuntimeError."""
    if t.dtype != dtype:
        raise TypeError(f'{name}: expected {dtype} got {t.dtype}')
    if ndim is not None and t.ndim != ndim:
        raise RuntimeError(f'{name}: expected {ndim} dims got {t.ndim}')
    if not t.is_contiguous():
        raise RuntimeError(f'{name} must be contiguous')
    return t


def ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def stream_for(device):
    """Activates ``device`` in the library's CUDA runtime and returns torch's current stream on it."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    check(lib().bl_set_device(idx), 'bl_set_device')
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


_exp_host = None
_exp_dev = {}


def exp_table_host():
    global _exp_host
    if _exp_host is None:
        out = np.empty(65536, dtype=np.float32)
        lib().bl_exp_table_host(out.ctypes.data_as(c_void_p))
        _exp_host = out
    return _exp_host


def exp_lut(device):
    """Device copy of expf over every binary16 pattern (host libm values), one per device."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    if key not in _exp_dev:
        _exp_dev[key] = torch.from_numpy(exp_table_host()).to(device)
    return _exp_dev[key]


_log_dev = {}


def log_lut(device):
    """Device copy of ``h.float().log().half()`` over every binary16 pattern, evaluated by torch on the host exactly
    as ``MCTS.root`` does on the reference's CPU path (boardlaw/mcts/__init__.py:147)."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    if key not in _log_dev:
        every = torch.arange(65536, dtype=torch.int32).to(torch.int16).view(torch.float16)
        _log_dev[key] = every.float().log().half().to(device)
    return _log_dev[key]
