"""Observation/action space descriptors, as ``boardlaw/heads.py:7-11``.

Only the space types the Hex path uses are kept; the network heads themselves (TensorIntake,
MaskedOutput, ValueOutput — boardlaw/heads.py:41-52,93-104,128-142) are fused into the forward kernels
of ``boardlaw_b200.networks``.
"""
from collections import namedtuple

Empty = namedtuple('Empty', ())
Discrete = namedtuple('Discrete', ('dim',))
Masked = namedtuple('Masked', ('dim',))
Vector = namedtuple('Vector', ('dim',))
Tensor = namedtuple('Tensor', ('dim',))
