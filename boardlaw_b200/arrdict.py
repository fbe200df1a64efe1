"""Dict-of-tensors records with attribute access and leaf-wise indexing.

The reference builds ``Hex`` and every decision/transition record on ``rebar.arrdict``
(rebar/arrdict.py:19-86, rebar/dotdict.py:7-30).  When ``rebar`` is importable (i.e. when this package is
dropped into a boardlaw checkout) that implementation is re-exported unchanged so records interoperate;
otherwise the small stand-alone implementation below provides the behaviours the hot path relies on:

  * ``d.key`` / ``d['key']`` access; ``d[idx]`` indexes every leaf; ``d[idx] = other`` assigns leaf-wise
  * ``d.method(...)`` / ``d.attr`` broadcast to the leaves (``.clone()``, ``.to(...)``, ``.half()``, ``.shape`` ...)
  * ``namedarrtuple(name, fields)`` fixed-field subclasses; ``stack`` / ``cat`` over sequences of records
"""
try:  # pragma: no cover - exercised only inside a boardlaw checkout
    from rebar.arrdict import arrdict, namedarrtuple, stack, cat, torchify, numpyify  # noqa: F401
    REBAR = True
except Exception:  # ModuleNotFoundError, or rebar's own optional imports failing
    REBAR = False

if not REBAR:
    from collections import OrderedDict

    import numpy as np
    import torch

    def _is_field(x):
        return isinstance(x, str) or (isinstance(x, tuple) and len(x) > 0 and all(isinstance(s, str) for s in x))

    class arrdict(OrderedDict):

        def __getattr__(self, key):
            if key.startswith('__') and key.endswith('__'):
                raise AttributeError(key)
            if key in self:
                return OrderedDict.__getitem__(self, key)
            try:
                got = [(k, getattr(v, key)) for k, v in self.items()]
            except AttributeError:
                raise AttributeError(f"There is no member called '{key}' and one of the leaves has no attribute '{key}'") from None
            return type(self)(got)

        def __call__(self, *args, **kwargs):
            return type(self)([(k, v(*args, **kwargs)) for k, v in self.items()])

        def __getitem__(self, x):
            if isinstance(x, str):
                return OrderedDict.__getitem__(self, x)
            return type(self)([(k, v[x]) for k, v in self.items()])

        def __setitem__(self, x, y):
            if _is_field(x):
                OrderedDict.__setitem__(self, x, y)
            elif isinstance(y, dict):
                for k in self:
                    OrderedDict.__getitem__(self, k)[x] = y[k]
            else:
                raise ValueError('Setting items must be done with a string key or by passing an arrdict')

        def __dir__(self):
            return sorted(set(list(super().__dir__()) + [k for k in self.keys() if isinstance(k, str)]))

        def __getstate__(self):
            return dict(self)

        def __setstate__(self, state):
            for k, v in state.items():
                OrderedDict.__setitem__(self, k, v)

        def __reduce__(self):
            return (_rebuild, (type(self), list(self.items())))

        def copy(self):
            return type(self)(list(self.items()))

        def map(self, f, *args, **kwargs):
            return type(self)([(k, v.map(f, *args, **kwargs) if isinstance(v, arrdict) else f(v, *args, **kwargs))
                               for k, v in self.items()])

        def __repr__(self):
            rows = [f'{type(self).__name__}:']
            for k, v in self.items():
                if hasattr(v, 'shape') and hasattr(v, 'dtype'):
                    rows.append(f'{k:<12s}{type(v).__name__}({tuple(v.shape)}, {v.dtype})')
                else:
                    rows.append(f'{k:<12s}{str(v).splitlines()[0] if str(v) else ""}')
            return '\n'.join(rows)

        __str__ = __repr__

    def _rebuild(cls, items):
        obj = cls.__new__(cls)
        OrderedDict.__init__(obj)
        for k, v in items:
            OrderedDict.__setitem__(obj, k, v)
        if hasattr(obj, '_restore'):
            obj._restore()
        return obj

    def _binary(name):
        def op(self, rhs):
            if isinstance(rhs, dict):
                return type(self)([(k, getattr(v, name)(rhs[k])) for k, v in self.items()])
            return type(self)([(k, getattr(v, name)(rhs)) for k, v in self.items()])
        op.__name__ = name
        return op

    for _n in ['lt', 'le', 'eq', 'ne', 'ge', 'gt', 'add', 'sub', 'mul', 'truediv', 'floordiv', 'mod', 'pow',
               'and', 'or', 'xor', 'radd', 'rsub', 'rmul', 'rtruediv']:
        setattr(arrdict, f'__{_n}__', _binary(f'__{_n}__'))
    arrdict.__hash__ = None

    def namedarrtuple(name='AnonymousNamedArrTuple', fields=()):
        fields = tuple(fields)

        def __init__(self, *args, **kwargs):
            OrderedDict.__init__(self)
            for k, v in OrderedDict(*args, **kwargs).items():
                OrderedDict.__setitem__(self, k, v)
            if set(fields) != set(self):
                raise KeyError(f'This NamedArrTuple subclass must be created with exactly the fields {fields}')

        def __setitem__(self, x, y):
            if _is_field(x) and (x not in fields):
                raise KeyError(f'Key "{x}" is not in this immutable NamedArrTuple, and so cannot be added')
            arrdict.__setitem__(self, x, y)

        def __delitem__(self, x):
            raise KeyError('Cannot delete keys from this immutable NameArrTuple subclass')

        return type(name, (arrdict,), {'__init__': __init__, '__setitem__': __setitem__, '__delitem__': __delitem__})

    def stack(xs, *args, **kwargs):
        x0 = xs[0]
        if isinstance(x0, dict):
            return type(x0)([(k, stack([y[k] for y in xs], *args, **kwargs)) for k in x0.keys()])
        if isinstance(x0, torch.Tensor):
            return torch.stack(list(xs), *args, **kwargs)
        if isinstance(x0, np.ndarray):
            return np.stack(xs, *args, **kwargs)
        if np.isscalar(x0):
            return np.array(xs, *args, **kwargs)
        raise ValueError(f"Can't stack {type(x0)}")

    def cat(xs, *args, **kwargs):
        x0 = xs[0]
        if isinstance(x0, dict):
            return type(x0)([(k, cat([y[k] for y in xs], *args, **kwargs)) for k in x0.keys()])
        if isinstance(x0, torch.Tensor):
            return torch.cat(list(xs), *args, **kwargs)
        if isinstance(x0, np.ndarray):
            return np.concatenate(xs, *args, **kwargs)
        if np.isscalar(x0):
            return np.array(xs)
        raise ValueError(f"Can't cat {type(x0)}")

    def torchify(a):
        if isinstance(a, dict):
            return type(a)([(k, torchify(v)) for k, v in a.items()])
        a = np.asarray(a)
        if np.issubdtype(a.dtype, np.floating):
            return torch.as_tensor(np.array(a), dtype=torch.float)
        if np.issubdtype(a.dtype, np.integer):
            return torch.as_tensor(np.array(a), dtype=torch.int)
        if np.issubdtype(a.dtype, np.bool_):
            return torch.as_tensor(np.array(a), dtype=torch.bool)
        raise ValueError(f"Can't handle {a.dtype}")

    def numpyify(t):
        if isinstance(t, dict):
            return type(t)([(k, numpyify(v)) for k, v in t.items()])
        if isinstance(t, torch.Tensor):
            return t.detach().cpu().numpy().copy()
        return t
