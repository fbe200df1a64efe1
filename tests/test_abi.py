"""CPU: the C-ABI library builds, loads and exports exactly what include/boardlaw_b200.h declares; the Python host
refuses to run without it and without a CUDA device (no fallback path)."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / 'include' / 'boardlaw_b200.h'


def declared_symbols():
    text = re.sub(r'/\*.*?\*/', '', HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r'\b(bl_[a-z0-9_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def built():
    from boardlaw_b200 import build
    return build.build()


def test_header_symbols_exported(built):
    lib = ctypes.CDLL(str(built))
    syms = declared_symbols()
    assert len(syms) >= 12
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f'declared in the header but not exported: {missing}'


def test_python_binding_covers_header(built):
    from boardlaw_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert _lib.lib().bl_abi_version() == _lib.ABI_VERSION


def test_no_torch_types_in_signatures():
    code = re.sub(r'/\*.*?\*/', '', HEADER.read_text(), flags=re.S)
    assert 'at::' not in code and 'torch' not in code.lower() and 'Tensor' not in code
    assert '#include <stdint.h>' in code and 'extern "C"' in code


def test_exp_table_is_host_libm(built):
    import numpy as np
    import oracle
    from boardlaw_b200 import _lib
    a, b = _lib.exp_table_host(), oracle.exp_table()
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_cpu_tensors_are_rejected(built):
    """There is no CPU path: the wrappers raise instead of falling back."""
    from boardlaw_b200.hex import cuda as hc
    board = torch.zeros((2, 3, 3), dtype=torch.uint8)
    with pytest.raises(RuntimeError, match='no CPU path'):
        hc.step(board, torch.zeros(2, dtype=torch.int32), torch.zeros(2, dtype=torch.int32))
    with pytest.raises(TypeError):
        hc.step(board.int(), torch.zeros(2, dtype=torch.int32), torch.zeros(2, dtype=torch.int32))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from boardlaw_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', tmp_path / 'nope.so')
    with pytest.raises(RuntimeError, match='no CPU or PyTorch fallback'):
        _lib.lib()


def test_product_never_touches_the_oracle():
    """Nothing under boardlaw_b200/ imports, loads or executes oracle/ (test infrastructure), and the only shared library the product
    opens is its own."""
    import re
    root = Path(__file__).resolve().parents[1] / 'boardlaw_b200'
    for f in root.rglob('*.py'):
        src = f.read_text()
        assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), f'{f} imports oracle'
        assert 'oracle/' not in src and 'oracle.' not in src.replace('the oracle.', ''), f'{f} refers to oracle'
        for m in re.finditer(r'CDLL\(([^)]*)\)', src):
            assert 'LIB' in m.group(1) or 'libboardlaw_b200' in m.group(1), f'{f} loads {m.group(1)}'
