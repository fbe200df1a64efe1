"""TEST INFRASTRUCTURE — builds the reference's own CPU kernels into ``oracle/_ref/``.

The recipe mirrors the reference's loader for a machine without CUDA
(``boardlaw/cuda.py:29-45`` ``load_cpu``: ``wrappers.cpp`` only, ``-std=c++17 -DNOCUDA``,
no ``-O`` flag, ``with_cuda=False``).  Sources are compiled *where they lie* under
``/root/reference``; nothing from the reference is copied into this repository, and only
the resulting shared objects land in ``oracle/_ref/`` (git-ignored, but shipped to the GPU
box with the snapshot so the checker works there too).

Two variants are built per module:
  * ``O0`` — the reference exactly as its own loader builds it (glibc ``powf``/``expf``
    calls, no optimisation).  This is "the reference" for the CPU baseline.
  * ``O2`` — the same unmodified sources with ``-O2`` (gcc folds ``powf(x, 2)`` to ``x*x``);
    this is the arithmetic the CUDA path is bit-exact against.

Only ``tests/``, ``__graft_entry__`` and ``bench.py``'s CPU-baseline legs may use the result.
"""
import os
import sys
import importlib.util
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_OUT = HERE / '_ref'
REFERENCE = Path(os.environ.get('BOARDLAW_REFERENCE', '/root/reference'))

MODULES = {
    'hexcuda': 'boardlaw/hex/cpp/wrappers.cpp',
    'mctscuda': 'boardlaw/mcts/cpp/wrappers.cpp',
}
VARIANTS = {'O0': [], 'O2': ['-O2']}


def modname(name, variant):
    # the pybind init symbol is derived from the extension name, so the two variants need
    # distinct names to be loadable side by side
    return name if variant == 'O0' else f'{name}_{variant.lower()}'


def so_path(name, variant):
    return REF_OUT / variant / name / f'{modname(name, variant)}.so'


def build(verbose=False):
    """Compile every (module, variant) that is missing.  Needs ``/root/reference``."""
    if not REFERENCE.exists():
        return False
    import sysconfig
    import torch.utils.cpp_extension as ext
    [torch_libdir] = ext.library_paths()
    python_libdir = sysconfig.get_config_var('LIBDIR')
    libpython_ver = sysconfig.get_config_var('LDVERSION')
    for variant, oflags in VARIANTS.items():
        for name, rel in MODULES.items():
            if so_path(name, variant).exists():
                continue
            bdir = REF_OUT / variant / name
            bdir.mkdir(parents=True, exist_ok=True)
            ext.load(
                name=modname(name, variant),
                sources=[str(REFERENCE / rel)],
                extra_cflags=['-std=c++17', '-DNOCUDA'] + oflags,
                with_cuda=False,
                extra_ldflags=[
                    f'-lpython{libpython_ver}', '-ltorch', '-ltorch_python',
                    f'-L{torch_libdir}', f'-Wl,-rpath,{torch_libdir}',
                    f'-L{python_libdir}', f'-Wl,-rpath,{python_libdir}'],
                build_directory=str(bdir),
                verbose=verbose,
                is_python_module=False)
    return True


def build_cuda(verbose=False):
    """Variant ``CUDA``: the reference's CUDA kernels (``wrappers.cpp`` + ``cuda.cu`` of each module) for sm_100a, by the recipe of
    its own loader (``boardlaw/cuda.py:10-27`` ``load_cuda``: ``--use_fast_math -lineinfo``) with one flag changed, ``-std=c++14``
    -> ``-std=c++17`` for the device compile, because the ATen headers of the installed torch refuse C++14.  nvcc
    cross-compiles here without a GPU; ~5 minutes.  Only ``bench.py``'s ``reference_cuda`` leg loads the result."""
    if not REFERENCE.exists():
        return False
    import sysconfig
    import torch.utils.cpp_extension as ext
    [torch_libdir] = ext.library_paths()
    python_libdir = sysconfig.get_config_var('LIBDIR')
    libpython_ver = sysconfig.get_config_var('LDVERSION')
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0a')
    for name, rel in MODULES.items():
        if so_path(name, 'CUDA').exists():
            continue
        bdir = REF_OUT / 'CUDA' / name
        bdir.mkdir(parents=True, exist_ok=True)
        src = (REFERENCE / rel).parent
        ext.load(
            name=modname(name, 'CUDA'),
            sources=[str(src / 'wrappers.cpp'), str(src / 'cuda.cu')],
            extra_cflags=['-std=c++17'],
            extra_cuda_cflags=['--use_fast_math', '-lineinfo', '-std=c++17'],
            extra_include_paths=['/usr/local/cuda/include'],
            with_cuda=True,
            extra_ldflags=[
                f'-lpython{libpython_ver}', '-ltorch', '-ltorch_python', '-lc10_cuda', '-lc10',
                f'-L{torch_libdir}', f'-Wl,-rpath,{torch_libdir}',
                f'-L{python_libdir}', f'-Wl,-rpath,{python_libdir}'],
            build_directory=str(bdir),
            verbose=verbose,
            is_python_module=False)
    return True


_loaded = {}


def load(name, variant='O0'):
    """Import a prebuilt reference extension (a pybind11 module) from ``oracle/_ref``."""
    key = (name, variant)
    if key not in _loaded:
        import torch  # noqa: F401  (libtorch must be loaded before the extension)
        path = so_path(name, variant)
        if not path.exists():
            raise FileNotFoundError(f'{path} missing: run `python oracle/build_ref.py` where /root/reference exists')
        # Two variants of the same pybind module name cannot coexist under one sys.modules key;
        # the classes are py::module_local so separate specs are fine.
        spec = importlib.util.spec_from_file_location(modname(name, variant), str(path))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _loaded[key] = mod
    return _loaded[key]


def available(variant='O0'):
    return all(so_path(n, variant).exists() for n in MODULES)


if __name__ == '__main__':
    ok = build(verbose='-v' in sys.argv)
    if '--cuda' in sys.argv:
        build_cuda(verbose='-v' in sys.argv)
    print('built' if ok else 'reference sources not present; nothing built')
    for v in VARIANTS:
        for n in MODULES:
            print(v, n, so_path(n, v).exists())
