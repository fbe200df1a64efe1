"""Builds ``boardlaw_b200/libboardlaw_b200.so`` in-tree with nvcc for sm_100a.

No torch headers are involved: the library is a plain C-ABI shared object (``include/boardlaw_b200.h``).
nvcc cross-compiles without a GPU, so this runs in the build container; the resulting ``.so`` is
git-ignored but travels to the GPU box with the snapshot.
"""
import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / 'csrc'
OBJ = HERE / 'csrc' / '_obj'
LIB = HERE / 'libboardlaw_b200.so'

ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
COMMON = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC',
          '--expt-relaxed-constexpr']

# per-file extra flags.  mcts.cu holds the bit-exact fp32 arithmetic (no FMA contraction, IEEE
# division/sqrt, denormals kept) — see DESIGN.md "arithmetic contract".
SOURCES = {
    'hex.cu': [],
    'mcts.cu': ['-fmad=false', '-prec-div=true', '-prec-sqrt=true', '-ftz=false'],
    'engine.cu': ['-fmad=false', '-prec-div=true', '-prec-sqrt=true', '-ftz=false'],
    'descend.cu': ['-fmad=false', '-prec-div=true', '-prec-sqrt=true', '-ftz=false'] + (['-DBL_CHILD_ILP=' + os.environ['BL_CHILD_ILP']] if 'BL_CHILD_ILP' in os.environ else []),
    'descend_mw.cu': ['-fmad=false', '-prec-div=true', '-prec-sqrt=true', '-ftz=false'],
    'descend_fx.cu': ['-fmad=false', '-prec-div=true', '-prec-sqrt=true', '-ftz=false'],
    'experimental.cu': [],
    'net.cu': [],
    'net_tc.cu': [],
    'net_tc_wide.cu': [],
    'learner.cu': [],
    'gemm_tc.cu': [],
    'host.cu': [],
}


# Measured-and-rejected descent variants (4 = producer/consumer warps, 6 = speculative evaluation of every node, 7 = packed pass lanes + warp-cooperative visits):
# bit-exact but slower than the default, so they stay out of the product library unless BL_EXPERIMENTAL=1 is set at build time
# (tests/test_gpu_mcts.py and tests/test_gpu_fx.py skip them when they are not compiled in).
EXPERIMENTAL = {
    'descend_all.cu': ['-fmad=false', '-prec-div=true', '-prec-sqrt=true', '-ftz=false'],
    'descend_pc.cu': ['-fmad=false', '-prec-div=true', '-prec-sqrt=true', '-ftz=false'],
    'descend_pk.cu': ['-fmad=false', '-prec-div=true', '-prec-sqrt=true', '-ftz=false'],
}
if os.environ.get('BL_EXPERIMENTAL') == '1':
    del SOURCES['experimental.cu']
    SOURCES.update(EXPERIMENTAL)


def nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not Path(exe).exists():
        raise RuntimeError('nvcc not found: the boardlaw_b200 kernels cannot be built')
    return exe


def _stale(target, deps):
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force=False, verbose=False):
    OBJ.mkdir(exist_ok=True)
    headers = list(CSRC.glob('*.cuh')) + [HERE.parent / 'include' / 'boardlaw_b200.h', Path(__file__)]
    objs, procs = [], []
    for stale in list(EXPERIMENTAL) + ['experimental.cu']:                 # objects of the other configuration must not be linked
        if stale not in SOURCES and (OBJ / (Path(stale).stem + '.o')).exists():
            (OBJ / (Path(stale).stem + '.o')).unlink()
            force = True
    for name, extra in SOURCES.items():
        src = CSRC / name
        if not src.exists():
            continue
        obj = OBJ / (src.stem + '.o')
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc()] + ARCH + COMMON + extra + ['-Xptxas', '-v', '-c', str(src), '-o', str(obj)]
            if verbose:
                print(' '.join(cmd))
            procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for name, p in procs:
        out, _ = p.communicate()
        (OBJ / (Path(name).stem + '.ptxas.log')).write_text(out)
        if p.returncode != 0:
            failed = True
            sys.stderr.write(out)
        elif verbose:
            print(out)
    if failed:
        raise RuntimeError('nvcc failed')
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc()] + ARCH + ['-shared', '-o', str(LIB)] + [str(o) for o in objs] + ['-cudart', 'static']
        if verbose:
            print(' '.join(cmd))
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == '__main__':
    print(build(force='-f' in sys.argv, verbose='-v' in sys.argv))
