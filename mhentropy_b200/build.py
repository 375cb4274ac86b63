"""Build the C-ABI CUDA library in-tree: ``mhentropy_b200/libmhentropy_b200.so`` (sm_100a only).

``python -m mhentropy_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles without a
GPU.  Objects go to ``build/`` (git-ignored); the ``.so`` stays next to the package so it travels
with the tree.
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
OBJ = os.path.join(ROOT, 'build', 'obj')
LIB = os.path.join(PKG, 'libmhentropy_b200.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
    '-Xcompiler', '-fPIC', '-DMHE_SM=100', '--expt-relaxed-constexpr',
] + os.environ.get('MHE_NVCC_EXTRA', '').split()


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found (needed to build libmhentropy_b200.so)')


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps_mtime() -> float:
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, 'include', 'mhentropy_b200.h'), __file__]
    return max(os.path.getmtime(p) for p in paths)


def needs_build() -> bool:
    return not os.path.exists(LIB) or os.path.getmtime(LIB) < _deps_mtime()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdr_mtime = _deps_mtime()

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= hdr_mtime:
            return obj
        cmd = [nvcc, *NVCC_FLAGS, '-c', src, '-o', obj]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(compile_one, sources()))
    tmp = LIB + '.tmp'
    r = subprocess.run([nvcc, '-shared', '-o', tmp, *objs], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    os.replace(tmp, LIB)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
