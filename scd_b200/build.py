"""In-tree build of the CUDA library (``scd_b200/libscd_b200.so``) for sm_100a.

``nvcc`` cross-compiles without a GPU.  The ``.so`` is git-ignored but travels to the GPU box with
the repo snapshot.  Run ``python -m scd_b200.build`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
LIB = os.path.join(PKG, 'libscd_b200.so')
SOURCES = [os.path.join(CSRC, 'scd_api.cu'), os.path.join(CSRC, 'hungarian.cpp'), os.path.join(CSRC, 'constrained.cpp')]
HEADERS = [os.path.join(CSRC, f) for f in ('ptx.cuh', 'peer_kernel.cuh', 'naming_kernel.cuh', 'kmeans_kernel.cuh', 'estep_tc_kernel.cuh', 'vote_kernel.cuh', 'eval_kernel.cuh')] + \
          [os.path.join(os.path.dirname(PKG), 'include', 'scd_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden']


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found (set NVCC=...)')


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB] + SOURCES
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + proc.stdout + proc.stderr)
    if verbose:
        sys.stderr.write(proc.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
