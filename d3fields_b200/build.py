"""Build libd3f.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m d3fields_b200.build [--force]

The library lands in d3fields_b200/_lib/libd3f.so: git-ignored, but it travels to the GPU box
with the working tree.  It links the static CUDA runtime only — no torch, no libcuda at link time.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
LIB_DIR = os.path.join(PKG, '_lib')
LIB_PATH = os.path.join(LIB_DIR, 'libd3f.so')
STAMP = os.path.join(LIB_DIR, 'libd3f.srchash')
INCLUDE = os.path.join(os.path.dirname(PKG), 'include')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xptxas=-v', '-shared', '-Xcompiler', '-fPIC']


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found (set NVCC=...)')


def _sources():
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh', '.h')))
    files.append(os.path.join(INCLUDE, 'd3f.h'))
    return files


def source_hash() -> str:
    h = hashlib.sha256()
    for f in _sources():
        h.update(os.path.basename(f).encode())
        with open(f, 'rb') as fh:
            h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    try:
        with open(STAMP) as fh:
            return os.path.exists(LIB_PATH) and fh.read().strip() == source_hash()
    except OSError:
        return False


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/d3f_abi.cu (which includes every kernel header) into libd3f.so."""
    if not force and is_current():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    import fcntl
    with open(os.path.join(LIB_DIR, '.build.lock'), 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)          # several ranks may import the package at once
        try:
            if not force and is_current():        # another process built it while we waited
                return LIB_PATH
            tmp = LIB_PATH + f'.tmp{os.getpid()}'
            cmd = [_nvcc()] + NVCC_FLAGS + ['-I', INCLUDE, '-o', tmp, os.path.join(CSRC, 'd3f_abi.cu')]
            proc = subprocess.run(cmd, capture_output=True, text=True)
            log = proc.stdout + proc.stderr
            with open(os.path.join(LIB_DIR, 'build.log'), 'w') as fh:
                fh.write(' '.join(cmd) + '\n' + log)
            if proc.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError('nvcc failed:\n' + log)
            os.replace(tmp, LIB_PATH)             # atomic: a reader never sees a half-written library
            if verbose:
                print(log)
            with open(STAMP, 'w') as fh:
                fh.write(source_hash())
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == '__main__':
    path = build(force='--force' in sys.argv, verbose=True)
    print(path)
