"""Builds the in-tree CUDA library `lib/libmfb200.so` for sm_100a with nvcc.

    python -m recommendation_gans_b200.build [--force]

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels with the repo
snapshot to the GPU box.
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libmfb200.so')
NVCC_FLAGS = ['-shared', '-Xcompiler', '-fPIC', '-gencode', 'arch=compute_100a,code=sm_100a',
              '-lineinfo', '-O3', '-std=c++17']


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: the mfb200 CUDA library cannot be built')


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + \
        [os.path.join(HERE, '..', 'include', 'mfb200.h')]
    return any(os.path.getmtime(p) > built for p in deps)


def build_library(force=False, verbose=False):
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB_PATH] + sources()
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError('nvcc failed:\n%s\n%s' % (' '.join(cmd), proc.stderr[-4000:]))
    if verbose:
        sys.stderr.write(proc.stderr)
    return LIB_PATH


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose='-v' in sys.argv))
