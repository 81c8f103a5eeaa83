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


def _dependencies():
    return sources() + sorted(glob.glob(os.path.join(CSRC, '*.cuh'))) + [os.path.join(HERE, '..', 'include', 'mfb200.h')]


def source_hash():
    """Content hash of everything the library is built from.  Staleness is decided by content, not by mtime: the
    snapshot that carries the prebuilt .so to a GPU box does not preserve a meaningful mtime order, and a spurious
    rebuild there would have N ranks rewriting the library under each other."""
    import hashlib
    h = hashlib.sha256()
    h.update(' '.join(NVCC_FLAGS).encode())
    for path in _dependencies():
        h.update(os.path.basename(path).encode())
        with open(path, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


HASH_PATH = LIB_PATH + '.srchash'


def is_stale():
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as f:
        return f.read().strip() != source_hash()


def _compile_one(args):
    src, obj, defines, verbose = args
    cmd = [_nvcc(), '-c', '-Xcompiler', '-fPIC'] + NVCC_FLAGS[3:] + ['-D' + d for d in defines] + \
        (['-Xptxas', '-v'] if verbose else []) + ['-o', obj, src]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError('nvcc failed:\n%s\n%s' % (' '.join(cmd), proc.stderr[-4000:]))
    return proc.stderr


def build_library(force=False, verbose=False, out=None, defines=()):
    """Compile csrc/*.cu into lib/libmfb200.so (one object per source, compiled in parallel, only the stale ones).
    `out` / `defines` build an experiment variant elsewhere (e.g. defines=['MFB_TC_EPI_WARPS=8']); load it with
    $MFB_LIB_PATH."""
    if out is None and not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    import fcntl
    with open(os.path.join(LIB_DIR, '.build.lock'), 'w') as lock:      # one builder at a time (ranks of one box)
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if out is None and not force and not is_stale():           # another process built it while we waited
                return LIB_PATH
            return _build_locked(force, verbose, out, defines)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force, verbose, out, defines):
    target = out or LIB_PATH
    obj_dir = os.path.join(LIB_DIR, 'obj' if out is None else 'obj_' + os.path.basename(target))
    os.makedirs(obj_dir, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, '*.cuh')) + [os.path.join(HERE, '..', 'include', 'mfb200.h')]
    hdr_time = max(os.path.getmtime(h) for h in headers)
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        fresh = os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_time)
        if force or defines or verbose or not fresh:
            jobs.append((src, obj, list(defines), verbose))
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
        logs = list(pool.map(_compile_one, jobs))
    tmp = target + '.tmp.%d' % os.getpid()
    cmd = [_nvcc()] + NVCC_FLAGS[:5] + ['-o', tmp] + objs
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError('link failed:\n%s\n%s' % (' '.join(cmd), proc.stderr[-4000:]))
    os.replace(tmp, target)                                            # readers never see a half-written library
    if out is None:
        with open(HASH_PATH + '.tmp', 'w') as f:
            f.write(source_hash())
        os.replace(HASH_PATH + '.tmp', HASH_PATH)
    if verbose:
        sys.stderr.write(''.join(logs))
    return target


if __name__ == '__main__':
    _out = sys.argv[sys.argv.index('--out') + 1] if '--out' in sys.argv else None
    _defs = [a[2:] for a in sys.argv if a.startswith('-D')]
    print(build_library(force='--force' in sys.argv, verbose='-v' in sys.argv, out=_out, defines=_defs))
