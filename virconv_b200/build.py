"""Build libvirconv_sm100.so (the C-ABI shared library of include/virconv_b200.h) in-tree with nvcc.

nvcc cross-compiles sm_100a without a GPU; the .so is git-ignored but travels with the gpurun snapshot.
    python -m virconv_b200.build [--verbose] [--force]
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libvirconv_sm100.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr',
         '-DVC_BUILDING']


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, 'rb') as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


def build(verbose: bool = False, force: bool = False, defines=(), suffix: str = '') -> str:
    """`defines` / `suffix`: experiment variants (e.g. -DVC_TC_STAGES=6 -> libvirconv_sm100_s6.so, picked up through
    the VIRCONV_LIB environment variable); the default build has neither."""
    os.makedirs(LIBDIR, exist_ok=True)
    if suffix:
        return _build_variant(list(defines), suffix, verbose)
    srcs = sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))] + [
        os.path.join(os.path.dirname(HERE), 'include', 'virconv_b200.h')]
    stamp = os.path.join(LIBDIR, '.build_digest')
    dig = _digest(deps)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    objs = []

    def compile_one(src):
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + '.o')
        cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    with open(stamp, 'w') as f:
        f.write(dig)
    return LIB


def _build_variant(defines, suffix, verbose):
    lib = os.path.join(LIBDIR, f'libvirconv_sm100_{suffix}.so')
    objs = []
    for src in sources():
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + f'_{suffix}.o')
        r = subprocess.run([NVCC] + FLAGS + [f'-D{d}' for d in defines] + ['-c', src, '-o', obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s' % (src, r.stderr))
        objs.append(obj)
    r = subprocess.run([NVCC, '-shared', '-o', lib] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart'],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stderr)
    return lib


if __name__ == '__main__':
    print(build(verbose='--verbose' in sys.argv, force='--force' in sys.argv))
