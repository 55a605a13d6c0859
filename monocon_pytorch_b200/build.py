"""In-tree nvcc build of the sm_100a engine -> monocon_pytorch_b200/libmonocon_b200.so.

The shared library carries hand-written CUDA only (no torch, no cuDNN/cuBLAS); cudart is linked
statically so that the .so loads without a GPU (symbol checks) and next to torch's own runtime.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libmonocon_b200.so')
OBJ = os.path.join(HERE, 'build')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
         '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr']


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps():
    inc = os.path.join(os.path.dirname(HERE), 'include')
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    out += [os.path.join(inc, f) for f in os.listdir(inc)]
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    if not os.path.exists(NVCC):
        raise RuntimeError(f'nvcc not found at {NVCC}; cannot build {LIB}')
    os.makedirs(OBJ, exist_ok=True)
    hdr_time = max(os.path.getmtime(p) for p in _deps() if not p.endswith('.cu'))

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_time):
            return obj
        cmd = [NVCC, *FLAGS, '-c', src, '-o', obj]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [NVCC, '-shared', '-o', LIB, *objs, '-gencode', 'arch=compute_100a,code=sm_100a', '-cudart', 'static']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
