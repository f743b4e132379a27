"""Build libtegb200.so (the C-ABI runtime + static kernels) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_PATH = os.path.join(HERE, 'libtegb200.so')
RUNTIME_HEADER = os.path.join(HERE, 'include', 'teg_b200_runtime.cuh')
ABI_INCLUDE = os.path.join(os.path.dirname(HERE), 'include')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '--fmad=false', '-shared', '-Xcompiler', '-fPIC']


def _nvcc() -> str:
    for cand in (shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: libtegb200.so cannot be built')


def _write_header_inc() -> str:
    """Embed teg_b200_runtime.cuh so NVRTC can serve `#include "teg_b200_runtime.cuh"` from memory."""
    inc = os.path.join(CSRC, 'tegb200_runtime_header.inc')
    with open(RUNTIME_HEADER) as f:
        text = f.read()
    assert ')TEGB"' not in text
    body = 'static const char TEGB_RUNTIME_HEADER[] = R"TEGB(' + text + ')TEGB";\n'
    if not os.path.exists(inc) or open(inc).read() != body:
        with open(inc, 'w') as f:
            f.write(body)
    return inc


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, 'tegb200.cu'), RUNTIME_HEADER, os.path.join(ABI_INCLUDE, 'tegb200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build_runtime(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    _write_header_inc()
    cuda_lib = '/usr/local/cuda/lib64'
    cmd = [_nvcc(), *NVCC_FLAGS, f'-I{ABI_INCLUDE}', f'-I{CSRC}', os.path.join(CSRC, 'tegb200.cu'),
           '-o', LIB_PATH, '-ldl']
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'nvcc failed:\n{" ".join(cmd)}\n{r.stdout}\n{r.stderr}')
    if verbose:
        print(r.stderr)
    return LIB_PATH


if __name__ == '__main__':
    print(build_runtime(force=True, verbose=True))
