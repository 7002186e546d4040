"""Build the C-ABI CUDA library in-tree: biom3_b200/libbiom3_b200.so (sm_100a only).

nvcc cross-compiles here without a GPU; the built .so travels to the GPU box with the snapshot."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'model.cu')
DEPS = [os.path.join(HERE, 'csrc', f) for f in ('model.cu', 'ptx.cuh', 'gemm_tcgen05.cuh', 'attention.cuh', 'kernels.cuh', 'fp32_path.cuh')]
DEPS.append(os.path.join(os.path.dirname(HERE), 'include', 'biom3_b200.h'))
LIB = os.path.join(HERE, 'libbiom3_b200.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
    '-shared', '-Xcompiler', '-fPIC', '--use_fast_math=false' if False else '-Xptxas=-v',
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + ['-o', LIB, SRC, '-lcudart']
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed building libbiom3_b200.so')
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


def build_variant(name: str, defines) -> str:
    """A second copy of the library with extra -D options, biom3_b200/libbiom3_b200.<name>.so, for A/B runs of compile-time
    choices (select it with BIOM3_LIB=<path>)."""
    out = os.path.join(HERE, f'libbiom3_b200.{name}.so')
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + [f'-D{d}' for d in defines] + ['-o', out, SRC, '-lcudart']
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f'nvcc failed building {out}')
    return out


if __name__ == '__main__':
    if '--variant' in sys.argv:           # python -m biom3_b200.build --variant ew8 BIOM3_BF16_EPI_WARPS=8
        i = sys.argv.index('--variant')
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build(force='--force' in sys.argv, verbose=True))
