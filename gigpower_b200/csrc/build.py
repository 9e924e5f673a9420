"""Build libgigpower_b200.so in-tree with nvcc for sm_100a (no torch dependency)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, 'libgigpower_b200.so')
SOURCES = ['gp_fbs.cu', 'gp_fbs_jit.cu', 'gp_nr3.cu', 'gp_study.cu', 'gp_peer.cu']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
         '-Xcompiler', '-fPIC', '-shared', '--fmad=true']
VERBOSE_FLAGS = ['-Xptxas=-v']      # register / spill report, only when asked for


def build(verbose=False, force=False):
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    deps = srcs + [os.path.join(HERE, 'gp_common.cuh'),
                   os.path.join(os.path.dirname(PKG), 'include', 'gigpower_b200.h')]
    deps += [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith('.cuh')]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    cmd = [NVCC] + FLAGS + (VERBOSE_FLAGS if verbose else []) + srcs + ['-o', OUT, '-lcudart', '-ldl']
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(' '.join(cmd) + '\n' + res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed building libgigpower_b200.so')
    return OUT


if __name__ == '__main__':
    build(verbose=True, force=True)
    print(OUT)
