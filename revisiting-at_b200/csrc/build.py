"""Build libb200at.so for sm_100a, in-tree (the .so travels to the GPU box with the snapshot).

    python revisiting-at_b200/csrc/build.py [--force] [--verbose]

One translation unit per kernel family.  The attack TU is built with -fmad=false: its update
arithmetic must round every operation like eager fp32.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, 'libb200at.so')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
COMMON = ['-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']
UNITS = [
    # (source, extra flags)
    ('b200at_attack.cu', ['-fmad=false']),
    ('b200at_convnext.cu', []),
    ('b200at_dwconv_mma.cu', []),
    ('b200at_gemm.cu', []),
    ('b200at_mlp.cu', []),
    ('b200at_stem.cu', []),
    ('b200at_attention.cu', []),
]


def _deps():
    return [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(('.cu', '.cuh', '.h', '.py'))] + \
           [os.path.join(HERE, '..', '..', 'include', h) for h in ('b200at.h', 'b200at_model.h')]


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in _deps()):
        return LIB
    objs = []
    for src, extra in UNITS:
        obj = os.path.join(HERE, src.replace('.cu', '.o'))
        cmd = ['nvcc'] + ARCH + COMMON + extra + (['-Xptxas', '-v'] if verbose else []) + \
              ['-c', os.path.join(HERE, src), '-o', obj]
        subprocess.check_call(cmd)
        objs.append(obj)
    subprocess.check_call(['nvcc'] + ARCH + ['-shared', '-o', LIB] + objs + ['-lcudart'])
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
