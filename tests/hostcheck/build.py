"""Build tests/hostcheck/libhostcheck.so (host compile of the kernel bodies; test infrastructure)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, 'libhostcheck.so')


def build(force=False):
    src = os.path.join(HERE, 'hostcheck.cpp')
    deps = [src] + [os.path.join(HERE, '..', '..', 'revisiting-at_b200', 'csrc', f)
                    for f in ('b200at_bodies.cuh', 'b200at_math.cuh')]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    subprocess.check_call(['g++', '-O2', '-ffp-contract=off', '-fno-fast-math', '-Wno-unknown-pragmas', '-shared',
                           '-fPIC', '-x', 'c++', '-std=c++17', src, '-o', LIB])
    return LIB


if __name__ == '__main__':
    print(build(force=True))
