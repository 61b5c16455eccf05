"""SASS census of libb200at.so: which Blackwell-native instructions each kernel contains (B200_PROFILING.md table).

    python profiles/sass_census.py > profiles/r02_sass_census.txt

UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA load/store, UTCBAR = tcgen05.commit,
HMMA = mma.sync (legacy tensor path), LDSM/STSM = ldmatrix/stmatrix, FFMA2 = packed fp32 FMA, SYNCS = mbarrier."""
import collections
import os
import re
import subprocess

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'revisiting-at_b200', 'csrc', 'libb200at.so')
PAT = re.compile(r'\b(UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|UTMASTG|UTCBAR|HMMA|LDSM|STSM|FFMA2|SYNCS|UBLKCP|LDGSTS|REDG|ATOMG|ATOMS|UCGABAR_ARV)\b')


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            order.append(cur)
            continue
        if cur:
            for op in PAT.findall(line):
                counts[cur][op] += 1
    names = subprocess.run(['c++filt'], input='\n'.join(order), capture_output=True, text=True).stdout.splitlines()
    for mangled, name in zip(order, names):
        demangle[mangled] = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', name).split('(')[0]
    print(f'# {os.path.relpath(LIB)}: {len(order)} kernels (sm_100a); instruction counts in the SASS of each kernel')
    for mangled in sorted(order, key=lambda k: demangle[k]):
        c = counts[mangled]
        tags = ' '.join(f'{k}={v}' for k, v in sorted(c.items())) or '-'
        print(f'{demangle[mangled][:70]:70s} {tags}')


if __name__ == '__main__':
    main()
