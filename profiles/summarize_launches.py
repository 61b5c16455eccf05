"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name.

    python profiles/summarize_launches.py gpurun_out/rXX_launches.csv [--top 40] [--rows A:B] > profiles/rXX_launches_summary.txt

--rows A:B keeps launches A..B-1 of the list (one step: from after one optimizer launch group to the end of the next).
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 40
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    rows = list(csv.DictReader(lines))
    if '--rows' in sys.argv:
        a, b = sys.argv[sys.argv.index('--rows') + 1].split(':')
        rows = rows[int(a):int(b)]
        path = f'{path} (launches {a}..{int(b) - 1})'
    for row in rows:
        try:
            v = float(row['Metric Value'].replace(',', ''))
        except (KeyError, ValueError):
            continue
        v *= {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(row['Metric Unit'], 1.0)
        name = row['Kernel Name'].replace('void ', '').replace('<unnamed>::', '')
        if name.startswith('at::'):
            m = re.search(r'(reduce_kernel|vectorized_layer_norm_kernel|layer_norm_grad\w*|GammaBeta\w*|\w+Functor\w*|'
                          r'\w*copy_kernel\w*|\w*[Gg]elu\w*|\w+_kernel_cuda\w*|index\w+|multi_tensor\w*)', name)
            name = 'at::' + (m.group(1) if m else name[4:64])
        else:
            name = re.sub(r'\(.*', '', name)[:80]
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f'# {path}: {sum(a[0] for a in agg.values())} launches, {tot / 1e6:.3f} ms of kernel time '
          f'(ncu per-launch times are cold-cache and serialised: compare shares, not absolutes)')
    print(f'{"ms":>10} {"share":>7} {"n":>6}  kernel')
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f'{t / 1e6:10.3f} {100 * t / tot:6.2f}% {c:6d}  {k}')


if __name__ == '__main__':
    main()
