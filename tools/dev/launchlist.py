#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py: the last bf16x3 step, per kernel.

    python tools/dev/launchlist.py gpurun_out/launches.csv [summary.csv]
"""
import collections
import csv
import re
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    h = rows[hdr]
    ki, vi = h.index('Kernel Name'), h.index('Metric Value')
    data = rows[hdr + 1:]
    names = [r[ki] for r in data]
    times = [float(r[vi]) / 1000.0 for r in data]
    idx = [i for i, n in enumerate(names) if 'adam_iterate_dev' in n]
    steps = [(idx[i - 1] + 1, idx[i] + 1) for i in range(1, len(idx))]
    x3 = [s for s in steps if any('split' in n for n in names[s[0]:s[1]])]
    lo, hi = x3[-1]
    seq = []
    agg = collections.OrderedDict()
    for n, t in zip(names[lo:hi], times[lo:hi]):
        k = re.sub(r'\(.*', '', n).replace('void ', '')[:52]
        seq.append((k, t))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    lines = ['# one graph-replayed C3 step, conv_math=bf16x3 (9 views batched), kernels aggregated by name; cold-cache serialised times: compare SHARES',
             '# %d kernels, sum %.1f us' % (hi - lo, tot), 'kernel,launches,total_us,share']
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append('%s,%d,%.1f,%.4f' % (k.replace(',', ';'), c, t, t / tot))
    print('\n'.join(lines))
    if len(sys.argv) > 2:
        open(sys.argv[2], 'w').write('# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --workload C3 --steps 2 --warmup 3 --no-cpu-baseline --quick\n' + '\n'.join(lines) + '\n')
    print('# in launch order:')
    for k, t in seq:
        print('#   %-52s %7.1f' % (k, t))


if __name__ == '__main__':
    main()
