#!/usr/bin/env python
"""Summarise an ncu --metrics gpu__time_duration.sum --csv launch list: per-kernel totals and shares."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    tot = collections.OrderedDict()
    seq = []
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
        short = re.sub(r'\(.*', '', row['Kernel Name'])
        short = re.sub(r'^void |<unnamed>::|ssdk::', '', short)
        seq.append((short, v, row.get('Grid Size')))
        tot[short] = tot.get(short, 0) + v
    T = sum(tot.values())
    print('total %.1f us over %d launches' % (T, len(seq)))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print('%-40s %10.1f us %5.1f%%  (%d launches)' % (k, v, 100 * v / T, sum(1 for s in seq if s[0] == k)))
    if '-v' in sys.argv:
        for i, (n, v, g) in enumerate(seq):
            print(i, n[:40], '%.1f' % v, g)


if __name__ == '__main__':
    main()
