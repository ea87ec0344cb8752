#!/usr/bin/env python
"""Per-source-line shares of executed instructions and stall samples for one kernel of an .ncu-rep.

usage: ncu_lines.py <report.ncu-rep> <kernel substring> <cubin source name, e.g. encode> [top N] [mangled substring]

The report's SASS page (instruction offsets + counters) is joined with `nvdisasm -g` of the cubin extracted from the in-tree
libssdk.so (which must be the build the report was taken from: the tool checks that the opcodes agree)."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_rows(rep, kern):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    res, hdr, take = [], None, False
    for r in rows:
        if r and r[0] == 'Kernel Name':
            take = kern in r[1]
            continue
        if r and r[0] == 'Address':
            hdr = r
            continue
        if take and hdr and len(r) >= len(hdr) - 2:
            res.append(dict(zip(hdr, r)))
        if res and not take:
            break
    return res


def line_table(cubin_name, kern):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(['cuobjdump', '-xelf', 'all', os.path.join(ROOT, 'ssd_keras_b200', '_lib', 'libssdk.so')], cwd=d, capture_output=True)
        f = [x for x in os.listdir(d) if x.startswith(cubin_name + '.')][0]
        txt = subprocess.run(['nvdisasm', '-g', os.path.join(d, f)], capture_output=True, text=True).stdout
    tab, cur, inside = [], None, False
    for l in txt.splitlines():
        if l.startswith('\t.section\t.text.'):
            inside = kern in l
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m:
            tab.append((int(m.group(1), 16), m.group(2).strip(), cur))
    return tab


def main():
    rep, kern, cub = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = sass_rows(rep, kern)
    tab = line_table(cub, sys.argv[5] if len(sys.argv) > 5 else kern)
    if len(rows) != len(tab):
        print('warning: %d profiled instructions vs %d in the in-tree build' % (len(rows), len(tab)))
    acc = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
    stall_cols = [k for k in rows[0] if k.startswith('stall_')]
    seen = set()
    stall_cols = [k for k in stall_cols if not (k in seen or seen.add(k))]
    mism = 0
    for r, (off, op, loc) in zip(rows, tab):
        if r['Source'].split()[0:1] != op.split()[0:1] and r['Source'].split()[1:2] != op.split()[0:1]:
            mism += 1
        a = acc[loc]
        a[0] += float(r['Instructions Executed'] or 0)
        a[1] += float(r['# Samples'] or 0)
    if mism:
        print('warning: %d opcode mismatches (report is of another build?)' % mism)
    ti = sum(a[0] for a in acc.values()) or 1
    ts = sum(a[1] for a in acc.values()) or 1
    print('warp instructions %.4g   samples %d' % (ti, ts))
    src = {}
    for loc, a in sorted(acc.items(), key=lambda kv: -kv[1][1])[:top]:
        if loc is None:
            continue
        fn = os.path.join(ROOT, 'ssd_keras_b200', 'csrc', loc[0])
        if fn not in src and os.path.exists(fn):
            src[fn] = open(fn).read().splitlines()
        text = src.get(fn, [''] * (loc[1] + 1))[loc[1] - 1].strip() if fn in src else ''
        print('%-12s %5d  inst %5.1f%%  samples %5.1f%%  %s' % (loc[0], loc[1], 100 * a[0] / ti, 100 * a[1] / ts, text[:100]))


if __name__ == '__main__':
    main()
