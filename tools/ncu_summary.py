#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small markdown table for profiles/."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg']


def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print('| # | kernel | ' + ' | '.join(k.split('.')[-3] if k.startswith('TPC') else k for k in KEYS if k in idx) + ' |')
    print('|---' * (2 + sum(k in idx for k in KEYS)) + '|')
    for n, r in enumerate(rows[2:]):
        name = r[idx['Kernel Name']].split('(')[0]
        vals = ['%s %s' % (r[idx[k]], units[idx[k]]) for k in KEYS if k in idx]
        print('| %d | %s | ' % (n, name) + ' | '.join(vals) + ' |')


if __name__ == '__main__':
    main()
