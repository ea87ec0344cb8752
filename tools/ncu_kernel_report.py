#!/usr/bin/env python
"""Per-kernel evidence table from an .ncu-rep (ncu --set full): time, DRAM bytes, throughputs, pipe utilisation, occupancy and
the warp-stall breakdown (stall reasons as average warps per issue-active cycle), one markdown section per kernel launch.

usage: ncu_kernel_report.py <report.ncu-rep> [kernel substring]"""
import csv
import subprocess
import sys

GENERAL = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'), ('launch__block_size', 'block'), ('launch__registers_per_thread', 'registers / thread'),
    ('launch__shared_mem_per_block_dynamic', 'dynamic smem / block'), ('launch__occupancy_limit_registers', 'CTAs/SM limit (registers)'),
    ('launch__occupancy_limit_shared_mem', 'CTAs/SM limit (smem)'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy'),
    ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput (% of peak)'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput (% of peak)'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_active', 'L1/TEX throughput (% of peak)'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput (% of peak)'),
    ('smsp__inst_executed.sum', 'warp instructions'), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy'),
    ('TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'tensor pipe active'),
    ('TPC.TriageCompute.sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'fp64 pipe active'),
    ('TPC.TriageCompute.sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed', 'ALU pipe'),
    ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'FMA pipe'),
    ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'LSU pipe'),
    ('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'LSU data-stage wavefronts (% of peak)'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'shared-memory bank conflicts'),
]
STALLS = ['barrier', 'long_scoreboard', 'short_scoreboard', 'wait', 'math_pipe_throttle', 'mio_throttle', 'lg_throttle', 'tex_throttle',
          'not_selected', 'no_instruction', 'branch_resolving', 'dispatch_stall', 'drain', 'imc_miss', 'membar', 'sleeping', 'selected']


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ''
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for n, r in enumerate(rows[2:]):
        name = r[idx['Kernel Name']].split('(')[0].replace('void ', '').replace('<unnamed>::', '').replace('ssdk::', '')
        if sub and sub not in name:
            continue
        print('### launch %d: `%s`\n' % (n, name))
        print('| metric | value |\n|---|---|')
        for k, label in GENERAL:
            if k in idx and r[idx[k]] not in ('', 'n/a'):
                v = r[idx[k]]
                try:
                    v = '%.4g' % float(v.replace(',', ''))
                except ValueError:
                    pass
                print('| %s | %s %s |' % (label, v, units[idx[k]]))
        st = []
        for s_ in STALLS:
            k = 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % s_
            if k in idx and r[idx[k]] not in ('', 'n/a'):
                st.append((float(r[idx[k]].replace(',', '')), s_))
        tot = sum(v for v, _ in st) or 1.0
        st.sort(reverse=True)
        print('| warp stalls (share of stalled warp-cycles) | %s |' % ', '.join('%s %.0f%%' % (s_, 100 * v / tot) for v, s_ in st[:6]))
        print()


if __name__ == '__main__':
    main()
