#!/usr/bin/env python
"""Summarise one or more `ncu --set full` reports (first kernel of each) as text:
duration, DRAM traffic, issue utilisation, occupancy, stall breakdown, bank conflicts.
    python scripts/ncu_summary.py gpurun_out/full_x.ncu-rep [...] > profiles/rNN_ncu_x.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
    'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_fma.sum', 'smsp__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_lsu.sum',
    'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum',
    'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(path, ': no data')
            continue
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        print('==', path, '::', d.get('Kernel Name', ('', '?'))[1][:90])
        for k in KEYS:
            if k in d:
                print('  %-62s %s %s' % (k, d[k][1], d[k][0]))
        stalls = [(h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')], float(v))
                  for h, u, v in zip(hdr, units, vals)
                  if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')
                  and 'not_issued' not in h]
        stalls.sort(key=lambda t: -t[1])
        print('  stall cycles per issued instruction: ' + ', '.join('%s %.2f' % s for s in stalls[:8]))


if __name__ == '__main__':
    main()
