#!/usr/bin/env python
"""Print the handful of ncu metrics we track per kernel from a .ncu-rep (raw page CSV)."""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__grid_size', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_elapsed', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second']
def main(path, kfilter=None):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        if kfilter and kfilter not in name:
            continue
        print('---', name[:90])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k); print(f'  {k:72s} {r[i]:>16s} {units[i]}')
        for i, h in enumerate(hdr):
            if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v >= 0.2:
                    print(f'  stall {h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]:30s} {v:8.2f}')
if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
