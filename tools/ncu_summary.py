#!/usr/bin/env python
"""Print the metrics that matter for this path from an .ncu-rep (ncu -i rep --page raw --csv)."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'sm__cycles_elapsed.max']
rows = list(csv.reader(subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout.splitlines()))
hdr = rows[0]
for r in rows[2:]:
    print('==', r[hdr.index('Kernel Name')])
    for i, h in enumerate(hdr):
        if h in WANT or (h.startswith('l1tex__') and ('pct_of_peak_sustained_elapsed' in h or 'pct_of_peak_sustained_active' in h)) or 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h:
            try:
                v = float(r[i])
            except ValueError:
                continue
            if 'issue_stalled' in h and v < 0.3:
                continue
            if h.startswith('l1tex__') and h not in WANT and (v < 5.0 or '.max.' in h or '.min.' in h or ('.sum.' in h and h.replace('.sum.', '.avg.') in hdr)):
                continue
            print(f'  {h:90s} {r[i]} {rows[1][i]}')
