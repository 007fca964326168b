#!/usr/bin/env python
"""ncu report -> markdown table of the metrics DESIGN.md / bench.py cite (one column per captured launch)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic',
        'sm__cycles_elapsed.avg', 'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
keys += [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h]
print('| metric | unit | ' + ' | '.join(f'launch {i + 1}' for i in range(len(rows) - 2)) + ' |')
print('|---|---|' + '---|' * (len(rows) - 2))
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print(f'| {k} | {units[i]} | ' + ' | '.join(r[i] for r in rows[2:]) + ' |')
