#!/usr/bin/env python
"""Summarise an ncu report: headline raw metrics + per-SASS-region stall samples (landmark instructions)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
marks = sys.argv[2].split(',') if len(sys.argv) > 2 else ['UBLKCP','LDGSTS','UTCHMMA','STTM','LDTM','SYNCS','BAR.','UTCBAR','EXIT','STG','ATOM','RED.','LDG','MUFU']
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
hdr,units=rows[0],rows[1]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','sm__cycles_elapsed.avg','smsp__warps_eligible.avg.per_cycle_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct']
for r in rows[2:]:
    print('==', r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '')
    for i,h in enumerate(hdr):
        if h in keys or ('stalled' in h and 'per_issue_active' in h):
            print(f'  {h} [{units[i]}] = {r[i]}')
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
data=[]
for r in rows[2:]:
    if len(r)<len(hdr): break
    data.append(r)
tot=sum(int(r[ix['# Samples']] or 0) for r in data)
print('total samples',tot,'instructions',len(data))
acc=0; last=0
for i,r in enumerate(data):
    s=int(r[ix['# Samples']] or 0); acc+=s
    srcl=r[ix['Source']]
    if any(k in srcl for k in marks):
        print(i, 'cum',acc, '+',acc-last,'ex',r[ix['Instructions Executed']], srcl[:70]); last=acc
print('--- top')
for r in sorted(data,key=lambda r:-int(r[ix['# Samples']] or 0))[:25]:
    print(r[ix['# Samples']], r[ix['Instructions Executed']], r[ix['Source']][:70], '| lsb',r[ix['stall_long_sb']],'ssb',r[ix['stall_short_sb']],'wait',r[ix['stall_wait']],'br',r[ix['stall_branch_resolving']],'bar',r[ix['stall_barrier']],'math',r[ix['stall_math']],'mio',r[ix['stall_mio']])
