#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` dump: one block per kernel launch with the
metrics the roofline / stall analysis needs.  usage: ncu_summary.py raw.csv"""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed.sum','sm__inst_executed_pipe_lsu.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__throughput.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed']
idx={h:i for i,h in enumerate(hdr)}
want+=[h for h in hdr if ('pipe_tensor' in h or 'tensor' in h.lower()) and h not in want]      # every tensor-pipe metric the report holds
stall=[h for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h]
for r in rows[2:]:
    print('----')
    for w in want:
        if w in idx: print(f"{w:80s} {r[idx[w]][:90]} {units[idx[w]]}")
    st=sorted(((float(r[idx[h]].replace(',','') or 0),h) for h in stall), reverse=True)[:6]
    for v,h in st: print(f"   stall {h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):30s} {v:.2f}")
