#!/usr/bin/env python
"""Headline metrics of the first kernel of an .ncu-rep (read here, no GPU): duration, instructions,
issue utilisation, occupancy, DRAM bytes, stall reasons per issue.  usage: python tools/ncu_brief.py x.ncu-rep"""
import csv,sys,subprocess,io
rep=sys.argv[1]
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
hdr=rows[0]; vals=rows[2]
print('kernel:', vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?')
want=['gpu__time_duration.sum','smsp__inst_executed.sum','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','dram__bytes_read.sum','dram__bytes_write.sum','smsp__warps_eligible.avg.per_cycle_active','smsp__warps_active.avg.per_cycle_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','launch__grid_size','launch__block_size','launch__waves_per_multiprocessor','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio']
for w in want:
    if w in hdr: print('%-86s %s %s'%(w, vals[hdr.index(w)], rows[1][hdr.index(w)]))
