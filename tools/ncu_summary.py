#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): headline metrics + SASS opcode mix of one kernel.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-index]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.max', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active']
r = rows[2 + idx]
for k in keys:
    if k in hdr:
        print("%-90s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))

sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                      capture_output=True, text=True).stdout
res, cur, h = [], [], None
for row in csv.reader(io.StringIO(sass)):
    if row and row[0] == 'Address':
        h = row
        if cur:
            res.append(cur)
        cur = []
        continue
    if h and len(row) == len(h):
        cur.append(row)
if cur:
    res.append(cur)
R = res[idx]
iE, iT, iS = h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('Source')
tot = sum(int(x[iE]) for x in R)
totT = sum(int(x[iT]) for x in R)
print("warp instructions %d, avg active threads %.1f" % (tot, totT / max(tot, 1)))
ops, opsT = collections.Counter(), collections.Counter()
for x in R:
    t = x[iS].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0].rstrip(';')
    ops[op] += int(x[iE])
    opsT[op] += int(x[iT])
for op, c in ops.most_common(28):
    print("%-10s %10d %5.1f%%  avg threads %5.1f" % (op, c, 100.0 * c / tot, opsT[op] / max(c, 1)))
