#!/usr/bin/env python3
"""Summarise an ncu report exported as CSV (raw page + source page): key metrics, stall reasons and
executed instructions per warp-step by code region.  usage: read_prof.py raw.csv src.csv n_paths n_steps"""
import csv
import sys

raw, src, n_paths, n_steps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
rows = list(csv.reader(open(raw)))
d = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__shared_mem_per_block_dynamic', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__cycles_elapsed.avg.per_second',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
for k in keys:
    if k in d:
        print(f"{k:85s} {d[k][0]:12s} {d[k][1]}")
for h in sorted(d):
    if h.startswith('smsp__average_warps_issue_stalled') and float(d[h][1]) > 0.05:
        print(f"{h:100s} {d[h][1]}")
rows = list(csv.reader(open(src)))
ix = {h: i for i, h in enumerate(rows[1])}
data = rows[2:]
warp_steps = n_paths * n_steps / 32
seg = [(r[ix['Source']].strip()[:60], int(r[ix['Instructions Executed']]), int(r[ix['# Samples']])) for r in data]
bounds = [0]
for i in range(1, len(seg)):
    a, b = seg[i - 1][1], seg[i][1]
    if a == 0 or b == 0 or max(a, b) / max(1, min(a, b)) > 1.8:
        bounds.append(i)
bounds.append(len(seg))
for lo, hi in zip(bounds[:-1], bounds[1:]):
    n = sum(x[1] for x in seg[lo:hi])
    if n / warp_steps > 0.4:
        s = sum(x[2] for x in seg[lo:hi])
        print(f"[{lo:4d},{hi:4d}) n_instr={hi-lo:4d} exec/instr={seg[lo][1]:>12d} inst/step={n/warp_steps:7.2f} samples={s}  first: {seg[lo][0]}")
print("total inst/step", sum(x[1] for x in seg) / warp_steps, "total samples", sum(x[2] for x in seg))
