#!/usr/bin/env python3
"""Summarise an .ncu-rep: headline metrics + per-CUDA-source-line instruction/stall shares."""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name','gpu__time_duration.sum','launch__grid_size','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers',
 'sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum',
 'smsp__thread_inst_executed_per_inst_executed.ratio','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio']
for r in rows[2:]:
    print("----")
    for i, h in enumerate(hdr):
        if h in keys: print(f"{h:80s} {units[i]:14s} {r[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; hd = None; agg = []
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hd = r; idx = {h: i for i, h in enumerate(hd)}; continue
    if r[0] != '' and hd:
        def g(n):
            try: return float(r[idx[n]])
            except Exception: return 0.0
        agg.append((cur, int(r[0]), r[1].strip(), g('# Samples'), g('Instructions Executed'), g('Thread Instructions Executed')))
ti = sum(a[4] for a in agg); ts = sum(a[3] for a in agg); tt = sum(a[5] for a in agg)
print(f"total warp inst {ti:.3e} thread inst {tt:.3e} avg lanes {tt/max(ti,1):.2f} samples {ts:.0f}")
agg.sort(key=lambda a: -a[4])
for a in agg[:top]:
    print(f"{a[0][:14]:14s} L{a[1]:4d} inst {100*a[4]/ti:5.1f}% thr {100*a[5]/tt:5.1f}% samp {100*a[3]/ts:5.1f}% lanes {a[5]/max(a[4],1):5.1f} | {a[2][:84]}")
