#!/usr/bin/env python3
"""Summarise an .ncu-rep (read on the CPU box): headline metrics + hot SASS regions.
usage: ncu_summary.py report.ncu-rep [--sass FROM TO]"""
import csv, subprocess, sys, math, io

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h, u, v = r[0], r[1], r[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__average_warp_latency_per_inst_issued.ratio", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores", "sass__inst_executed_global_loads", "sass__inst_executed_global_stores",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "sm__maximum_warps_per_active_cycle_pct"]
for i, k in enumerate(h):
    if k in want or k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and float(v[i] or 0) > 0.3 or k.startswith("launch__occupancy_limit"):
        print(f"{k:95s} {u[i]:12s} {v[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hh = rows[1]; data = rows[2:]
isrc = hh.index("Source"); iinst = hh.index("Instructions Executed"); isamp = hh.index("# Samples"); ithr = hh.index("Thread Instructions Executed")
tot = sum(int(x[iinst]) for x in data); tots = sum(int(x[isamp]) for x in data)
print("total warp-instr", tot, "samples", tots)
if "--sass" in sys.argv:
    a = int(sys.argv[sys.argv.index("--sass") + 1]); b = int(sys.argv[sys.argv.index("--sass") + 2])
    for n, x in enumerate(data[a:b]):
        c = int(x[iinst]); s = int(x[isamp])
        print(f"{n + a:4d} {c:9d} {s / tots * 100:5.2f}% {int(x[ithr]) / max(c, 1):5.1f} {x[isrc]}")
else:
    regions = []; cur = None
    for n, x in enumerate(data):
        c = int(x[iinst])
        if cur and c > 0 and abs(math.log((c + 1) / (cur['c'] + 1))) < 0.15:
            cur['n'] += 1; cur['tot'] += c; cur['s'] += int(x[isamp]); cur['end'] = n; cur['thr'] += int(x[ithr])
        else:
            cur = {'start': n, 'end': n, 'c': c, 'n': 1, 'tot': c, 's': int(x[isamp]), 'thr': int(x[ithr])}; regions.append(cur)
    for g in regions:
        if g['tot'] / tot > 0.004 or g['s'] / tots > 0.01:
            print(f"{g['start']:4d}-{g['end']:4d} n={g['n']:3d} exec/instr={g['c']:9d} share={g['tot'] / tot * 100:5.1f}% samples={g['s'] / tots * 100:5.1f}% thr={g['thr'] / max(g['tot'], 1):4.1f}")
