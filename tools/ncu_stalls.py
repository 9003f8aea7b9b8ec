#!/usr/bin/env python3
"""Per-instruction stall samples of an .ncu-rep (read on the CPU box): totals per stall reason and the top instructions
of each of the largest reasons.   usage: ncu_stalls.py report.ncu-rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 12
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hh = rows[1]; data = rows[2:]
cols = [c for c in hh if c.startswith("stall_") and "Not Issued" not in c]
tot = {c: sum(int(r[hh.index(c)] or 0) for r in data) for c in cols}
alls = sum(tot.values())
print("samples", alls, " ".join(f"{c[6:]}={v / alls * 100:.1f}%" for c, v in sorted(tot.items(), key=lambda x: -x[1]) if v / alls > 0.005))
for c, v in sorted(tot.items(), key=lambda x: -x[1])[:4]:
    i = hh.index(c)
    top = sorted(range(len(data)), key=lambda n: -int(data[n][i] or 0))[:N]
    print(f"--- {c} ({v / alls * 100:.1f}% of samples): top instructions")
    for n in sorted(top):
        print(f"  {n:5d} {int(data[n][i] or 0) / alls * 100:5.2f}%  {data[n][hh.index('Source')][:90]}")
