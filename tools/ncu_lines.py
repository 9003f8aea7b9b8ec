#!/usr/bin/env python3
"""Per-source-line view of an .ncu-rep captured with --import-source on (-lineinfo build):
share of executed warp-instructions and of stall samples per CUDA source line, top N by instructions.
usage: ncu_lines.py report.ncu-rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hh = rows[hi]
iinst = hh.index("Instructions Executed"); isamp = hh.index("# Samples")
def num(x):
    try: return int(x)
    except ValueError: return 0
lines = [r for r in rows[hi + 1:] if len(r) > iinst and r[0].strip().isdigit()]
tot = sum(num(r[iinst]) for r in lines) or 1; tots = sum(num(r[isamp]) for r in lines) or 1
print(f"total warp-instr {tot}  samples {tots}  ({len(lines)} source lines with code)")
ranked = sorted(lines, key=lambda r: -num(r[iinst]))[:top]
for r in sorted(ranked, key=lambda r: int(r[0])):
    print(f"{r[0]:>5s} inst {num(r[iinst]) / tot * 100:5.2f}%  samples {num(r[isamp]) / tots * 100:5.2f}%  {r[1].strip()[:140]}")
