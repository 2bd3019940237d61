#!/usr/bin/env python3
"""Opcode histogram and hottest SASS instructions of an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv)."""
import csv
import collections
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iN, iI, iT = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
body = [r for r in rows[2:] if len(r) > iT and r[iI].isdigit()]
tot_s = sum(int(r[iN]) for r in body)
tot_i = sum(int(r[iI]) for r in body)
print(f"total warp-instructions {tot_i:.4g}  samples {tot_s}")
ops = collections.Counter()
smp = collections.Counter()
for r in body:
    src = r[iS].strip()
    parts = src.split()
    op = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "?")
    op = op.split(".")[0]
    ops[op] += int(r[iI])
    smp[op] += int(r[iN])
print("opcode            %instr  %samples")
for op, c in ops.most_common(28):
    print(f"{op:16s} {100*c/tot_i:6.2f}  {100*smp[op]/tot_s:6.2f}")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("\nhottest instructions by samples (index, %samples, %instr, sass)")
order = sorted(range(len(body)), key=lambda k: -int(body[k][iN]))[:n]
for k in sorted(order):
    r = body[k]
    print(f"{k:5d} {100*int(r[iN])/tot_s:5.2f} {100*int(r[iI])/tot_i:5.2f}  {r[iS].strip()[:100]}")
