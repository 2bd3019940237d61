#!/usr/bin/env python3
"""Samples and executed instructions PER SOURCE LINE of one kernel.

    ncu -i X.ncu-rep --page source --csv --kernel-id :::N > k.csv
    python profiles/sass_lines.py k.csv brille_b200/libbrille_b200.so <kernel-name-substring> [top]

The ncu source page lists the SASS of the kernel in address order with its sampling data but without line numbers; nvdisasm -g
of the cubin (extracted from the shared object with cuobjdump) lists the same SASS with `//## File ..., line N` marks (the
innermost inlined location).  Joined by instruction index.  Built with -lineinfo.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def line_table(so, kernel):
    if so.endswith(".cubin"):
        tmp, names = os.path.dirname(os.path.abspath(so)), [os.path.basename(so)]
    else:
        tmp = tempfile.mkdtemp()
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
        names = sorted(os.listdir(tmp))
    for cubin in names:
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
        lines = out.splitlines()
        start = None
        for i, ln in enumerate(lines):
            if ln.startswith(".text.") and kernel in ln and ln.rstrip().endswith(":"):
                start = i
                break
        if start is None:
            continue
        table = []
        cur = (None, 0)
        for ln in lines[start + 1:]:
            if ln.startswith("//-----") or ln.startswith(".text."):
                break
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                table.append((int(m.group(1), 16), cur, m.group(2).strip()))
        return table
    raise SystemExit(f"kernel {kernel} not found in {so}")


def main():
    path, so, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    iS, iN, iI = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    iT = hdr.index("Thread Instructions Executed")
    stall_cols = [k for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    body = [r for r in rows[2:] if len(r) > iT and r[iI].isdigit()]
    table = line_table(so, kernel)
    if len(body) == 2 * len(table):  # (ncu lists the function twice when the report holds the source and the SASS view)
        body = body[: len(table)]
    if len(table) < len(body):
        print(f"warning: {len(body)} profiled instructions, {len(table)} disassembled", file=sys.stderr)
    smp, ins, thr = collections.Counter(), collections.Counter(), collections.Counter()
    stalls = collections.defaultdict(collections.Counter)
    for k, r in enumerate(body):
        loc = table[k][1] if k < len(table) else (None, 0)
        smp[loc] += int(r[iN])
        ins[loc] += int(r[iI])
        thr[loc] += int(r[iT])
        for c in stall_cols:
            if r[c].isdigit() and int(r[c]):
                stalls[loc][hdr[c]] += int(r[c])
    ts, ti = sum(smp.values()) or 1, sum(ins.values()) or 1
    print(f"{kernel}: {len(body)} SASS instructions, {ti:.4g} warp instructions executed, {ts} samples")
    src = {}
    print("  %smp  %inst lanes  file:line  top stalls | source")
    for loc, c in smp.most_common(top):
        f, ln = loc
        text = ""
        if f:
            if f not in src:
                here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
                for root in (os.environ.get("SASS_LINES_SRC", ""), os.path.join(here, "brille_b200/csrc"), os.path.join(here, "include")):
                    p = os.path.join(root, f)
                    if os.path.exists(p):
                        src[f] = open(p).read().splitlines()
                        break
                else:
                    src[f] = []
            text = src[f][ln - 1].strip()[:90] if 0 < ln <= len(src[f]) else ""
        st = ", ".join(f"{k.replace('stall_', '')} {100 * v / max(c, 1):.0f}%" for k, v in stalls[loc].most_common(2))
        lanes = thr[loc] / ins[loc] if ins[loc] else 0
        print(f"{100 * c / ts:6.2f} {100 * ins[loc] / ti:6.2f} {lanes:5.1f}  {f}:{ln}  [{st}] | {text}")


if __name__ == "__main__":
    main()
