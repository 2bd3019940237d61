#!/usr/bin/env python3
"""SASS of the hot kernels of libbrille_b200.so: mnemonic histogram + the instructions that prove the memory path
(UBLKCP = bulk asynchronous copy / TMA without tensor map, SYNCS = mbarrier, STG.E.ENL2.256 = 32-byte stores, ATOMS, REDUX ...).

    python profiles/sass_summary.py [full]      ->  profiles/sass_<kernel>.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "brille_b200", "libbrille_b200.so")
HOT = {
    "k_interp_cell_tma_4_full": "k_interp_cell_tmaILi4ELb0E",
    "k_interp_cell_tma_4_sf": "k_interp_cell_tmaILi4ELb1E",
    "k_locate_trellis_split": "k_locateILi0ELb1E",
    "k_trellis_in_node_coop": "k_trellis_in_node_coop",
    "k_locate_in_node_nest": "k_locate_in_nodeILi1E",
    "k_pair_match": "k_pair_match",
    "k_powder_bin": "k_powder_bin",
}
PROOF = re.compile(r"UBLKCP|SYNCS|UTMA|STG\.E\.ENL2\.256|STG\.E\.128|ATOMS|ATOMG|RED\.|REDUX|MATCH|LDGSTS|DFMA|SHFL")


def main():
    full = len(sys.argv) > 1 and sys.argv[1] == "full"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", sass)
    for name, key in HOT.items():
        blk = next((b for b in blocks if b.split("\n", 1)[0].find(key) >= 0), None)
        if blk is None:
            print("not found:", key)
            continue
        lines = [ln for ln in blk.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", ln)]
        ops = collections.Counter()
        for ln in lines:
            m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
            if m:
                ops[m.group(1).split(".")[0]] += 1
        out = [f"# {blk.splitlines()[0].strip()}", f"# {len(lines)} SASS instructions (cuobjdump -sass brille_b200/libbrille_b200.so, sm_100a)", "",
               "mnemonic histogram:"]
        out += [f"  {op:14s} {c}" for op, c in ops.most_common(40)]
        out += ["", "instructions that show the memory / synchronisation path (first 60 of each kind are listed):"]
        seen = collections.Counter()
        for ln in lines:
            m = PROOF.search(ln)
            if m and seen[m.group(0)] < (60 if m.group(0) not in ("DFMA", "SHFL") else 6):
                seen[m.group(0)] += 1
                out.append("  " + re.sub(r"\s+", " ", ln.strip()))
        if full or name == "k_interp_cell_tma_4_full":
            out += ["", "full listing:"] + [re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", ln.rstrip()) for ln in lines]
        open(os.path.join(ROOT, "profiles", f"sass_{name}.txt"), "w").write("\n".join(out) + "\n")
        print(name, len(lines), dict(seen))


if __name__ == "__main__":
    main()
