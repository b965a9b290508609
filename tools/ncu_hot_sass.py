#!/usr/bin/env python3
"""Executed-instruction mix of a kernel from an .ncu-rep captured with --import-source on:
   tools/ncu_hot_sass.py file.ncu-rep [min_exec_fraction]
Prints executed warp-instructions per opcode class and the stall samples per opcode class."""
import collections, csv, subprocess, sys

def cls(op):
    base = op.split(".")[0]
    if base == "IMAD":
        if ".WIDE" in op: return "IMAD.WIDE"
        if any(x in op for x in (".MOV", ".IADD", ".SHL")): return "IMAD.MOV/IADD/SHL"
        if ".HI" in op: return "IMAD.HI"
        if ".X" in op: return "IMAD.X"
    return base

def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    i_src, i_ex, i_smp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ex, smp = collections.Counter(), collections.Counter()
    for r in rows[2:]:
        if len(r) <= i_ex: continue
        toks = r[i_src].split()
        if not toks: continue
        op = toks[1] if toks[0].startswith("@") else toks[0]
        c = cls(op)
        ex[c] += int(r[i_ex]); smp[c] += int(r[i_smp])
    tot, tots = sum(ex.values()), sum(smp.values())
    print(f"executed warp-instructions {tot}, samples {tots}")
    for k, v in ex.most_common(24):
        print(f"  {k:20s} {v:12d} {100.0 * v / tot:5.1f}%   samples {100.0 * smp[k] / max(1, tots):5.1f}%")

main()
