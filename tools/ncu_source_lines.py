#!/usr/bin/env python3
"""Executed warp instructions and stall samples per SOURCE LINE of every kernel in an .ncu-rep captured with
--import-source on (library built with -lineinfo):  tools/ncu_source_lines.py file.ncu-rep [top_n]
One block per (kernel launch, source file); launches of the same kernel repeat."""
import csv, subprocess, sys


def main():
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    secs, i = [], 0
    while i < len(rows):
        r = rows[i]
        if r and r[0] == "File Path":
            secs.append({"file": r[1], "fn": rows[i + 1][1], "hdr": rows[i + 2], "rows": []})
            i += 3
            continue
        if secs:
            secs[-1]["rows"].append(r)
        i += 1
    seen = set()
    for s in secs:
        key = (s["fn"], s["file"])
        h = s["hdr"]
        iln, isrc, iex, ism = h.index("Line No"), h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        per, tot, smp = [], 0, 0
        for r in s["rows"]:
            if len(r) <= iex or not r[iln].strip().isdigit():
                continue
            try:
                e, m = int(r[iex]), int(r[ism])
            except ValueError:
                continue
            tot += e
            smp += m
            per.append((e, m, int(r[iln]), r[isrc].strip()[:130]))
        if key in seen or tot == 0:
            continue
        seen.add(key)
        print(f"===== {s['fn'][:70]} | {s['file'].split('/')[-1]} | executed warp instructions {tot}, stall samples {smp}")
        for e, m, ln, src in sorted(per, reverse=True)[:top]:
            print(f"  {e:11d} {100.0 * e / tot:5.1f}%  smp {m:6d}  L{ln:<5d} {src}")


main()
