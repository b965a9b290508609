#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share.
usage: summarize_launches.py launches.csv [first_launch] [n_launches]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    n = int(sys.argv[3]) if len(sys.argv) > 3 else len(data)
    mi = hdr.index("Metric Name")
    data = [r for r in data if len(r) > vi and r[mi] == "gpu__time_duration.sum"][first:first + n]   # other metrics may share the log
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    tot, cnt, mx = collections.defaultdict(float), collections.Counter(), collections.defaultdict(float)
    for r in data:
        us = float(r[vi].replace(",", "")) * scale[r[ui]]
        name = r[ki].split("(")[0].replace("void ", "")
        tot[name] += us
        cnt[name] += 1
        mx[name] = max(mx[name], us)
    T = sum(tot.values())
    print(f"launches {len(data)}, total device time {T / 1e3:.3f} ms (per-launch times are cold-cache and serialised under ncu)")
    print(f"{'kernel':32s} {'launches':>8s} {'total us':>12s} {'share':>7s} {'max us':>10s}")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"{k:32s} {cnt[k]:8d} {v:12.1f} {100 * v / T:6.1f}% {mx[k]:10.1f}")


if __name__ == "__main__":
    main()
