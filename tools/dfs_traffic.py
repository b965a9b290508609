#!/usr/bin/env python3
"""DRAM traffic per k_phase_dfs launch from an ncu CSV log (metrics dram__bytes_read.sum, dram__bytes_write.sum,
gpu__time_duration.sum over one or more whole proofs):  tools/dfs_traffic.py log.csv instances out.json
Writes the average per launch (bytes read + written), which bench.py reports as roofline.traffic."""
import collections, csv, json, re, sys

def num(v, unit):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(unit, 1)

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]
ik, im, iu, iv, iid = (h.index(x) for x in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
per = collections.defaultdict(dict)
for r in rows[hi + 1:]:
    if len(r) > iv and "k_phase_dfs" in r[ik]:
        per[int(r[iid])][r[im]] = num(r[iv], r[iu])
n = len(per)
rd = sum(d.get("dram__bytes_read.sum", 0) for d in per.values())
wr = sum(d.get("dram__bytes_write.sum", 0) for d in per.values())
t = sum(d.get("gpu__time_duration.sum", 0) for d in per.values())
out = {"kernel": "k_phase_dfs", "workload": f"SHA256_64 x {sys.argv[2]}", "instances": int(sys.argv[2]), "launches": n,
       "dram_bytes_read_per_launch": rd / n, "dram_bytes_write_per_launch": wr / n, "traffic_bytes_per_launch": (rd + wr) / n,
       "ncu_seconds_per_launch": t / n,
       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_phase_dfs"}
json.dump(out, open(sys.argv[3], "w"), indent=1)
print(json.dumps(out))
