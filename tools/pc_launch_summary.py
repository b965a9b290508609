"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv) of `tools/prof_target.py pc LOG_LEN`."""
import collections, csv, sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
tot = collections.OrderedDict()
for r in csv.DictReader(lines):
    k = r["Kernel Name"].split("(")[0]
    v = float(r["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}[r["Metric Unit"]]
    t = tot.setdefault(k, [0, 0.0])
    t[0] += 1
    t[1] += v
s = sum(v[1] for v in tot.values())
print("%d launches, %.3f ms (each launch timed alone by ncu: cold caches, serialised)" % (sum(v[0] for v in tot.values()), s / 1e3))
for k, v in sorted(tot.items(), key=lambda x: -x[1][1]):
    print("  %-28s %4d launches %9.3f ms  %5.1f %%" % (k, v[0], v[1] / 1e3, 100 * v[1] / s))
