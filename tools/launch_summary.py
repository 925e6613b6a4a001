"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: total / count / average per kernel."""
import collections
import csv
import re
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:100]
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
print("%10s %6s %10s  %s" % ("total ms", "n", "avg ms", "kernel"))
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print("%10.3f %6d %10.4f  %s" % (t, n, t / n, k))
