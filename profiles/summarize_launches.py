"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list into the per-kernel table kept under profiles/."""
import collections
import csv
import re
import sys

rows = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(rows):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("unnamed>::", "")
    v, u = float(row["Metric Value"].replace(",", "")), row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
    a = agg.setdefault(k, [0, 0.0, 1e30, 0.0])
    a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total us | mean us | min us | max us | share |")
print("|---|---:|---:|---:|---:|---:|---:|")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("| `%s` | %d | %.1f | %.1f | %.1f | %.1f | %.3f |" % (k, a[0], a[1], a[1] / a[0], a[2], a[3], a[1] / tot))
