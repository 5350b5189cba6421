"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python profiles/launches_summary.py launches.csv > launches_summary.txt"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
tot = collections.OrderedDict()
first = {}
n = 0
for r in rows[1:]:
    if len(r) <= i_val or not r[i_val]:
        continue
    name = re.sub(r"\(.*$", "", r[i_name]).replace("void ", "")
    ms = float(r[i_val].replace(",", "")) * scale.get(r[i_unit], 1e-6)
    a = tot.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
    first.setdefault(name, n)
    n += 1
total = sum(a[1] for a in tot.values())
print(f"total {total:.1f} ms over {n} launches")
for name, (cnt, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:70]:70s} n={cnt:4d}  total {ms:9.2f} ms  avg {ms / cnt:8.3f} ms  {100 * ms / total:5.1f}%  first at launch #{first[name]}")
