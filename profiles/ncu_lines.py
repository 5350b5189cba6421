"""Aggregates an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.
usage: python profiles/ncu_lines.py dump.csv [top N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
agg = {}
fname = ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        i_s = hdr.index("# Samples")
        i_i = hdr.index("Instructions Executed")
        i_w = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
        i_x = hdr.index("L1 Wavefronts Shared Excessive") if "L1 Wavefronts Shared Excessive" in hdr else None
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    if r[2] != "-":            # SASS rows repeat under the line row; the line row has Address "-"
        continue
    key = (fname, int(r[0]))
    a = agg.setdefault(key, [0, 0, 0, 0, r[1].strip()[:110]])
    a[0] += int(r[i_s] or 0)
    a[1] += int(r[i_i] or 0)
    if i_w is not None:
        a[2] += int(r[i_w] or 0)
        a[3] += int(r[i_x] or 0)
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
print(f"total samples {tot_s}, instructions {tot_i}")
print(f"{'file:line':28s} {'smp%':>6s} {'inst%':>6s} {'smemWF':>10s} {'excess':>9s}  source")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0] + ':' + str(k[1]):28s} {100 * a[0] / tot_s:6.2f} {100 * a[1] / tot_i:6.2f} {a[2]:10d} {a[3]:9d}  {a[4]}")
