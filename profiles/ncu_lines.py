"""Aggregate an ncu report's source page by (file, line): samples, instructions, main stall reasons.
usage: python profiles/ncu_lines.py report.ncu-rep out.tsv [top_n]"""
import collections
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 120
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
keys = ("# Samples", "Instructions Executed", "stall_no_inst", "stall_wait", "stall_short_sb", "stall_barrier",
        "stall_long_sb", "stall_selected", "stall_not_selected", "stall_math", "stall_mio", "stall_branch_resolving",
        "stall_dispatch", "stall_lg", "stall_membar", "stall_sleep")
agg, tot = collections.OrderedDict(), collections.Counter()
seen_addr = set()
for s, i0 in enumerate(hdr_idx):
    end = hdr_idx[s + 1] if s + 1 < len(hdr_idx) else len(rows)
    fname = rows[i0 - 2][1].split("/")[-1] if i0 >= 2 else ""
    hdr = rows[i0]
    col = {k: hdr.index(k) for k in keys if k in hdr}
    ac = hdr.index("Address")
    cur_line, cur_src = "", ""
    for r in rows[i0 + 1:end]:
        if len(r) <= max(col.values()):
            continue
        if r[0]:                        # a CUDA source line: the SASS rows that follow belong to it
            cur_line, cur_src = r[0], r[1].strip()[:100]
        if not r[3]:
            continue
        try:
            vals = {k: int(r[c] or 0) for k, c in col.items()}
        except ValueError:
            continue
        key = (fname, cur_line, cur_src)
        agg.setdefault(key, collections.Counter()).update(vals)
        tot.update(vals)
with open(out, "w") as f:
    f.write("# totals " + " ".join(f"{k}={v}" for k, v in tot.items()) + "\n")
    for (fn, ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top_n]:
        f.write(f"{fn}:{ln}\tsamp%={100 * a['# Samples'] / max(tot['# Samples'], 1):.2f}\tinst%="
                f"{100 * a['Instructions Executed'] / max(tot['Instructions Executed'], 1):.2f}\t" +
                " ".join(f"{k[6:]}={a[k]}" for k in keys[2:] if a[k]) + f"\t| {src}\n")
print("wrote", out)
