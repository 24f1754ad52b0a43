"""Per-launch table of an `ncu --csv --metrics ...` log: one line per launch with every metric captured.
    python tools/ncu_table.py launches.csv [kernel-substring]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, mi, vi, ii, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID"), hdr.index("Metric Unit")
per, units = collections.OrderedDict(), {}
for r in data:
    if len(r) > vi and (len(sys.argv) < 3 or sys.argv[2] in r[ki]):
        per.setdefault((int(r[ii]), r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")), {})[r[mi]] = float(r[vi].replace(",", ""))
        units[r[mi]] = r[ui]
names = sorted({m for v in per.values() for m in v})
short = lambda m: m.replace("smsp__", "").replace("sm__", "").replace(".pct_of_peak_sustained_active", "%").replace(".pct_of_peak_sustained_elapsed", "%e").replace("_executed", "").replace("gpu__time_duration.sum", "time_ns")
print("id kernel | " + " | ".join(f"{short(m)}[{units[m]}]" for m in names))
for (i, k), v in per.items():
    print(f"{i:3d} {k:32s} | " + " | ".join(f"{v.get(m, float('nan')):.4g}" for m in names))
