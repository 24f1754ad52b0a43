"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` launch list
per kernel: launches, mean time, mean DRAM bytes.  usage: python tools/ncu_launches.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for r in data:
    if len(r) > vi:
        agg[r[ki].split("(")[0].replace("<unnamed>::", "")][r[mi]].append(float(r[vi].replace(",", "")))
tot = 0.0
for k, v in agg.items():
    t = v["gpu__time_duration.sum"]
    rd, wr = v.get("dram__bytes_read.sum", [0.0]), v.get("dram__bytes_write.sum", [0.0])
    tot += sum(t)
    print(f"{k:28s} launches={len(t):4d} mean_us={sum(t) / len(t) / 1e3:9.1f} dram_read_MB={sum(rd) / len(rd) / 1e6:9.1f} "
          f"dram_write_MB={sum(wr) / len(wr) / 1e6:9.1f}")
print(f"total_us={tot / 1e3:.1f}")
