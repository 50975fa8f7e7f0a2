#!/usr/bin/env python3
"""Per-launch table of an ncu --csv launch list: kernel, time, DRAM bytes, warp instructions.
Usage: python tools/launches.py <launches.csv> [first_launch] [count]"""
import csv, sys
from collections import OrderedDict
lines = open(sys.argv[1]).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.reader(lines[start:]))
h = rows[0]
ki, mi, vi, idi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
per = OrderedDict()
for r in rows[1:]:
    if len(r) > vi:
        per.setdefault((int(r[idi]), r[ki].split("(")[0].replace("void ", "").replace("mlv::", "")), {})[r[mi]] = float(r[vi].replace(",", "") or 0)
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
count = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
tot = 0.0
for (i, k), m in per.items():
    if i < first or i >= first + count:
        continue
    t = m.get("gpu__time_duration.sum", 0) / 1e3
    tot += t
    print("%4d %-44s %8.1f us  rd %7.2f MB wr %7.2f MB  inst %9.0f k  regs %3d grid %6d" % (i, k[:44], t, m.get("dram__bytes_read.sum", 0) / 1e6, m.get("dram__bytes_write.sum", 0) / 1e6,
          m.get("smsp__inst_executed.sum", 0) / 1e3, m.get("launch__registers_per_thread", 0), m.get("launch__grid_size", 0)))
print("total %.1f us" % tot)
