#!/usr/bin/env python
"""Group an `ncu --csv --metrics ...` launch list by (kernel, grid): median duration and last DRAM bytes.
    python profiles/summarize_launches.py gpurun_out/launches.csv"""
import csv
import sys
from collections import OrderedDict

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(lines))
d = OrderedDict()
for r in rows:
    e = d.setdefault(r["ID"], {"name": r["Kernel Name"][:70], "grid": r["Grid Size"], "block": r["Block Size"]})
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    if r["Metric Unit"] == "Kbyte": v *= 1e3
    if r["Metric Unit"] == "Mbyte": v *= 1e6
    if r["Metric Unit"] == "Gbyte": v *= 1e9
    if r["Metric Unit"] in ("us", "usecond"): v *= 1e3
    if r["Metric Unit"] in ("ms", "msecond"): v *= 1e6
    e[r["Metric Name"]] = v
seen = OrderedDict()
for v in d.values():
    seen.setdefault((v["name"], v["grid"], v["block"]), []).append(v)
for (name, grid, block), vs in seen.items():
    ts = sorted(x.get("gpu__time_duration.sum", 0.0) for x in vs)
    t = ts[len(ts) // 2]
    rd = vs[-1].get("dram__bytes_read.sum", 0.0)
    wr = vs[-1].get("dram__bytes_write.sum", 0.0)
    print(f"{name:72s} grid {grid:>12s} blk {block:>12s} n={len(vs):3d} median {t / 1e3:9.1f} us"
          + (f"  dram rd {rd / 1e6:8.1f} MB wr {wr / 1e6:8.1f} MB" if rd + wr else ""))
