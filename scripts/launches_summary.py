#!/usr/bin/env python
"""gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum --csv) -> the per-kernel table kept in profiles/."""
import collections
import csv
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
lines = [l for l in open(path) if not l.startswith("==")]
agg, n = collections.OrderedDict(), 0
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
    agg.setdefault((name, row["Grid Size"], row["Block Size"]), []).append(v)
    n += 1
tot = sum(sum(v) for v in agg.values())
title = sys.argv[2] if len(sys.argv) > 2 else "launch list"
print(f"# {title}")
print("# ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none (per-launch times are cold-cache and serialised: compare SHARES)")
print(f"# total device time {tot:.1f} us over {n} launches")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{sum(v):11.1f} us {100 * sum(v) / tot:5.1f}%  n={len(v):4d} avg={sum(v) / len(v):10.2f} us  {k[0]} grid={k[1]} block={k[2]}")
