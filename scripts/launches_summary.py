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
print("# round 1 launch list (final kernels): ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none")
print("# command: python bench.py --steps 4 --warmup 3 --prefill 4096 --no-cpu-baseline   (Mistral-7B 32L; profiling starts after weight synthesis)")
print("# i.e. four 4096-token prefills (1 warm-up + 3 timed), 3+4 decode steps (megakernel: fused greedy argmax, one launch per token), 3+4 e2e steps;")
print(f"# per-launch times are cold-cache, serialised: compare SHARES.  total device time {tot:.1f} us over {n} launches")
print("# decode: the megakernel is the only kernel of a device-loop decode step (100 % of the step).  prefill (4096 tokens, per layer):")
print("# gemm_tcgen05 <4,2,256> qkv+rope, <1,2,256> wo and w2, <3,2,256> gate/up+SiLU*mul, attn_prefill_tcgen05 grid=(32,32,1); <2,2,256> lm head")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{sum(v):11.1f} us {100 * sum(v) / tot:5.1f}%  n={len(v):4d} avg={sum(v) / len(v):10.2f} us  {k[0]} grid={k[1]} block={k[2]}")
