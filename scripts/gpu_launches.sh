#!/bin/bash
# launch list of one 4096-token prefill (first N launches after profiling starts)
MB200_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 4 --warmup 3 --prefill 4096 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
python - <<'PY'
import csv, collections, re
lines=[l for l in open('gpurun_out/launches.csv') if not l.startswith('==')]
agg=collections.OrderedDict()
for row in csv.DictReader(lines):
    name=re.sub(r'\(.*','',row['Kernel Name'])
    v=float(row['Metric Value'].replace(',','')); unit=row['Metric Unit']
    v = v/1000 if unit=='ns' else (v*1000 if unit=='ms' else v)
    agg.setdefault((name[:70], row['Grid Size']),[]).append(v)
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:10]:
    print(f"{sum(v):10.1f} us n={len(v):4d} avg={sum(v)/len(v):9.2f}  {k[0]} grid={k[1]}")
PY
