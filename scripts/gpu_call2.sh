#!/bin/bash
# round 2, call 2: parity suite (one process per test file: a crash cannot poison the rest), smoke, megakernel preload A/B, Nemo B=32 (stream-K)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
for f in test_gpu_model test_gpu_ops test_gpu_expert_parallel; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q -p no:cacheprovider --timeout 900 -s > gpurun_out/pytest_$f.log 2>&1
  echo "pytest $f exit $?"
  grep -E "passed|failed|FAILED|ERROR|watchdog" gpurun_out/pytest_$f.log | sort | uniq -c | sort -rn | head -40
done
for pre in 1 0; do
  MB200_MK_PRELOAD=$pre timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline $([ $pre = 0 ] && echo --no-parity) > gpurun_out/bench_pre$pre.json 2> gpurun_out/bench_pre$pre.err
  echo "bench preload=$pre exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_pre$pre.json"))
    print({k:d[k] for k in ("value","ms_per_step")}, d["roofline"]["frac"], d["roofline"]["us_per_launch"], d["e2e"]["value"], d["prefill"]["ms"], d.get("parity"), d["clocks"])
except Exception as e:
    print("parse failed", e)
PY
  tail -3 gpurun_out/bench_pre$pre.err
done
timeout 900 python bench.py --model mistral-nemo-12b --batch 32 --prefill 1024 --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/bench_nemo.json 2> gpurun_out/bench_nemo.err
echo "bench nemo exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_nemo.json"))
    print({k:d[k] for k in ("value","ms_per_step")}, d["roofline"]["frac"], d["e2e"]["value"], d["prefill"])
except Exception as e:
    print("parse failed", e)
PY
tail -3 gpurun_out/bench_nemo.err
MB200_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_nemo.csv \
  python bench.py --model mistral-nemo-12b --batch 32 --prefill 1024 --layers 4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_nemo.log 2>&1
echo "launches exit $?"; python scripts/launches_summary.py gpurun_out/launches_nemo.csv 2>/dev/null | grep -v "^#" | head -24
