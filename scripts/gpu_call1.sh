#!/bin/bash
# round 2, call 1: parity suite with the tightened tolerances, smoke, megakernel register-preload A/B, Nemo B=32 baseline + launch list
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -s > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
grep -E "passed|failed|error|FAILED|ERROR" gpurun_out/pytest.log | tail -40
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
echo "bench nemo exit $?"; cat gpurun_out/bench_nemo.json | head -c 3000; tail -3 gpurun_out/bench_nemo.err
MB200_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_nemo.csv \
  python bench.py --model mistral-nemo-12b --batch 32 --prefill 1024 --layers 4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_nemo.log 2>&1
echo "launches exit $?"; python scripts/launches_summary.py gpurun_out/launches_nemo.csv 2>/dev/null | head -30
