#!/bin/bash
# round 2, call 6: co-resident stream-K rings + fast MoE routing for decode: tests, Nemo B=32, Mixtral-8x7B B=8, linear micro-benchmark
set -u
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print({k: d[k] for k in ("value", "ms_per_step")}, "frac", d["roofline"]["frac"], "experts", d["roofline"].get("distinct_experts_per_layer"), "e2e", d["e2e"]["value"],
          "prefill", d["prefill"]["ms"], d["prefill"]["tflops"], d["clocks"])
except Exception as e:
    print("parse failed", e)
PY
}
for f in test_gpu_ops test_gpu_model test_gpu_expert_parallel; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q -p no:cacheprovider --timeout 900 -s > gpurun_out/pytest_$f.log 2>&1
  echo "pytest $f exit $?"
  grep -E "passed|failed|FAILED|ERROR|watchdog" gpurun_out/pytest_$f.log | sort | uniq -c | sort -rn | head -30
done
timeout 900 python bench.py --model mistral-nemo-12b --batch 32 --prefill 1024 --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/bench_nemo.json 2> gpurun_out/bench_nemo.err
echo "bench nemo exit $?"; show gpurun_out/bench_nemo.json; tail -3 gpurun_out/bench_nemo.err
timeout 600 python scripts/bench_linear.py 32 > gpurun_out/bench_linear_T32_cores.txt 2>&1; echo "bench_linear exit $?"; cat gpurun_out/bench_linear_T32_cores.txt
timeout 1200 python bench.py --model mixtral-8x7b --batch 8 --prefill 2048 --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/bench_mixtral_b8.json 2> gpurun_out/bench_mixtral_b8.err
echo "bench mixtral-8x7b B=8 exit $?"; show gpurun_out/bench_mixtral_b8.json; tail -3 gpurun_out/bench_mixtral_b8.err
MB200_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_nemo.csv \
  python bench.py --model mistral-nemo-12b --batch 32 --prefill 1024 --layers 4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_nemo.log 2>&1
echo "launches exit $?"; python scripts/launches_summary.py gpurun_out/launches_nemo.csv 2>/dev/null | grep -v "^#" | grep -v "at::" | head -16
