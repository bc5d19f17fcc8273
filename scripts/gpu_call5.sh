#!/bin/bash
# round 2, call 5: programmatic dependent launch on the batched decode path (tests, Nemo B=32 with/without PDL, launch list), Mixtral-8x7B B=8 with distinct prompts
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
for f in test_gpu_ops test_gpu_model; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q -p no:cacheprovider --timeout 900 -s > gpurun_out/pytest_$f.log 2>&1
  echo "pytest $f exit $?"
  grep -E "passed|failed|FAILED|ERROR|watchdog" gpurun_out/pytest_$f.log | sort | uniq -c | sort -rn | head -30
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
for pdl in 1 0; do
  MB200_PDL=$pdl timeout 900 python bench.py --model mistral-nemo-12b --batch 32 --prefill 1024 --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/bench_nemo_pdl$pdl.json 2> gpurun_out/bench_nemo_pdl$pdl.err
  echo "bench nemo pdl=$pdl exit $?"; show gpurun_out/bench_nemo_pdl$pdl.json; tail -3 gpurun_out/bench_nemo_pdl$pdl.err
done
timeout 600 python scripts/bench_linear.py 32 > gpurun_out/bench_linear_T32_pdl.txt 2>&1; echo "bench_linear exit $?"; cat gpurun_out/bench_linear_T32_pdl.txt
timeout 1200 python bench.py --model mixtral-8x7b --batch 8 --prefill 2048 --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/bench_mixtral_b8.json 2> gpurun_out/bench_mixtral_b8.err
echo "bench mixtral-8x7b B=8 exit $?"; show gpurun_out/bench_mixtral_b8.json; tail -3 gpurun_out/bench_mixtral_b8.err
MB200_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_mixtral.csv \
  python bench.py --model mixtral-8x7b --batch 8 --prefill 2048 --layers 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_mixtral.log 2>&1
echo "launches mixtral exit $?"; python scripts/launches_summary.py gpurun_out/launches_mixtral.csv 2>/dev/null | grep -v "^#" | head -30
