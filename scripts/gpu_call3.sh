#!/bin/bash
# round 2, call 3: megakernel structure A/B (round-1 kernel vs warpgroup/setmaxnreg build), op + model tests with the TMA decode attention,
# linear micro-benchmark, Nemo B=32, Mixtral-8x7B B=8 on one GPU, compute-sanitizer
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print({k: d[k] for k in ("value", "ms_per_step")}, "frac", d["roofline"]["frac"], "us", d["roofline"]["us_per_launch"], "e2e", d["e2e"]["value"],
          "prefill", d["prefill"]["ms"], d["prefill"]["tflops"], "parity", d.get("parity"), d["clocks"])
except Exception as e:
    print("parse failed", e)
PY
}
for lib in "" mkwg; do
  if [ -n "$lib" ]; then export MB200_LIB_PATH=$PWD/mistral_inference_b200/libmb200_$lib.so; else unset MB200_LIB_PATH; fi
  timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline $([ -n "$lib" ] && echo --no-parity) > gpurun_out/bench_mk_${lib:-r1}.json 2> gpurun_out/bench_mk_${lib:-r1}.err
  echo "bench megakernel ${lib:-r1} exit $?"; show gpurun_out/bench_mk_${lib:-r1}.json; tail -2 gpurun_out/bench_mk_${lib:-r1}.err
done
unset MB200_LIB_PATH
for f in test_gpu_ops test_gpu_model; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q -p no:cacheprovider --timeout 900 -s > gpurun_out/pytest_$f.log 2>&1
  echo "pytest $f exit $?"
  grep -E "passed|failed|FAILED|ERROR|watchdog" gpurun_out/pytest_$f.log | sort | uniq -c | sort -rn | head -30
done
timeout 600 python scripts/bench_linear.py 32 > gpurun_out/bench_linear_T32.txt 2>&1; echo "bench_linear exit $?"; cat gpurun_out/bench_linear_T32.txt
timeout 900 python bench.py --model mistral-nemo-12b --batch 32 --prefill 1024 --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/bench_nemo.json 2> gpurun_out/bench_nemo.err
echo "bench nemo exit $?"; show gpurun_out/bench_nemo.json; tail -3 gpurun_out/bench_nemo.err
MB200_ATTN_DECODE=plain timeout 900 python bench.py --model mistral-nemo-12b --batch 32 --prefill 1024 --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/bench_nemo_plainattn.json 2> gpurun_out/bench_nemo_plainattn.err
echo "bench nemo (plain attention) exit $?"; show gpurun_out/bench_nemo_plainattn.json
timeout 1200 python bench.py --model mixtral-8x7b --batch 8 --prefill 2048 --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/bench_mixtral_b8.json 2> gpurun_out/bench_mixtral_b8.err
echo "bench mixtral-8x7b B=8 exit $?"; show gpurun_out/bench_mixtral_b8.json; tail -3 gpurun_out/bench_mixtral_b8.err
timeout 1500 bash scripts/sanitize.sh memcheck racecheck synccheck
