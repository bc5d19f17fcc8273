#!/bin/bash
# round 2, call 8: megakernel with non-blocking register prefetch across grid barriers (libmb200_mkpf.so) vs the round-1 kernel: parity tests, then A/B
set -u
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print({k: d[k] for k in ("value", "ms_per_step")}, "frac", d["roofline"]["frac"], "us", d["roofline"]["us_per_launch"], "e2e", d["e2e"]["value"], "parity", d.get("parity"), d["clocks"])
except Exception as e:
    print("parse failed", e)
PY
}
MB200_LIB_PATH=$PWD/mistral_inference_b200/libmb200_mkpf.so timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -p no:cacheprovider --timeout 600 -s -k "megakernel or config1 or golden" > gpurun_out/pytest_mkpf.log 2>&1
echo "pytest (prefetch build) exit $?"; grep -E "passed|failed|FAILED|ERROR|watchdog" gpurun_out/pytest_mkpf.log | sort | uniq -c | sort -rn | head
for lib in mkpf "" mkpf ""; do
  if [ -n "$lib" ]; then export MB200_LIB_PATH=$PWD/mistral_inference_b200/libmb200_$lib.so; else unset MB200_LIB_PATH; fi
  timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/bench_ab_${lib:-r1}.json 2> gpurun_out/bench_ab_${lib:-r1}.err
  echo "bench ${lib:-r1} exit $?"; show gpurun_out/bench_ab_${lib:-r1}.json; tail -2 gpurun_out/bench_ab_${lib:-r1}.err
done
unset MB200_LIB_PATH
timeout 900 python bench.py --model mixtral-8x7b --steps 64 --warmup 8 --no-cpu-baseline --no-parity > gpurun_out/bench_mixtral_b1_r1.json 2> gpurun_out/bench_mixtral_b1_r1.err
echo "bench mixtral B=1 r1 exit $?"; show gpurun_out/bench_mixtral_b1_r1.json
MB200_LIB_PATH=$PWD/mistral_inference_b200/libmb200_mkpf.so timeout 900 python bench.py --model mixtral-8x7b --steps 64 --warmup 8 --no-cpu-baseline --no-parity > gpurun_out/bench_mixtral_b1_pf.json 2> gpurun_out/bench_mixtral_b1_pf.err
echo "bench mixtral B=1 prefetch exit $?"; show gpurun_out/bench_mixtral_b1_pf.json
