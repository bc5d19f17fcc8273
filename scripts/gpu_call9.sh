#!/bin/bash
# round 2, call 9: grouped GEMM with 2-CTA cluster pairs (tests, Mixtral-8x7B B=8 prefill), ncu --set full captures of the round-2 kernels
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
  grep -E "passed|failed|FAILED|ERROR|watchdog" gpurun_out/pytest_$f.log | sort | uniq -c | sort -rn | head -20
done
timeout 1200 python bench.py --model mixtral-8x7b --batch 8 --prefill 2048 --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/bench_mixtral_b8_cl2.json 2> gpurun_out/bench_mixtral_b8_cl2.err
echo "bench mixtral-8x7b B=8 (cluster grouped GEMM) exit $?"; show gpurun_out/bench_mixtral_b8_cl2.json; tail -3 gpurun_out/bench_mixtral_b8_cl2.err
# ---- ncu --set full: stream-K GEMM (gate/up of a Nemo decode step), TMA decode attention, grouped cluster GEMM (Mixtral prefill), grouped stream-K (Mixtral decode)
MB200_PROFILE=1 MB200_DECODE_GRAPH=0 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_streamk_kernel -s 20 -c 4 -f -o gpurun_out/prof_r02_streamk \
  python bench.py --model mistral-nemo-12b --batch 32 --prefill 1024 --layers 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_streamk.log 2>&1; echo "ncu streamk exit $?"
MB200_PROFILE=1 MB200_DECODE_GRAPH=0 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_decode_tma_kernel -s 4 -c 2 -f -o gpurun_out/prof_r02_attn_decode_tma \
  python bench.py --model mistral-nemo-12b --batch 32 --prefill 1024 --layers 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?"
MB200_PROFILE=1 MB200_DECODE_GRAPH=0 timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tcgen05_grouped_kernel -s 2 -c 2 -f -o gpurun_out/prof_r02_grouped \
  python bench.py --model mixtral-8x7b --batch 8 --prefill 2048 --layers 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_grouped.log 2>&1; echo "ncu grouped exit $?"
MB200_PROFILE=1 MB200_DECODE_GRAPH=0 timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_streamk_grouped_kernel -s 8 -c 2 -f -o gpurun_out/prof_r02_streamk_grouped \
  python bench.py --model mixtral-8x7b --batch 8 --prefill 2048 --layers 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_streamk_grouped.log 2>&1; echo "ncu streamk grouped exit $?"
ls -la gpurun_out/*.ncu-rep
