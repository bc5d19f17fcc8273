#!/bin/bash
# round 2, call 12: 2-D blocked tile order for large-T GEMMs: tests, Nemo / Mixtral prefill, 7B default line, ncu of the grouped GEMM
set -u
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print({k: d[k] for k in ("value", "ms_per_step")}, "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "prefill", d["prefill"]["ms"], d["prefill"]["tflops"], d["clocks"])
except Exception as e:
    print("parse failed", e)
PY
}
for f in test_gpu_ops test_gpu_model; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q -p no:cacheprovider --timeout 900 -s > gpurun_out/pytest_$f.log 2>&1
  echo "pytest $f exit $?"
  grep -E "passed|failed|FAILED|ERROR|watchdog" gpurun_out/pytest_$f.log | sort | uniq -c | sort -rn | head -20
done
timeout 900 python bench.py --model mistral-nemo-12b --batch 32 --prefill 1024 --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/bench_nemo.json 2> gpurun_out/bench_nemo.err
echo "bench nemo exit $?"; show gpurun_out/bench_nemo.json; tail -3 gpurun_out/bench_nemo.err
timeout 1200 python bench.py --model mixtral-8x7b --batch 8 --prefill 2048 --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/bench_mixtral_b8.json 2> gpurun_out/bench_mixtral_b8.err
echo "bench mixtral-8x7b B=8 exit $?"; show gpurun_out/bench_mixtral_b8.json; tail -3 gpurun_out/bench_mixtral_b8.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default exit $?"; show gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "bench reference exit $?"; head -c 1500 gpurun_out/bench_reference.json; echo; tail -3 gpurun_out/bench_reference.err
MB200_PROFILE=1 MB200_DECODE_GRAPH=0 timeout 900 ncu --profile-from-start off --set full --clock-control none -k "regex:gemm_tcgen05_grouped_kernel" -s 2 -c 2 -f -o /tmp/prof_grouped \
  python bench.py --model mixtral-8x7b --batch 8 --prefill 2048 --layers 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_grouped.log 2>&1
python scripts/ncu_summary.py /tmp/prof_grouped.ncu-rep "round 2: ncu --set full -k regex:gemm_tcgen05_grouped_kernel (Mixtral-8x7B shapes, 8 x 2048-token prefill: grouped gate/up and down GEMMs over 8 experts, 2-CTA cluster pairs, 2-D blocked tile order)" > gpurun_out/r02_ncu_grouped.txt 2>> gpurun_out/ncu_grouped.log
rm -f /tmp/prof_grouped.ncu-rep
grep -E "Kernel Name|gpu__time_duration|dram__bytes_read.sum |pipe_tensor" gpurun_out/r02_ncu_grouped.txt | head -8
MB200_PROFILE=1 MB200_DECODE_GRAPH=0 timeout 900 ncu --profile-from-start off --set full --clock-control none -k "regex:gemm_tcgen05_kernel" -s 6 -c 3 -f -o /tmp/prof_dense \
  python bench.py --model mistral-nemo-12b --batch 32 --prefill 1024 --layers 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_dense.log 2>&1
python scripts/ncu_summary.py /tmp/prof_dense.ncu-rep "round 2: ncu --set full -k regex:gemm_tcgen05_kernel (Nemo-12B shapes, 32 x 1024-token prefill: dense linears at T = 32768, 2-CTA clusters, 2-D blocked tile order)" > gpurun_out/r02_ncu_gemm_tcgen05_T32768.txt 2>> gpurun_out/ncu_dense.log
rm -f /tmp/prof_dense.ncu-rep
grep -E "Kernel Name|gpu__time_duration|dram__bytes_read.sum |pipe_tensor" gpurun_out/r02_ncu_gemm_tcgen05_T32768.txt | head -12
