#!/bin/bash
# round 2, call 10: ncu --set full captures of the round-2 kernels, summarised ON THE BOX (the .ncu-rep files exceed gpurun's 64 MiB return limit),
# Mixtral-8x7B B=8 bench + launch list with the cluster grouped GEMM, Nemo B=32 with one-wave attention splits
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
cap() {  # cap <name> <kernel regex> <skip> <count> <header> -- <bench args>
  local name=$1 pat=$2 skip=$3 cnt=$4 hdr=$5; shift 6
  MB200_PROFILE=1 MB200_DECODE_GRAPH=0 timeout 900 ncu --profile-from-start off --set full --clock-control none -k "regex:$pat" -s $skip -c $cnt -f -o /tmp/prof_$name \
    python bench.py "$@" --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name exit $?"
  python scripts/ncu_summary.py /tmp/prof_$name.ncu-rep "$hdr" > gpurun_out/r02_ncu_$name.txt 2>> gpurun_out/ncu_$name.log
  rm -f /tmp/prof_$name.ncu-rep
  grep -E "Kernel Name|gpu__time_duration|dram__bytes_read.sum |dram__bytes_read.sum.per_second|pipe_tensor" gpurun_out/r02_ncu_$name.txt | head -12
}
timeout 900 python bench.py --model mistral-nemo-12b --batch 32 --prefill 1024 --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/bench_nemo.json 2> gpurun_out/bench_nemo.err
echo "bench nemo exit $?"; show gpurun_out/bench_nemo.json; tail -3 gpurun_out/bench_nemo.err
timeout 1200 python bench.py --model mixtral-8x7b --batch 8 --prefill 2048 --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/bench_mixtral_b8.json 2> gpurun_out/bench_mixtral_b8.err
echo "bench mixtral-8x7b B=8 exit $?"; show gpurun_out/bench_mixtral_b8.json; tail -3 gpurun_out/bench_mixtral_b8.err
MB200_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_mixtral.csv \
  python bench.py --model mixtral-8x7b --batch 8 --prefill 2048 --layers 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_mixtral.log 2>&1
echo "launches mixtral exit $?"; python scripts/launches_summary.py gpurun_out/launches_mixtral.csv 2>/dev/null | grep -v "^#" | grep -v "at::" | head -14
cap streamk "gemm_streamk_kernel" 24 3 "round 2: ncu --set full --clock-control none -k regex:gemm_streamk_kernel (python bench.py --model mistral-nemo-12b --batch 32 --prefill 1024 --layers 2; decode-step linears at T = 32, eager launches)" -- --model mistral-nemo-12b --batch 32 --prefill 1024 --layers 2
cap attn_decode_tma "attn_decode_tma_kernel" 4 1 "round 2: ncu --set full -k regex:attn_decode_tma_kernel (Nemo-12B shapes, batch 32, kv_len ~1030: 135 MB of K/V per launch)" -- --model mistral-nemo-12b --batch 32 --prefill 1024 --layers 2
cap grouped "gemm_tcgen05_grouped_kernel" 2 2 "round 2: ncu --set full -k regex:gemm_tcgen05_grouped_kernel (Mixtral-8x7B shapes, 8 x 2048-token prefill: grouped gate/up and down GEMMs over 8 experts, 2-CTA cluster pairs)" -- --model mixtral-8x7b --batch 8 --prefill 2048 --layers 2
cap streamk_grouped "gemm_streamk_grouped_kernel" 8 2 "round 2: ncu --set full -k regex:gemm_streamk_grouped_kernel (Mixtral-8x7B shapes, batch-8 decode step: ~7 experts x (gate/up 235 MB | down 117 MB) streamed once)" -- --model mixtral-8x7b --batch 8 --prefill 2048 --layers 2
du -sh gpurun_out
