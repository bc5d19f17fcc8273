#!/bin/bash
# round 2, call 21: evidence after the stream-K reduction change -- Nemo-12B batch 32 launch list (4-layer slice) and an
# ncu --set full capture of the stream-K GEMM, both summarised on the box
set -u
mkdir -p gpurun_out
MB200_PROFILE=1 timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_nemo.csv \
  python bench.py --model mistral-nemo-12b --batch 32 --prefill 1024 --layers 4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_nemo.log 2>&1
echo "launches exit $?"
python scripts/launches_summary.py gpurun_out/launches_nemo.csv "round 2 (call 21, after the parallel split-tile reduction): python bench.py --model mistral-nemo-12b --batch 32 --prefill 1024 --layers 4 --steps 2 --warmup 3 (4-layer slice of BASELINE configs[2])" > gpurun_out/r02_nemo_b32_launches_after.txt 2>/dev/null
grep -v "at::" gpurun_out/r02_nemo_b32_launches_after.txt | head -16
rm -f gpurun_out/launches_nemo.csv
MB200_PROFILE=1 MB200_DECODE_GRAPH=0 timeout 500 ncu --profile-from-start off --set full --clock-control none -k "regex:gemm_streamk_kernel" -s 24 -c 4 -f -o /tmp/prof_streamk \
  python bench.py --model mistral-nemo-12b --batch 32 --prefill 1024 --layers 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_streamk.log 2>&1
echo "ncu streamk exit $?"
python scripts/ncu_summary.py /tmp/prof_streamk.ncu-rep "round 2 (call 21): ncu --set full --clock-control none -k regex:gemm_streamk_kernel -s 24 -c 4 (python bench.py --model mistral-nemo-12b --batch 32 --prefill 1024 --layers 2; decode-step linears at T = 32, eager launches, after the parallel split-tile reduction)" > gpurun_out/r02_ncu_streamk_after.txt 2>> gpurun_out/ncu_streamk.log
rm -f /tmp/prof_streamk.ncu-rep
grep -E "Kernel Name|gpu__time_duration|dram__bytes_read.sum |dram__bytes_read.sum.per_second|dram__bytes_write.sum " gpurun_out/r02_ncu_streamk_after.txt | head -20
