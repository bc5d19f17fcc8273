#!/bin/bash
# One gpurun call: GPU parity tests, a bench line, the ncu launch list and (optionally) a full capture of the top kernel.
# usage: gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh [tests] [bench] [launches] [ncu:<kernel regex>]'
set -u
mkdir -p gpurun_out
STEPS="${*:-tests bench launches}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for s in $STEPS; do
  case "$s" in
    tests)
      timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 -s > gpurun_out/pytest.log 2>&1
      echo "pytest exit $?" >> gpurun_out/pytest.log
      grep -E "^\[parity\]|passed|failed|error|FAILED|ERROR" gpurun_out/pytest.log | tail -120
      ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
      ;;
    bench)
      timeout 900 python bench.py --steps 64 --warmup 8 > gpurun_out/bench.json 2> gpurun_out/bench.err
      echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
      ;;
    benchfast)
      timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
      echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
      ;;
    launches)
      MB200_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
        python bench.py --steps 4 --warmup 3 --prefill 4096 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
      echo "launches exit $?"; wc -l gpurun_out/launches.csv
      ;;
    ncu:*)
      pat="${s#ncu:}"
      MB200_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:${pat}" -s "${NCU_SKIP:-0}" -c "${NCU_COUNT:-4}" -f -o "gpurun_out/prof_${pat//[^a-zA-Z0-9_]/_}" \
        python bench.py --steps 2 --warmup 3 ${NCU_BENCH_ARGS:---prefill 4096 --layers 2} --no-cpu-baseline > "gpurun_out/ncu_${pat//[^a-zA-Z0-9_]/_}.log" 2>&1
      echo "ncu $pat exit $?"
      ;;
  esac
done
