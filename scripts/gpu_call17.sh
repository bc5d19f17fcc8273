#!/bin/bash
# round 2, call 17: stream-K owner sums the contributors' slots while its own last k-blocks are still in the tensor core
set -u
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
try:
    txt = open(sys.argv[1]).read()
    d = json.loads(txt[txt.index('{"metric'):])
    print({k: d[k] for k in ("value", "ms_per_step")}, "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "prefill", d["prefill"]["ms"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("parse failed", e)
PY
}
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "gemm or moe or lm_head or linear" 2>&1 | tail -3
timeout 300 python scripts/bench_linear.py 32 2>&1 | sed 's/   copy.*//' | tee gpurun_out/bench_linear_t32.txt
timeout 600 python bench.py --model mistral-nemo-12b --batch 32 --prefill 1024 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_nemo_b32.json 2> gpurun_out/bench_nemo_b32.err
echo "nemo b32 exit $?"; show gpurun_out/bench_nemo_b32.json; tail -2 gpurun_out/bench_nemo_b32.err
