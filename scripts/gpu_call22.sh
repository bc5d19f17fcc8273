#!/bin/bash
# round 2, call 22: A-box slack only where the stage is smaller than the MMA's 16 KB read -> 102 KB per stream-K CTA, two fit an SM.
# main library = new layout (tests + bench); libmb200_wideslack.so = the layout of calls 1-21 (bench only, same box)
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
timeout 200 python bench.py --model mistral-nemo-12b --batch 32 --prefill 1024 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_nemo_tight.json 2> gpurun_out/bench_nemo_tight.err
echo "nemo b32 tight slack exit $?"; show gpurun_out/bench_nemo_tight.json; tail -2 gpurun_out/bench_nemo_tight.err
MB200_LIB_PATH=$PWD/mistral_inference_b200/libmb200_wideslack.so timeout 200 python bench.py --model mistral-nemo-12b --batch 32 --prefill 1024 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_nemo_wide.json 2> gpurun_out/bench_nemo_wide.err
echo "nemo b32 wide slack exit $?"; show gpurun_out/bench_nemo_wide.json; tail -2 gpurun_out/bench_nemo_wide.err
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "gemm or moe or lm_head or linear or qkv" 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -3
