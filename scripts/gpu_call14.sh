#!/bin/bash
# round 2, call 14: persistent ffn_block kernel -- op test, model tests, Nemo B=32 bench with and without it
set -u
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
try:
    txt = open(sys.argv[1]).read()
    d = json.loads(txt[txt.index('{"metric'):])
    print({k: d[k] for k in ("value", "ms_per_step")}, "frac", d["roofline"]["frac"], "us", d["roofline"]["us_per_launch"], "e2e", d["e2e"]["value"], "prefill", d["prefill"]["ms"], d["parity"] and d["parity"]["ok"], d["clocks"])
except Exception as e:
    print("parse failed", e)
PY
}
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "ffn_block" 2>&1 | tail -15
echo "ffn_block op test exit $?"
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -8
for fb in 1 0; do
  MB200_FFN_BLOCK=$fb timeout 600 python bench.py --model mistral-nemo-12b --batch 32 --prefill 1024 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_nemo_fb$fb.json 2> gpurun_out/bench_nemo_fb$fb.err
  echo "nemo b32 ffn_block=$fb exit $?"; show gpurun_out/bench_nemo_fb$fb.json; tail -3 gpurun_out/bench_nemo_fb$fb.err
done
MB200_FFN_BLOCK=1 timeout 600 python bench.py --model mistral-7b --batch 16 --prefill 1024 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_7b_b16_fb1.json 2> gpurun_out/bench_7b_b16_fb1.err
echo "7b b16 fb=1 exit $?"; show gpurun_out/bench_7b_b16_fb1.json
MB200_FFN_BLOCK=0 timeout 600 python bench.py --model mistral-7b --batch 16 --prefill 1024 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_7b_b16_fb0.json 2> gpurun_out/bench_7b_b16_fb0.err
echo "7b b16 fb=0 exit $?"; show gpurun_out/bench_7b_b16_fb0.json
