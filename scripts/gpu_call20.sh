#!/bin/bash
# round 2, call 20: stream-K reduction with parallel flag polling + one round trip for <= 4 contributors; full GPU suite, smoke, benches
set -u
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
try:
    txt = open(sys.argv[1]).read()
    d = json.loads(txt[txt.index('{"metric'):])
    print({k: d[k] for k in ("value", "ms_per_step")}, "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "prefill", d["prefill"]["ms"], d["parity"] and d["parity"]["ok"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("parse failed", e)
PY
}
timeout 300 python scripts/trace_streamk.py 32 2>&1 | tee gpurun_out/trace_streamk_t32.txt | grep -A9 "== wo\|per layer"
timeout 600 python bench.py --model mistral-nemo-12b --batch 32 --prefill 1024 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_nemo_b32.json 2> gpurun_out/bench_nemo_b32.err
echo "nemo b32 exit $?"; show gpurun_out/bench_nemo_b32.json; tail -2 gpurun_out/bench_nemo_b32.err
timeout 1800 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_final.log 2>&1; echo "pytest -m gpu exit $?"; tail -4 gpurun_out/pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "our arm exit $?"; show gpurun_out/bench_default.json
timeout 900 python bench.py --model mixtral-8x7b --batch 8 --prefill 2048 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_mixtral_b8.json 2> gpurun_out/bench_mixtral_b8.err
echo "mixtral b8 exit $?"; show gpurun_out/bench_mixtral_b8.json; tail -2 gpurun_out/bench_mixtral_b8.err
