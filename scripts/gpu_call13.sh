#!/bin/bash
# round 2, call 13: sanity of the bench flow (e2e loop first) on one GPU before the 8-GPU run
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
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default exit $?"; show gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout 900 python bench.py --model mixtral-8x7b --batch 8 --prefill 512 --layers 4 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mixtral_quick.json 2> gpurun_out/bench_mixtral_quick.err
echo "bench mixtral quick exit $?"; show gpurun_out/bench_mixtral_quick.json; tail -3 gpurun_out/bench_mixtral_quick.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
