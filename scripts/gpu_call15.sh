#!/bin/bash
# round 2, call 15: L2 prefetch of the next weight tiles across dependent stream-K launches (MB200_SK_PREFETCH boxes per CTA)
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
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "gemm or ffn_block or lm_head or linear" 2>&1 | tail -4
for pf in 0 16 48; do
  echo "== MB200_SK_PREFETCH=$pf"
  MB200_SK_PREFETCH=$pf timeout 300 python scripts/bench_linear.py 32 2>&1 | sed 's/   copy.*//' | tee gpurun_out/bench_linear_pf$pf.txt
done
for pf in 16 0 48; do
  MB200_SK_PREFETCH=$pf timeout 600 python bench.py --model mistral-nemo-12b --batch 32 --prefill 1024 --steps 30 --warmup 5 --no-cpu-baseline --no-parity > gpurun_out/bench_nemo_pf$pf.json 2> gpurun_out/bench_nemo_pf$pf.err
  echo "nemo b32 prefetch=$pf exit $?"; show gpurun_out/bench_nemo_pf$pf.json; tail -2 gpurun_out/bench_nemo_pf$pf.err
done
