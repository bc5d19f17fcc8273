#!/bin/bash
# round 2, call 11 (8 GPUs): BASELINE configs[4]: Mixtral-8x22B shape expert-parallel over 8 GPUs, batch 16, 4k prefill
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/gpu8.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --parallel expert \
  --model mixtral-8x22b --batch 16 --prefill 4096 --steps 32 --warmup 4 > gpurun_out/bench_ep8_8x22b.json 2> gpurun_out/bench_ep8_8x22b.err
echo "bench EP8 mixtral-8x22b exit $?"; python - <<'PY'
import json
try:
    txt = open("gpurun_out/bench_ep8_8x22b.json").read()
    d = json.loads(txt[txt.index('{"metric'):])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["roofline"], d["prefill"], d["e2e"]["value"])
except Exception as e:
    print("parse failed", e)
PY
tail -8 gpurun_out/bench_ep8_8x22b.err
