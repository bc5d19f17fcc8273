#!/bin/bash
# round 2, call 4 (2 GPUs): expert-parallel tests over NVLink, config 4 (Mixtral-8x7B expert-sharded over 2 GPUs, B=8, 2k prefill), N=2 default line
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/gpu2.txt 2>&1
nvidia-smi topo -m >> gpurun_out/gpu2.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_expert_parallel.py -m gpu -q -p no:cacheprovider --timeout 600 -s > gpurun_out/pytest_ep2.log 2>&1
echo "pytest EP (2 GPUs) exit $?"; grep -E "passed|failed|FAILED|ERROR|watchdog|skipped" gpurun_out/pytest_ep2.log | tail -8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --parallel expert \
  --model mixtral-8x7b --batch 8 --prefill 2048 --steps 32 --warmup 4 > gpurun_out/bench_ep2_mixtral.json 2> gpurun_out/bench_ep2_mixtral.err
echo "bench EP mixtral-8x7b exit $?"; head -c 2500 gpurun_out/bench_ep2_mixtral.json; echo; tail -5 gpurun_out/bench_ep2_mixtral.err
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 \
  > gpurun_out/bench_n2_default.json 2> gpurun_out/bench_n2_default.err
echo "bench N=2 default exit $?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_n2_default.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["roofline"]["frac"], "sharded:", d.get("sharded"))
except Exception as e:
    print("parse failed", e)
PY
tail -5 gpurun_out/bench_n2_default.err
