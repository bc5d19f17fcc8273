#!/bin/bash
# what the driver runs at round end, on one box: the GPU suite in one process, smoke(), both bench arms with its flags
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_final.log 2>&1; echo "pytest -m gpu exit $?"; tail -4 gpurun_out/pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm exit $?"; head -c 400 gpurun_out/bench_reference.json; echo
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "our arm exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["roofline"]["frac"], d["e2e"], d["clocks"], d["parity"]["ok"], d["cpu_baseline"]["value"])
PY
tail -3 gpurun_out/bench_default.err
