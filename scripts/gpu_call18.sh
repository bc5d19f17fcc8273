#!/bin/bash
# round 2, call 18: time line of dependent stream-K launches (tracing build); decode attention split sweep in a chain
set -u
mkdir -p gpurun_out
timeout 300 python scripts/trace_streamk.py 32 2>&1 | tee gpurun_out/trace_streamk_t32.txt
timeout 300 python scripts/bench_attn_decode.py 32 1106 8 32 2>&1 | tee gpurun_out/bench_attn_decode_b32.txt
timeout 300 python scripts/bench_attn_decode.py 8 2100 8 32 2>&1 | tee gpurun_out/bench_attn_decode_b8.txt
