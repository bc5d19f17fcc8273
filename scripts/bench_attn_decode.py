#!/usr/bin/env python
"""Decode attention in a chain (CUDA graph over distinct KV rings, as in a decode step): microseconds per launch for several split
counts.  Usage: python scripts/bench_attn_decode.py [B] [kv_len] [KV] [H]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mistral_inference_b200 import _abi  # noqa: E402
from mistral_inference_b200.transformer_layers import decode_splits  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
kv_len = int(sys.argv[2]) if len(sys.argv) > 2 else 1106
KV = int(sys.argv[3]) if len(sys.argv) > 3 else 8
H = int(sys.argv[4]) if len(sys.argv) > 4 else 32
hd, L = 128, 24
W = ((kv_len + 255) // 256) * 256
dev = torch.device("cuda")
ws = _abi.Workspace(_abi.workspace_bytes(128, 5120, H, KV, hd, 14336, 0, B), dev)
rings = [(torch.randn(B, W, KV, hd, device=dev, dtype=torch.bfloat16), torch.randn(B, W, KV, hd, device=dev, dtype=torch.bfloat16)) for _ in range(L)]
q = torch.randn(B, H * hd, device=dev, dtype=torch.bfloat16)
out = torch.empty_like(q)
lens = torch.full((B,), kv_len, dtype=torch.int32, device=dev)
nbytes = 2 * B * kv_len * KV * hd * 2
print(f"B={B} kv_len={kv_len} KV={KV} H={H}: {nbytes / 1e6:.1f} MB of keys and values per launch; default splits {decode_splits(B, KV, W)}")
for splits in (1, 2, 3, 4, 6, 8):
    def step():
        for ck, cv in rings:
            _abi.attn_decode(q, ck, cv, lens, out, H, KV, hd, splits, ws)
    step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    for _ in range(2):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (5 * L)
    print(f"  splits={splits}: {us:7.2f} us = {nbytes / us / 1e3:7.1f} GB/s", flush=True)
