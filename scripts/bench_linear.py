#!/usr/bin/env python
"""Micro-benchmark of the decode-sized weight-streaming linears (stream-K tcgen05 GEMM) against a plain device copy of the same
bytes.  Weights rotate over enough copies to defeat the 126 MB L2.  Usage: python scripts/bench_linear.py [T]"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mistral_inference_b200 import _abi  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda")
ws = _abi.Workspace(_abi.workspace_bytes(128, 16384, 48, 8, 128, 16384, 0, 32), dev)
shapes = [("qkv 6144x5120", 6144, 5120), ("wo 5120x4096", 5120, 4096), ("gate/up 28672x5120", 28672, 5120), ("down 5120x14336", 5120, 14336),
          ("7B qkv 6144x4096", 6144, 4096), ("7B down 4096x14336", 4096, 14336), ("lm head 131072x5120", 131072, 5120)]


def timed(fn, n):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


for name, N, K in shapes:
    nbytes = N * K * 2
    R = max(2, (400 << 20) // nbytes + 1)
    w = [torch.randn(N, K, device=dev, dtype=torch.bfloat16) * K ** -0.5 for _ in range(R)]
    x = torch.randn(T, K, device=dev, dtype=torch.bfloat16)
    out = torch.empty(T, N, device=dev, dtype=torch.bfloat16)
    dst = torch.empty(N, K, device=dev, dtype=torch.bfloat16)
    iters = max(8, min(64, (4 << 30) // nbytes))
    t_gemm = timed(lambda i: _abi.linear_residual(x, w[i % R], None, out, ws), iters)
    t_copy = timed(lambda i: dst.copy_(w[i % R]), iters)
    # the same launches from a CUDA graph (no host launch cost between kernels)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(R):
            _abi.linear_residual(x, w[i], None, out, ws)
    t_graph = timed(lambda i: g.replay(), 8) / R
    print(f"T={T} {name:24s} {nbytes / 1e6:8.1f} MB  gemm {t_gemm:8.2f} us = {nbytes / t_gemm / 1e3:7.1f} GB/s   graph {t_graph:8.2f} us = {nbytes / t_graph / 1e3:7.1f} GB/s"
          f"   copy(r+w) {t_copy:8.2f} us = {2 * nbytes / t_copy / 1e3:7.1f} GB/s", flush=True)
    del w, dst
