#!/usr/bin/env python
"""Where do the microseconds between dependent weight-streaming launches go?  Runs the four stream-K GEMMs of a Nemo-12B layer
(qkv, wo, [rmsnorm] gate/up, down) at T tokens over several layers from a CUDA graph with the tracing build of the library
(-DMB200_SK_TRACE: every CTA stamps %globaltimer at eight points) and prints, per GEMM, the time line relative to the moment the
previous launch's last CTA exited.  Usage: python scripts/trace_streamk.py [T]   (builds libmb200_sktrace.so if it is missing)"""
import ctypes
import os
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
lib_path = REPO / "mistral_inference_b200" / "libmb200_sktrace.so"
os.environ["MB200_LIB_PATH"] = str(lib_path)  # read when the package is first imported
from mistral_inference_b200.build import build_variant  # noqa: E402

if not lib_path.exists():
    build_variant("sktrace", ["MB200_SK_TRACE"])

import numpy as np  # noqa: E402
import torch  # noqa: E402

from mistral_inference_b200 import _abi  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda")
dim, q_dim, qkv_n, hidden, L = 5120, 4096, 6144, 14336, 8
ws = _abi.Workspace(_abi.workspace_bytes(128, dim, 32, 8, 128, hidden, 0, 32), dev)
mk = lambda n, k: torch.randn(n, k, device=dev, dtype=torch.bfloat16) * k ** -0.5
layers = [dict(wqkv=mk(qkv_n, dim), wo=mk(dim, q_dim), w13=mk(2 * hidden, dim), w2=mk(dim, hidden)) for _ in range(L)]
nw = torch.ones(dim, device=dev, dtype=torch.bfloat16)
x = torch.randn(T, dim, device=dev, dtype=torch.bfloat16)
qkv = torch.empty(T, qkv_n, device=dev, dtype=torch.bfloat16)
a = torch.randn(T, q_dim, device=dev, dtype=torch.bfloat16)
h = torch.empty(T, dim, device=dev, dtype=torch.bfloat16)
g = torch.empty(T, hidden, device=dev, dtype=torch.bfloat16)
out = torch.empty(T, dim, device=dev, dtype=torch.bfloat16)


def step():
    for w in layers:
        _abi.linear_residual(x, w["wqkv"], None, qkv, ws)
        _abi.linear_residual(a, w["wo"], x, h, ws)
        _abi.ffn_gateup(h, nw, w["w13"], g, 1e-5, ws)  # rmsnorm kernel + GEMM
        _abi.linear_residual(g, w["w2"], h, out, ws)


for _ in range(2):
    step()
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    step()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    graph.replay()
e0.record()
for _ in range(4):
    graph.replay()
e1.record()
torch.cuda.synchronize()
print(f"T={T}: {e0.elapsed_time(e1) * 1e3 / (4 * L):.2f} us per layer (4 GEMMs + 1 rmsnorm), graph replay")

NL, NC, NP = 64, 160, 8
buf = np.zeros((NL, NC, NP, 2), dtype=np.uint64)
count = ctypes.c_uint(0)
fn = _abi.lib().mb200_debug_sk_trace
fn.restype = ctypes.c_int
rc = fn(ctypes.c_void_p(buf.ctypes.data), ctypes.c_size_t(buf.nbytes), ctypes.byref(count))
assert rc == 0, rc
G = torch.cuda.get_device_properties(0).multi_processor_count
n_launch = count.value // G
names = ["qkv", "wo", "gate/up", "down"]
points = ["set up", "dep. wait returned", "first MMA", "last tile requested", "last MMA issued", "last accumulator ready", "epilogues done", "exit"]
rows = {n: [] for n in names}
for li in range(n_launch - NL + 2, n_launch):  # launches still in the ring, with their predecessor
    cur, prev = buf[li % NL, :G, :, 0].astype(np.int64), buf[(li - 1) % NL, :G, :, 0].astype(np.int64)
    origin = prev[:, 7].max()
    rows[names[li % 4]].append((cur - origin) / 1e3)  # us
for n in names:
    r = np.stack(rows[n])  # [launches, G, points]
    print(f"\n== {n}: {len(rows[n])} launches; us relative to the exit of the previous launch's last CTA  (min / median / max over CTAs, mean over launches)")
    for pi, pn in enumerate(points):
        v = r[:, :, pi]
        print(f"  {pn:24s} {v.min(1).mean():8.2f} {np.median(v, 1).mean():8.2f} {v.max(1).mean():8.2f}")
    clk = buf[(n_launch - 1 - (3 - names.index(n))) % NL, :G, :, 1].astype(np.int64)
    print(f"  (clock64, one launch, median over CTAs: last accumulator -> epilogues done {np.median(clk[:, 6] - clk[:, 5]) / 1.965e3:.2f} us, "
          f"dep. wait -> first MMA {np.median(clk[:, 2] - clk[:, 1]) / 1.965e3:.2f} us -- stamps 1 and 2 come from different warps)")
