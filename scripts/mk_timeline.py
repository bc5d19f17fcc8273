"""Debug: per-phase timeline of the decode megakernel (CTA 0), Mistral-7B shape at kv_len 4096.
usage (GPU box): python scripts/mk_timeline.py [n_layers]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from mistral_inference_b200 import _abi, synth  # noqa: E402
from mistral_inference_b200.cache import BufferCache  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
p = synth.shape("mistral-7b", n_layers=L)
model = bench.build_gpu_model(p, 1)
cache = BufferCache(L, 1, 4096 + 64, p["n_kv_heads"], p["head_dim"], p["sliding_window"]).to(model.device, model.dtype)
for i in cache.cache_k:
    cache.cache_k[i].normal_()
    cache.cache_v[i].normal_()
cache._kv_seqlens_host = [5000]
tok = torch.tensor([17], device="cuda")
for _ in range(3):
    model.decode_static(tok, cache)
buf = torch.zeros(L * 12, dtype=torch.int64, device="cuda")
_abi.set_decode_timeline(buf)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
model.decode_static(tok, cache)
e1.record()
torch.cuda.synchronize()
_abi.set_decode_timeline(None)
t = buf.cpu().view(L, 12).double()
names = ["stage_x+norm", "QKV gemv", "barrier1", "attention", "barrier2", "stage+WO gemv", "barrier3", "stage+GATEUP", "barrier4",
         "stage+DOWN", "barrier5"]
d = (t[:, 1:] - t[:, :-1]) / 1000.0  # us
print(f"kernel total {e0.elapsed_time(e1) * 1000:.1f} us for {L} layers; per-layer phase durations of CTA 0 (us), median over layers 1..:")
med = d[1:].median(0).values
for n, v in zip(names, med.tolist()):
    print(f"  {n:16s} {v:8.2f}")
print(f"  {'layer total':16s} {((t[1:, 11] - t[1:, 0]) / 1000).median().item():8.2f}   (ideal HBM time per layer at 6583 GB/s: {(436.2e6 + 16.8e6) / 6583.5e3:.1f} us)")
