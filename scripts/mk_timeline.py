"""Debug: per-phase timeline of the decode megakernel (CTA 0), Mistral-7B shape at kv_len 4096.
usage (GPU box): python scripts/mk_timeline.py [n_layers]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
import synth  # noqa: E402
from mistral_inference_b200 import _abi  # noqa: E402
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
buf = torch.zeros(8 * L * 16, dtype=torch.int64, device="cuda")
_abi.set_decode_timeline(buf)
sm_count = _abi.device_info()[0]
bbuf = torch.zeros(sm_count * L * 6 * 2, dtype=torch.int64, device="cuda")
_abi.set_barrier_timeline(bbuf)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
model.decode_static(tok, cache)
e1.record()
torch.cuda.synchronize()
_abi.set_decode_timeline(None)
_abi.set_barrier_timeline(None)
t = buf.cpu().view(8, L, 16).double() / 1000.0  # us; 8 sampled CTAs (0, 21, ..., 147)
names = ["stage_x+norm", "QKV gemv", "barrier1", "attention", "barrier2", "stage+WO gemv", "barrier3", "stage+GATEUP", "barrier4",
         "stage+DOWN", "barrier5"]
d = t[:, :, 1:12] - t[:, :, :11]
print(f"kernel total {e0.elapsed_time(e1) * 1000:.1f} us for {L} layers; phase durations (us), median over layers 1.., per sampled CTA min / CTA 0 / max:")
med = d[:, 1:].median(1).values  # [8, 11]
for i, n in enumerate(names):
    print(f"  {n:16s} min {med[:, i].min():7.2f}   cta0 {med[0, i]:7.2f}   max {med[:, i].max():7.2f}")
arr = t[:, 1:, [2, 4, 6, 8, 10]]  # arrival times at the 5 barriers
lea = t[:, 1:, [3, 5, 7, 9, 11]]  # leave times
skew = (arr.max(0).values - arr.min(0).values).median(0).values
lat = (lea.min(0).values - arr.max(0).values).median(0).values
a = t[:6, 1:]  # active sampled CTAs
sub = torch.stack([a[:, :, 14] - a[:, :, 3], a[:, :, 15] - a[:, :, 3], a[:, :, 12] - a[:, :, 3], a[:, :, 13] - a[:, :, 12], a[:, :, 4] - a[:, :, 13]], -1).median(1).values
print("  inside attention (us; first KV stage ready | last KV stage ready | slice partial done (all since phase start) | mid barrier | slice merge):")
for row in sub.tolist():
    print("     ", [round(x, 2) for x in row])
print("  barrier arrival skew across sampled CTAs (us):", [round(x, 2) for x in skew.tolist()])
print("  barrier latency last-arrival -> first-leave (us):", [round(x, 2) for x in lat.tolist()])
print(f"  layer total (cta0) {(t[0, 1:, 11] - t[0, 1:, 0]).median().item():8.2f}   (ideal HBM time per layer at 6583 GB/s: {(436.2e6 + 16.8e6) / 6583.5e3:.1f} us)")

bt = bbuf.cpu().view(sm_count, L, 6, 2).double() / 1000.0
arr, lea = bt[:, 1:, :, 0], bt[:, 1:, :, 1]
print("  ALL CTAs, per barrier (median over layers): arrival skew (last - first arrive) | latency (first leave - last arrive) | leave spread")
for b, nm in enumerate(["after QKV", "after slice partial", "after slice merge", "after WO", "after GATEUP", "after DOWN"]):
    skew = (arr[:, :, b].max(0).values - arr[:, :, b].min(0).values).median().item()
    lat = (lea[:, :, b].min(0).values - arr[:, :, b].max(0).values).median().item()
    spread = (lea[:, :, b].max(0).values - lea[:, :, b].min(0).values).median().item()
    slow = arr[:, :, b].argmax(0).mode().values.item()
    print(f"    {nm:20s} skew {skew:6.2f}   latency {lat:6.2f}   leave spread {spread:6.2f}   (most often last: CTA {slow})")

# per-CTA arrival pattern at the gate/up barrier (index 4) and the wo barrier (3), layer 3
for b, nm in [(4, "after GATEUP"), (3, "after WO"), (5, "after DOWN")]:
    arrv = bt[:, 3, b, 0]
    rel = arrv - arrv.min()
    order = rel.argsort()
    print(f"  {nm}: arrival offsets (us) percentiles 10/50/90/100: {rel.quantile(0.1):.2f} {rel.quantile(0.5):.2f} {rel.quantile(0.9):.2f} {rel.max():.2f};"
          f" earliest CTAs {order[:8].tolist()} latest CTAs {order[-12:].tolist()}")
    lv = bt[:, 3, b, 1]
    print(f"      leave - last arrive per CTA: min {(lv - arrv.max()).min():.2f} median {(lv - arrv.max()).median():.2f} max {(lv - arrv.max()).max():.2f}")
