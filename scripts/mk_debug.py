"""Debug driver for the decode megakernel: tiny + 7B-shape models, prints progress unbuffered."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mistral_inference_b200 as mi  # noqa: E402
import synth  # noqa: E402
from mistral_inference_b200.cache import BufferCache  # noqa: E402
from mistral_inference_b200.transformer import Transformer  # noqa: E402


def log(*a):
    print(*a, flush=True)


def model_for(p, max_batch=1):
    args = mi.TransformerArgs.from_dict(dict(p))
    args.max_batch_size = max_batch
    with torch.device("cuda"):
        m = Transformer(args).to(torch.bfloat16)
    with torch.no_grad():
        for k, shp in synth.state_dict_shapes(p):
            m._assign(k, synth.synth_tensor(k, shp, 1, torch.bfloat16, "cuda"))
    return m.eval()


for name, over, prompt_len, steps in [("tiny", {}, 9, 4), ("tiny", {"sliding_window": 6}, 9, 8),
                                      ("mistral-7b", {"n_layers": 2, "vocab_size": 4096, "sliding_window": 64}, 100, 4),
                                      ("mistral-7b", {"n_layers": 8}, 0, 4), ("tiny-moe", {}, 9, 4),
                                      ("mixtral-8x7b", {"n_layers": 2, "vocab_size": 4096}, 0, 4)]:
    p = synth.shape(name, **over)
    log("==", name, over)
    m = model_for(p)
    cache = BufferCache(p["n_layers"], 1, max(prompt_len, 1) + steps + 4200, p["n_kv_heads"], p["head_dim"], p.get("sliding_window"))
    cache.to(m.device, m.dtype)
    for i in cache.cache_k:
        cache.cache_k[i].zero_()
        cache.cache_v[i].zero_()
    if prompt_len:
        m.forward(torch.tensor(synth.synth_prompt(prompt_len, p["vocab_size"], 3), device="cuda"), [prompt_len], cache)
        torch.cuda.synchronize()
        log("  prefill ok")
    else:
        cache._kv_seqlens_host = [5000]
    for s in range(steps):
        t0 = time.time()
        lg = m.decode_static(torch.tensor([s + 1], device="cuda"), cache)
        torch.cuda.synchronize()
        log(f"  step {s} ok {1e3 * (time.time() - t0):.2f} ms  logits[:3]={lg[0, :3].tolist()}")
log("ALL OK")
