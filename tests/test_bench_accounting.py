"""bench.py's roofline numerators against an independent count made from the checkpoint's own tensor shapes:
algorithmic bytes of one batch-1 decode step = every weight byte the step must read once + the visible KV rows."""
import math

import bench
import synth


def _weight_bytes_from_shapes(p: dict) -> int:
    moe = p.get("moe") or {}
    total = 0
    for key, shape in synth.state_dict_shapes(p):
        if key == "tok_embeddings.weight":
            continue  # one row gathered per token: noise
        n = math.prod(shape) * 2
        if moe and ".experts." in key:
            total += n * moe["num_experts_per_tok"] / moe["num_experts"]  # k of E experts are streamed at batch 1
        else:
            total += n
    return int(total)


def test_decode_bytes_dense_and_moe():
    for name in ("mistral-7b", "mixtral-8x7b", "tiny", "tiny-moe"):
        p = synth.shape(name)
        kv_len = 4096
        kv = 2 * p["n_layers"] * 2 * kv_len * p["n_kv_heads"] * p["head_dim"]
        assert bench.decode_bytes_per_step(p, kv_len) == _weight_bytes_from_shapes(p) + kv, name
    assert bench.decode_bytes_per_step(synth.shape("mistral-7b"), 4096) == 14_758_191_104  # the bench line's bytes_per_launch


def test_prefill_flops_counts_every_linear_and_the_causal_half():
    p, T = synth.shape("mistral-7b"), 4096
    linear_params = sum(math.prod(s) for k, s in synth.state_dict_shapes(p) if len(s) == 2 and k != "tok_embeddings.weight")
    attn = 4.0 * p["n_layers"] * p["n_heads"] * p["head_dim"] * (T * (T + 1) // 2)  # QK^T and PV over the visible keys
    assert bench.prefill_flops(p, T) == 2.0 * T * linear_params + attn
