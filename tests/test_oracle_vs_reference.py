"""Pins the oracle restatement bit-exactly against the reference's own modules (run unmodified
behind oracle/ref_shims.py).  Needs /root/reference, so it runs in the build container only."""
import pytest
import torch

import synth
from oracle import ref_shims
from oracle import restatement as R

from .util import oracle_args

pytestmark = pytest.mark.skipif(not ref_shims.reference_available(), reason="reference tree not present (GPU box)")


def _ref_model(p, max_batch, dtype, seed=3):
    ref = ref_shims.import_reference()
    args = ref.args.TransformerArgs.from_dict(dict(p))
    args.max_batch_size = max_batch
    with torch.device("meta"):
        m = ref.transformer.Transformer(args)
    m.load_state_dict(synth.synth_state_dict(p, seed, dtype), assign=True, strict=True)
    return ref, m.eval()


@pytest.mark.parametrize("shape,over", [
    ("tiny", {}), ("tiny", {"sliding_window": 5}), ("tiny", {"sliding_window": [4, None]}),
    ("tiny-moe", {}), ("tiny-moe", {"sliding_window": 3}),
])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_generate_bit_exact(shape, over, dtype):
    p = synth.shape(shape, **over)
    ref, rm = _ref_model(p, 3, dtype)
    om = R.OracleTransformer(oracle_args(p, 3), synth.synth_state_dict(p, 3, dtype))
    prompts = [synth.synth_prompt(n, p["vocab_size"], 40 + i) for i, n in enumerate([11, 9, 10])]
    t_ref, lp_ref = ref.generate.generate(prompts, rm, max_tokens=9, temperature=0.0, chunk_size=4)
    t_or, lp_or = R.generate(prompts, om, max_tokens=9, chunk_size=4)
    assert t_ref == t_or
    assert lp_ref == lp_or  # python floats from identical fp32 tensors


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_forward_without_cache_bit_exact(dtype):
    """cache=None: unmasked, cross-sequence attention."""
    p = synth.shape("tiny")
    ref, rm = _ref_model(p, 2, dtype)
    om = R.OracleTransformer(oracle_args(p, 2), synth.synth_state_dict(p, 3, dtype))
    toks = torch.tensor(synth.synth_prompt(13, p["vocab_size"], 5))
    with torch.inference_mode():
        a = rm.forward(toks, seqlens=[6, 7])
        b = om.forward(toks, [6, 7])
    assert torch.equal(a, b)


def test_elementwise_ops_bit_exact():
    ref = ref_shims.import_reference()
    import mistral_inference.rope as r_rope
    import mistral_inference.transformer_layers as r_layers

    torch.manual_seed(0)
    x = torch.randn(7, 256).to(torch.bfloat16)
    w = (1 + 0.1 * torch.randn(256)).to(torch.bfloat16)
    norm = r_layers.RMSNorm(256, eps=1e-5)
    norm.weight.data = w
    assert torch.equal(norm(x), R.rms_norm(x, w, 1e-5))
    table = r_rope.precompute_freqs_cis(128, 1000, 1e6)
    assert torch.equal(torch.view_as_real(table), torch.view_as_real(R.rope_table(128, 1000, 1e6)))
    q = torch.randn(7, 4, 128).to(torch.bfloat16)
    k = torch.randn(7, 2, 128).to(torch.bfloat16)
    fc = table[torch.tensor([0, 1, 2, 500, 501, 998, 999])]
    for a, b in zip(r_rope.apply_rotary_emb(q, k, fc), R.apply_rope(q, k, fc)):
        assert torch.equal(a, b)
