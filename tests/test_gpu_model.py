"""GPU parity of the whole path through the public API (Transformer.forward / generate), against
  (a) the committed outputs of the reference (tests/golden/), teacher-forced;
  (b) the oracle restatement run on this machine's CPU on the same weights;
  (c) the reference's own consistency property decode == (chunked) re-prefill (tests/test_generate.py:36-69,199-230).

Tolerance: logits are bf16 values (stored as fp32) of magnitude <= ~2.4 here, i.e. 1 ulp = 2^-7..2^-6.  After L
layers of bf16 roundings the CUDA path and the CPU oracle differ by a few ulps on a few logits: LOGIT_ATOL below.
The north-star's "rtol 1e-3 / atol 1e-5" is tighter than one bf16 ulp (2^-8 relative) and therefore not
attainable element-wise between ANY two correct bf16 implementations with different summation order (the oracle on
another CPU does not meet it against itself: tests/test_oracle_golden.py); measured deltas are printed.
Greedy token ids must match wherever the reference's own top-2 margin exceeds the tolerance.
"""
import pytest
import torch

import mistral_inference_b200 as mi
from mistral_inference_b200 import synth
from mistral_inference_b200.cache import BufferCache
from mistral_inference_b200.transformer import Transformer
from oracle import restatement as R

from .util import GOLDEN_CASES, case_params_prompts, load_golden, oracle_model

pytestmark = pytest.mark.gpu
LOGIT_ATOL = 0.06
BF16_CASES = [c for c in GOLDEN_CASES if not c.endswith("fp32")]


def gpu_model(p: dict, max_batch: int, seed: int = 1) -> Transformer:
    args = mi.TransformerArgs.from_dict(dict(p))
    args.max_batch_size = max_batch
    with torch.device("cuda"):
        m = Transformer(args).to(torch.bfloat16)
    m.load_state_dict(synth.synth_state_dict(p, seed, torch.bfloat16, "cuda"))
    return m.eval()


def new_cache(m: Transformer, max_seq: int) -> BufferCache:
    a = m.args
    c = BufferCache(m.n_local_layers, a.max_batch_size, max_seq, a.n_kv_heads, a.head_dim, a.sliding_window)
    c.to(m.device, m.dtype)
    for i in c.cache_k:  # uninitialised in the reference (cache.py:166): poison so that a masking bug is loud
        c.cache_k[i].fill_(float("nan"))
        c.cache_v[i].fill_(float("nan"))
    c.reset()
    return c


def report(tag, got, want):
    d = (got.float().cpu() - want.float().cpu()).abs()
    print(f"\n[parity] {tag}: max|d|={d.max():.4f} mean|d|={d.mean():.5f} exact={(d == 0).float().mean():.3f}")
    return d


def check_logits(d: torch.Tensor, moe: bool) -> None:
    """Dense: every logit within LOGIT_ATOL.  MoE: top-k routing is discontinuous -- when two router logits are
    within a bf16 ulp the CUDA path and the oracle may legitimately pick different experts for a token, which moves
    that token's logits by more than rounding noise; tolerate a few such rows, bound everything else."""
    if not moe:
        assert d.max() <= LOGIT_ATOL, d.max()
        return
    bad_rows = (d.max(dim=-1).values > LOGIT_ATOL)
    assert bad_rows.float().mean() <= 0.2, f"{int(bad_rows.sum())}/{bad_rows.numel()} rows beyond tolerance"
    assert d[~bad_rows].max() <= LOGIT_ATOL and d.mean() <= 0.01


@pytest.mark.parametrize("name", BF16_CASES)
def test_golden_teacher_forced(name):
    case, gold, _ = load_golden(name)
    p, prompts = case_params_prompts(case)
    m = gpu_model(p, case["max_batch"])
    B = len(prompts)
    cache = new_cache(m, max(len(x) for x in prompts) + case["max_tokens"])
    logits = m.forward(torch.tensor(sum(prompts, []), device="cuda"), [len(x) for x in prompts], cache)
    moe = p.get("moe") is not None
    d = report(f"{name} prefill", logits, gold["prefill_logits"])
    check_logits(d, moe)
    toks = gold["tokens"]  # [B, max_tokens] the reference's greedy choices
    agree = total = 0
    for step in range(toks.shape[1]):
        logits = m.forward(toks[:, step].to("cuda"), [1] * B, cache)
        want = gold["decode_logits"][step]
        d = report(f"{name} decode step {step}", logits, want)
        check_logits(d, moe)
        if step + 1 < toks.shape[1] and not moe:  # the next greedy token, wherever the reference's margin is decisive
            top2 = want.topk(2, dim=-1).values
            decisive = (top2[:, 0] - top2[:, 1]) > 2 * LOGIT_ATOL
            pick = logits.argmax(-1).cpu()
            assert torch.equal(pick[decisive], toks[:, step + 1][decisive])
            agree += int((pick == toks[:, step + 1]).sum())
            total += B
    print(f"[parity] {name}: greedy token agreement with the reference {agree}/{total}")


@pytest.mark.parametrize("shape,over,lens,chunk", [
    ("tiny", {}, [11, 9, 10], 4),
    ("tiny", {"sliding_window": 5}, [11, 9, 10], 4),
    ("tiny", {"sliding_window": [4, None]}, [70, 68], 33),
    ("tiny-moe", {"sliding_window": 3}, [11, 12], 5),
    ("ref-test", {}, [8, 4, 4, 4], None),  # the shape of the reference's own tests (tests/test_generate.py:40-50)
])
def test_generate_vs_oracle_and_self_consistency(shape, over, lens, chunk):
    p = synth.shape(shape, **over)
    if shape == "ref-test":
        p["vocab_size"] = 4096  # keep the CPU oracle quick
    prompts = [synth.synth_prompt(n, p["vocab_size"], 60 + i) for i, n in enumerate(lens)]
    B, max_tokens = len(prompts), 6
    m = gpu_model(p, B + 1)  # max_batch_size > B exercises cache[:B] (tests/test_generate.py:212)
    om = oracle_model(p, B + 1)
    # (b) teacher-forced on the oracle's greedy tokens
    o_toks, o_lp, o_step = R.generate(prompts, om, max_tokens=max_tokens, chunk_size=chunk, return_logits=True)
    cache = new_cache(m, max(lens) + max_tokens)
    last = None
    for s in range(0, max(lens), chunk or max(lens)):
        chunks = [pr[s:s + (chunk or max(lens))] for pr in prompts]
        logits = m.forward(torch.tensor(sum(chunks, []), device="cuda"), [len(c) for c in chunks], cache)
        last = logits[torch.tensor([len(c) for c in chunks]).cumsum(0) - 1]
    for step in range(max_tokens):
        d = report(f"{shape}{over} step {step}", last, o_step[step])
        check_logits(d, p.get("moe") is not None)
        last = m.forward(torch.tensor([t[step] for t in o_toks], device="cuda"), [1] * B, cache)
    # (c) the reference's property through the public generate(): decode == chunked re-prefill
    toks, lp = mi.generate(prompts, m, max_tokens=max_tokens, temperature=0.0)
    assert len(toks) == B and all(len(t) == max_tokens for t in toks)
    full = [pr + t for pr, t in zip(prompts, toks)]
    if chunk is not None:  # every prompt needs a token in every chunk (generate.py:94)
        n_chunks = -(-max(len(f) for f in full) // chunk)
        if min(len(f) for f in full) <= chunk * (n_chunks - 1):
            chunk = None
    gen2, lp2 = mi.generate(full, m, max_tokens=0, temperature=0.0, chunk_size=chunk)
    assert gen2 == []
    worst = max(abs(a - b) for x, y in zip(lp, lp2) for a, b in zip(x, y))
    print(f"[parity] {shape}{over}: decode vs re-prefill logprob max|d|={worst:.4f}")
    assert all(len(x) == len(y) for x, y in zip(lp, lp2))
    assert worst < 0.12  # bf16; the reference's 5e-4 is its fp32 bound


def test_forward_without_cache():
    p = synth.shape("tiny")
    m, om = gpu_model(p, 2), oracle_model(p, 2)
    toks = torch.tensor(synth.synth_prompt(13, p["vocab_size"], 5))
    d = report("no-cache forward", m.forward(toks.cuda(), [6, 7]), om.forward(toks, [6, 7]))
    assert d.max() <= LOGIT_ATOL


@pytest.mark.parametrize("over,lens", [
    ({}, [700]),                          # one long sequence: 2-CTA cluster GEMMs (T >= 512) + tcgen05 flash attention, ragged last tile
    ({"sliding_window": 200}, [640, 300]),  # two sequences, window shorter than the prompt: window edge tiles + ring wrap on write
])
def test_long_first_prefill_vs_oracle(over, lens):
    """The tensor-core prefill kernels at model level: a first prefill long enough for the tcgen05 GEMM (cluster pairs, row-chunk
    epilogues: RoPE + ring scatter, residual, SiLU*mul, fp32 logits) and the tcgen05 attention, against the CPU oracle on the
    same weights; then one decode step off the cache that prefill wrote."""
    p = synth.shape("tiny", **over)
    m, om = gpu_model(p, len(lens)), oracle_model(p, len(lens))
    toks = torch.tensor(synth.synth_prompt(sum(lens), p["vocab_size"], 11))
    cache, ocache = new_cache(m, max(lens) + 8), om.new_cache(max(lens) + 8)
    d = report(f"long first prefill {over} {lens}", m.forward(toks.cuda(), lens, cache), om.forward(toks, lens, ocache))
    assert d.max() <= LOGIT_ATOL and d.mean() <= 0.004
    nxt = torch.tensor([3 + b for b in range(len(lens))])
    d = report(f"decode after long prefill {over} {lens}", m.forward(nxt.cuda(), [1] * len(lens), cache), om.forward(nxt, [1] * len(lens), ocache))
    assert d.max() <= LOGIT_ATOL


def test_sampling_path_runs():
    p = synth.shape("tiny")
    m = gpu_model(p, 2)
    torch.manual_seed(0)
    toks, lp = mi.generate([[1, 2, 3], [4, 5, 6, 7]], m, max_tokens=5, temperature=0.7, eos_id=None)
    assert len(toks) == 2 and all(len(t) == 5 for t in toks) and all(0 <= x < p["vocab_size"] for t in toks for x in t)


def test_full_size_7b_layer_properties():
    """Mistral-7B layer shapes (BASELINE.json configs[1]) at sizes the CPU oracle cannot do quickly, through
    size-independent properties: decode == re-prefill after the 4096-slot ring has wrapped."""
    p = synth.shape("mistral-7b", n_layers=2, vocab_size=4096, sliding_window=256)
    m = gpu_model(p, 1)
    prompt = synth.synth_prompt(300, p["vocab_size"], 9)  # > W: the ring wraps during prefill
    toks, lp = mi.generate([prompt], m, max_tokens=40, temperature=0.0, chunk_size=128)
    gen2, lp2 = mi.generate([prompt + toks[0]], m, max_tokens=0, temperature=0.0, chunk_size=170)
    worst = max(abs(a - b) for a, b in zip(lp[0], lp2[0]))
    print(f"[parity] 7B-shape 2-layer ring-wrap consistency: max|d logprob|={worst:.4f}")
    assert worst < 0.12


@pytest.mark.parametrize("shape,over,prompt_len,steps", [
    ("tiny", {}, 9, 12),
    ("tiny", {"sliding_window": 6}, 9, 12),                 # ring wraps
    ("tiny", {"sliding_window": [5, None]}, 20, 8),
    ("mistral-7b", {"n_layers": 2, "vocab_size": 4096, "sliding_window": 64}, 100, 6),  # real layer shapes (K chunks of 3584)
    ("tiny-moe", {}, 9, 10),                                  # in-kernel router + experts (moe.py:24-32)
    ("tiny-moe", {"sliding_window": 5}, 12, 8),
    ("mixtral-8x7b", {"n_layers": 2, "vocab_size": 4096}, 40, 4),  # real expert shapes: 2 x (8 experts x 176 M params)
    ("mistral-nemo-12b", {"n_layers": 2, "vocab_size": 4096}, 140, 4),  # BASELINE config 3 shape: dim 5120 != H*hd, no window
    ("mixtral-8x22b", {"n_layers": 2, "vocab_size": 4096}, 40, 3),      # BASELINE config 5 shape: H/KV = 6, dim 6144, hidden 16384
])
def test_decode_megakernel(shape, over, prompt_len, steps, monkeypatch):
    """The persistent one-kernel-per-token decode step against (a) the per-op kernel path and (b) the CPU oracle."""
    p = synth.shape(shape, **over)
    m = gpu_model(p, 1)
    prompt = synth.synth_prompt(prompt_len, p["vocab_size"], 21)
    toks = synth.synth_prompt(steps, p["vocab_size"], 22)  # teacher-forced continuation

    def run(megakernel: bool):
        monkeypatch.setenv("MB200_MEGAKERNEL", "1" if megakernel else "0")
        cache = new_cache(m, prompt_len + steps + 1)
        out = [m.forward(torch.tensor(prompt, device="cuda"), [prompt_len], cache)[-1:].clone()]
        for t in toks:
            out.append(m.forward(torch.tensor([t], device="cuda"), [1], cache).clone())
        return torch.cat(out[1:], 0), cache

    mk, c1 = run(True)
    # fused greedy argmax of the last step == torch.argmax of the logits it produced (first index on ties)
    assert int(m.last_argmax.item()) == int(mk[-1].argmax().item())
    per_op, c2 = run(False)
    moe = p.get("moe") is not None
    d = report(f"megakernel vs per-op {shape}{over}", mk, per_op)
    check_logits(d, moe)
    for i in ([] if moe else c1.cache_k):  # (an expert flip changes later layers' K/V legitimately: dense models only)  # the rings written by both paths agree (1 bf16 ulp on a few elements)
        a, b = c1.cache_k[i][0].float(), c2.cache_k[i][0].float()
        ok = torch.isfinite(b)
        assert torch.equal(torch.isfinite(a), ok)
        assert (a[ok] - b[ok]).abs().max() <= 0.07
    if shape.startswith("tiny"):
        om = oracle_model(p, 1)
        oc = om.new_cache(prompt_len + steps + 1)
        om.forward(torch.tensor(prompt), [prompt_len], oc)
        want = torch.cat([om.forward(torch.tensor([t]), [1], oc) for t in toks], 0)
        d = report(f"megakernel vs oracle {shape}{over}", mk, want)
        check_logits(d, moe)
