"""The oracle restatement reproduces the committed outputs of the reference (tests/golden/)."""
import pytest
import torch

from oracle import restatement as R

from .util import GOLDEN_CASES, case_params_prompts, load_golden, oracle_model, same_machine_as_golden


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_golden(name):
    case, gold, meta = load_golden(name)
    p, prompts = case_params_prompts(case)
    dtype = getattr(torch, case["dtype"])
    model = oracle_model(p, case["max_batch"], dtype=dtype)
    toks, logprobs, step_logits = R.generate(prompts, model, max_tokens=case["max_tokens"], return_logits=True)

    # full prefill logits
    cache = model.new_cache(max(len(x) for x in prompts) + case["max_tokens"])
    prefill = model.forward(torch.tensor(sum(prompts, [])), [len(x) for x in prompts], cache)
    exact = same_machine_as_golden(meta)
    if exact:
        # same torch build + same CPU ISA level as the fixture generator: the restatement is bit-exact
        assert torch.equal(prefill, gold["prefill_logits"])
        assert toks == gold["tokens"].tolist()
        assert torch.equal(torch.stack(step_logits[1:], 0), gold["decode_logits"][:-1])
        assert torch.equal(torch.tensor(sum(logprobs, []), dtype=torch.float64), gold["logprobs"])
    else:  # a different CPU may pick other GEMM kernels: fp32-accumulation-order noise through bf16 roundings
        tol = 1e-4 if dtype == torch.float32 else 6e-2
        torch.testing.assert_close(prefill, gold["prefill_logits"], rtol=0, atol=tol)

    # the reference's own property (tests/test_generate.py:36-69,199-230): decode == (chunked) re-prefill
    full = [pr + t for pr, t in zip(prompts, toks)]
    gen2, logprobs2 = R.generate(full, model, max_tokens=0, chunk_size=case["chunk"])
    assert gen2 == []
    if exact:
        assert torch.equal(torch.tensor(sum(logprobs2, []), dtype=torch.float64), gold["reprefill_logprobs"])
    bound = 5e-4 if dtype == torch.float32 else 0.12  # fp32 bound is the reference's; bf16 has 8 mantissa bits
    for a, b in zip(logprobs, logprobs2):
        assert len(a) == len(b)
        assert max(abs(x - y) for x, y in zip(a, b)) < bound


def test_oracle_matches_config1_golden():
    """BASELINE.json configs[0] (Mistral-7B shape, 1 layer, batch 1, 128 + 32): the oracle restatement against the committed
    outputs of the reference's generate() -- bit-exact on the fixture's machine type, decisive picks identical anywhere."""
    import synth

    from .util import oracle_args

    case, gold, meta = load_golden("config1_7b_1layer")
    seed = int(meta["seed"])
    p = synth.shape(case["shape"], **case["over"])
    prompts = [synth.synth_prompt(n, p["vocab_size"], seed * 100 + i) for i, n in enumerate(case["prompt_lens"])]
    model = R.OracleTransformer(oracle_args(p, 1), synth.synth_state_dict(p, seed, torch.bfloat16))
    toks, logprobs, step_logits = R.generate(prompts, model, max_tokens=case["max_tokens"], return_logits=True)
    if same_machine_as_golden(meta):
        assert toks == gold["tokens"].tolist()
        assert torch.equal(torch.tensor(sum(logprobs, []), dtype=torch.float64), gold["logprobs"])
        top = torch.cat(step_logits, 0).topk(64, dim=-1)
        assert torch.equal(top.values, gold["topk_values"])
    else:
        prefix = int(meta["decisive_prefix"])
        assert toks[0][:prefix] == gold["tokens"][0].tolist()[:prefix]
