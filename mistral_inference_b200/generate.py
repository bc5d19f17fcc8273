"""Prefill-by-chunks + decode driver (API mirror of mistral_inference/generate.py:43-170)."""
from typing import List, Optional, Tuple

import torch

from .cache import BufferCache
from .transformer import Transformer


@torch.inference_mode()
def generate(encoded_prompts: List[List[int]], model: Transformer, images: List[List] = [], *, max_tokens: int,  # noqa: B006
             temperature: float, chunk_size: Optional[int] = None, eos_id: Optional[int] = None
             ) -> Tuple[List[List[int]], List[List[float]]]:
    """Same contract as generate.py:43-148: returns (generated tokens, logprobs of prompt tokens 1.. and of the
    generated tokens); greedy when temperature == 0, else top-p with p = 0.8; stops when every sequence has emitted
    eos (finished sequences keep generating); `[]` tokens when max_tokens == 0."""
    assert not images, "vision inputs are outside the accelerated hot path"
    model = model.eval()
    B, V = len(encoded_prompts), model.args.vocab_size
    seqlens = [len(x) for x in encoded_prompts]

    # Cache (generate.py:68-78)
    cache_window = max(seqlens) + max_tokens
    cache = BufferCache(model.n_local_layers, model.args.max_batch_size, cache_window, model.args.n_kv_heads, model.args.head_dim,
                        model.args.sliding_window)
    cache.to(device=model.device, dtype=model.dtype)
    cache.reset()

    logprobs: List[List[float]] = [[] for _ in range(B)]
    last_token_prelogits = None
    max_prompt_len = max(seqlens)
    if chunk_size is None:
        chunk_size = max_prompt_len

    # Encode prompt by chunks (generate.py:92-118)
    for s in range(0, max_prompt_len, chunk_size):
        prompt_chunks = [p[s:s + chunk_size] for p in encoded_prompts]
        assert all(len(p) > 0 for p in prompt_chunks)
        flat = sum(prompt_chunks, [])
        prelogits = model.forward(torch.tensor(flat, device=model.device, dtype=torch.long), seqlens=[len(p) for p in prompt_chunks],
                                  cache=cache)
        logits = torch.log_softmax(prelogits, dim=-1)
        if last_token_prelogits is not None:
            last_token_logits = torch.log_softmax(last_token_prelogits, dim=-1)
            firsts = torch.tensor([p[0] for p in prompt_chunks], device=logits.device)
            for i_seq, lp in enumerate(last_token_logits.gather(1, firsts[:, None])[:, 0].tolist()):
                logprobs[i_seq].append(lp)
        # logprob of token i+1 under the distribution at token i: one gather + one D2H instead of a .item() per token
        nxt = torch.tensor(flat[1:] + [0], device=logits.device)
        picked = logits.gather(1, nxt[:, None])[:, 0].tolist()
        offset = 0
        for i_seq, sequence in enumerate(prompt_chunks):
            logprobs[i_seq].extend(picked[offset:offset + len(sequence) - 1])
            offset += len(sequence)
        last_idx = torch.tensor([len(p) for p in prompt_chunks], device=prelogits.device).cumsum(dim=0) - 1
        last_token_prelogits = prelogits.index_select(0, last_idx)
        assert last_token_prelogits.shape == (B, V)

    # decode (generate.py:120-140)
    generated_tensors = []
    is_finished = torch.tensor([False for _ in range(B)])
    assert last_token_prelogits is not None
    for _ in range(max_tokens):
        next_token = sample(last_token_prelogits, temperature=temperature, top_p=0.8)
        if eos_id is not None:
            is_finished = is_finished | (next_token == eos_id).cpu()
        if is_finished.all():
            break
        last_token_logits = torch.log_softmax(last_token_prelogits, dim=-1)
        for i, lp in enumerate(last_token_logits.gather(1, next_token[:, None])[:, 0].tolist()):
            logprobs[i].append(lp)
        generated_tensors.append(next_token[:, None])
        last_token_prelogits = model.forward(next_token, seqlens=[1] * B, cache=cache)
        assert last_token_prelogits.shape == (B, V)

    generated_tokens: List[List[int]]
    if generated_tensors:
        generated_tokens = torch.cat(generated_tensors, 1).tolist()
    else:
        generated_tokens = []
    return generated_tokens, logprobs


def sample(logits: torch.Tensor, temperature: float, top_p: float) -> torch.Tensor:
    """generate.py:151-158."""
    if temperature > 0:
        probs = torch.softmax(logits / temperature, dim=-1)
        next_token = sample_top_p(probs, top_p)
    else:
        next_token = torch.argmax(logits, dim=-1).unsqueeze(0)
    return next_token.reshape(-1)


def sample_top_p(probs: torch.Tensor, p: float) -> torch.Tensor:
    """generate.py:161-170."""
    assert 0 <= p <= 1
    probs_sort, probs_idx = torch.sort(probs, dim=-1, descending=True)
    probs_sum = torch.cumsum(probs_sort, dim=-1)
    mask = probs_sum - probs_sort > p
    probs_sort[mask] = 0.0
    probs_sort.div_(probs_sort.sum(dim=-1, keepdim=True))
    next_token = torch.multinomial(probs_sort, num_samples=1)
    return torch.gather(probs_idx, -1, next_token)
