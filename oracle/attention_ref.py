"""Oracle attention arithmetic (TEST INFRASTRUCTURE -- see oracle/__init__.py).

fp32 restatement of what the reference obtains from
`xformers.ops.fmha.memory_efficient_attention(q, k, v, attn_bias)` at
transformer_layers.py:88, for the three mask kinds cache.py builds (cache.py:240,243-248,250-254;
SURVEY.md Appendix B):

  softmax(q k^T * hd^-0.5 + mask) v      inputs [1, S, H, hd], output contiguous [1, Sq, H, hd]
                                         in q's dtype; scores / softmax / PV in fp32.

xformers 0.0.26.post1 is un-vendored and CUDA-only, so this file is the oracle's definition of
that boundary ("parity unpinned" there except via the decode == re-prefill self-consistency).
Both the import shim (oracle/ref_shims.py) and the restatement (oracle/restatement.py) call
`attend_block`, so the two agree bit-for-bit by construction.
"""
from typing import Optional

import torch


def attend_block(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, allowed: Optional[torch.Tensor]) -> torch.Tensor:
    """One sequence block.  q [s, H, hd], k/v [n, H, hd] (kv heads already repeated, as
    transformer_layers.py:84 does), allowed: bool [s, n] or None (= attend to everything).
    Returns [s, H, hd] in q.dtype."""
    s, H, hd = q.shape
    qf = q.float().permute(1, 0, 2)  # [H, s, hd]
    kf = k.float().permute(1, 2, 0)  # [H, hd, n]
    vf = v.float().permute(1, 0, 2)  # [H, n, hd]
    scores = torch.matmul(qf, kf) * (hd ** -0.5)  # default xformers scale; Attention.scale is unused (SURVEY E-3)
    if allowed is not None:
        scores = scores.masked_fill(~allowed[None, :, :], float("-inf"))
    probs = torch.softmax(scores, dim=-1)
    out = torch.matmul(probs, vf)  # [H, s, hd]
    return out.permute(1, 0, 2).contiguous().to(q.dtype)


def local_causal_allowed(s: int, n: int, window: Optional[int], device=None) -> torch.Tensor:
    """Bottom-right aligned causal + local window: query i (0..s-1) may see key j (0..n-1) iff
         j <= i + (n - s)   and   j > i + (n - s) - window
    With n == s this is BlockDiagonalCausalMask.make_local_attention(window) (cache.py:240);
    with n > s it is BlockDiagonalMask.make_local_attention_from_bottomright(window) (cache.py:243-248).
    window None = no lower bound."""
    i = torch.arange(s, device=device)[:, None]
    j = torch.arange(n, device=device)[None, :]
    off = n - s
    allowed = j <= i + off
    if window is not None:
        allowed &= j > i + off - window
    return allowed
