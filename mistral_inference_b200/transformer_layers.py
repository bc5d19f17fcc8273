"""Attention / FeedForward / RMSNorm / TransformerBlock (API mirror of mistral_inference/transformer_layers.py).

The classes keep the reference's names and wiring, but every FLOP goes through libmb200 (C ABI in
include/mistral_b200.h); weights are stored pre-packed for the fused kernels:
  Attention.wqkv  [(H + 2*KV) * hd, dim] = wq ++ wk ++ wv        (one GEMM / one weight stream)
  FeedForward.w13 [2 * hidden, dim], row 2i = w1[i], row 2i+1 = w3[i]  (SiLU*mul in the epilogue)
`wq/wk/wv/w1/w3` are exposed as zero-copy views for state-dict compatibility.
"""
from typing import Optional, Tuple

import torch
from torch import nn

from . import _abi
from .args import LoraArgs, MoeArgs
from .cache import CacheView
from .moe import MoeLayer


class _WeightView:
    """Stands in for an nn.Linear whose weight is a view into a packed parameter."""

    def __init__(self, getter):
        self._getter = getter

    @property
    def weight(self) -> torch.Tensor:
        return self._getter()


class RMSNorm(nn.Module):
    """transformer_layers.py:109-120."""

    def __init__(self, dim: int, eps: float = 1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim), requires_grad=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _abi.rmsnorm(x, self.weight, self.eps)


class Attention(nn.Module):
    """transformer_layers.py:31-93."""

    def __init__(self, dim: int, n_heads: int, head_dim: int, n_kv_heads: int, lora: Optional[LoraArgs] = None):
        super().__init__()
        assert lora is None, "LoRA adapters must be merged into the weights (lora.py:118-139); unmerged LoRA is out of scope"
        self.dim = dim
        self.n_heads = n_heads
        self.head_dim = head_dim
        self.n_kv_heads = n_kv_heads
        self.repeats = n_heads // n_kv_heads
        self.scale = head_dim ** -0.5
        self.q_dim = n_heads * head_dim
        self.kv_dim = n_kv_heads * head_dim
        self.wqkv = nn.Parameter(torch.empty(self.q_dim + 2 * self.kv_dim, dim), requires_grad=False)
        self.wo_weight = nn.Parameter(torch.empty(dim, self.q_dim), requires_grad=False)

    # state-dict compatible views
    @property
    def wq(self) -> _WeightView:
        return _WeightView(lambda: self.wqkv[: self.q_dim])

    @property
    def wk(self) -> _WeightView:
        return _WeightView(lambda: self.wqkv[self.q_dim: self.q_dim + self.kv_dim])

    @property
    def wv(self) -> _WeightView:
        return _WeightView(lambda: self.wqkv[self.q_dim + self.kv_dim:])

    @property
    def wo(self) -> _WeightView:
        return _WeightView(lambda: self.wo_weight)

    def attend(self, x: torch.Tensor, norm_w: torch.Tensor, eps: float, rope: torch.Tensor, positions: torch.Tensor,
               cache: Optional[CacheView], ws: "_abi.Workspace") -> torch.Tensor:
        """RMSNorm + QKV + RoPE + cache phase + attention core.  Returns the pre-`wo` output [T, H*hd]."""
        T = x.shape[0]
        q = torch.empty(T, self.q_dim, dtype=x.dtype, device=x.device)
        k = torch.empty(T, self.kv_dim, dtype=x.dtype, device=x.device)
        v = torch.empty(T, self.kv_dim, dtype=x.dtype, device=x.device)
        out = torch.empty(T, self.q_dim, dtype=x.dtype, device=x.device)
        H, KV, hd = self.n_heads, self.n_kv_heads, self.head_dim
        if cache is None:
            # cache-less forward: unmasked over the whole flattened batch
            _abi.attn_qkv(x, norm_w, self.wqkv, rope, positions, q, k, v, None, None, None, H, KV, hd, eps, ws)
            _abi.attn_prefill(q, k, v, None, None, None, None, out, 1, T, 0, H, KV, hd, causal=False)
            return out
        md = cache.metadata
        if md.prefill:
            # read the old ring, THEN write (transformer_layers.py:75-76)
            _abi.attn_qkv(x, norm_w, self.wqkv, rope, positions, q, k, v, None, None, None, H, KV, hd, eps, ws)
            _abi.attn_prefill(q, k, v, cache.cache_k, cache.cache_v, md.q_start, md.seqpos, out, len(md.seqlens), md.max_seqlen,
                              md.window, H, KV, hd, causal=True, first_prefill=md.first_prefill)
            _abi.kv_ring_write(k, v, cache.cache_k, cache.cache_v, md.cache_rows, KV, hd)
        else:
            # write, THEN read the ring (transformer_layers.py:78-81); the scatter is the QKV kernel's epilogue
            _abi.attn_qkv(x, norm_w, self.wqkv, rope, positions, q, k, v, cache.cache_k, cache.cache_v, md.cache_rows, H, KV, hd, eps, ws)
            B = len(md.seqlens)
            _abi.attn_decode(q, cache.cache_k, cache.cache_v, md.kv_len, out, H, KV, hd, decode_splits(B, KV, md.window), ws)
        return out


def decode_splits(B: int, KV: int, W: int, n_sm: int = 148) -> int:
    """KV splits per (sequence, kv head): as many as fit ONE wave of two CTAs per SM (a second, partial wave costs more than the
    idle SMs of an incomplete first one; with B * KV >= 2 * n_sm / 2 there is no split and no combine step at all), at least 64
    keys per split."""
    s = max(1, (2 * n_sm) // (B * KV))
    return int(max(1, min(s, 64, (W + 63) // 64)))


class FeedForward(nn.Module):
    """transformer_layers.py:96-106."""

    def __init__(self, dim: int, hidden_dim: int, lora: Optional[LoraArgs] = None):
        super().__init__()
        assert lora is None
        self.dim = dim
        self.hidden_dim = hidden_dim
        self.w13 = nn.Parameter(torch.empty(2 * hidden_dim, dim), requires_grad=False)
        self.w2_weight = nn.Parameter(torch.empty(dim, hidden_dim), requires_grad=False)

    @property
    def w1(self) -> _WeightView:
        return _WeightView(lambda: self.w13.view(self.hidden_dim, 2, self.dim)[:, 0])

    @property
    def w3(self) -> _WeightView:
        return _WeightView(lambda: self.w13.view(self.hidden_dim, 2, self.dim)[:, 1])

    @property
    def w2(self) -> _WeightView:
        return _WeightView(lambda: self.w2_weight)

    def run(self, x: torch.Tensor, norm_w: Optional[torch.Tensor], eps: float, residual: Optional[torch.Tensor],
            ws: "_abi.Workspace") -> torch.Tensor:
        """[norm] -> gate/up -> silu*mul -> down [+ residual]."""
        T = x.shape[0]
        g = torch.empty(T, self.hidden_dim, dtype=x.dtype, device=x.device)
        _abi.ffn_gateup(x, norm_w, self.w13, g, eps, ws)
        out = torch.empty(T, self.dim, dtype=x.dtype, device=x.device)
        _abi.linear_residual(g, self.w2_weight, residual, out, ws)
        return out

    def forward(self, x: torch.Tensor, ws: Optional["_abi.Workspace"] = None) -> torch.Tensor:
        ws = ws or _abi.Workspace(_abi.workspace_bytes(x.shape[0], self.dim, 1, 1, 128, self.hidden_dim, 0, 1), x.device)
        return self.run(x, None, 0.0, None, ws)


class TransformerBlock(nn.Module):
    """transformer_layers.py:123-169: pre-norm residual wiring; FeedForward or MoeLayer."""

    def __init__(self, dim: int, hidden_dim: int, n_heads: int, n_kv_heads: int, head_dim: int, norm_eps: float,
                 lora: Optional[LoraArgs] = None, moe: Optional[MoeArgs] = None, expert_shard: Tuple[int, int] = (0, 1), expert_group=None):
        super().__init__()
        self.n_heads = n_heads
        self.dim = dim
        self.norm_eps = norm_eps
        self.attention = Attention(dim=dim, n_heads=n_heads, head_dim=head_dim, n_kv_heads=n_kv_heads, lora=lora)
        self.attention_norm = RMSNorm(dim, eps=norm_eps)
        self.ffn_norm = RMSNorm(dim, eps=norm_eps)
        self.feed_forward: nn.Module
        if moe is not None:
            g, G = expert_shard  # this rank allocates only the experts it owns (e % G == g): SURVEY.md 8(e)
            self.feed_forward = MoeLayer(experts={e: FeedForward(dim=dim, hidden_dim=hidden_dim, lora=lora) for e in range(moe.num_experts) if e % G == g},
                                         gate_weight=nn.Parameter(torch.empty(moe.num_experts, dim), requires_grad=False), moe_args=moe,
                                         expert_shard=expert_shard, expert_group=expert_group)
        else:
            self.feed_forward = FeedForward(dim=dim, hidden_dim=hidden_dim, lora=lora)

    def forward(self, x: torch.Tensor, rope: torch.Tensor, positions: torch.Tensor, cache: Optional[CacheView],
                ws: "_abi.Workspace") -> torch.Tensor:
        # r = attention(attention_norm(x)); h = x + r        (transformer_layers.py:165-166)
        a = self.attention.attend(x, self.attention_norm.weight, self.norm_eps, rope, positions, cache, ws)
        h = torch.empty_like(x)
        _abi.linear_residual(a, self.attention.wo_weight, x, h, ws)
        # r = feed_forward(ffn_norm(h)); out = h + r          (transformer_layers.py:167-168)
        if isinstance(self.feed_forward, MoeLayer):
            hn = _abi.rmsnorm(h, self.ffn_norm.weight, self.norm_eps)
            return self.feed_forward.run(hn, h, ws)  # router + grouped experts + ordered combine + residual, no host sync
        return self.feed_forward.run(h, self.ffn_norm.weight, self.norm_eps, h, ws)
