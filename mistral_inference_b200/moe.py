"""Mixture-of-experts layer (API mirror of mistral_inference/moe.py:16-32).

Round-1 status: routing bookkeeping (top-k on the bf16 router logits, fp32 softmax over the k selected,
ascending-expert bf16 `+=` accumulation, moe.py:25-31) is host-orchestrated like the reference, with all
GEMMs (router, gate/up + SiLU*mul, down) running in libmb200.  The fused router and the grouped expert
kernel replace the Python loop next.
"""
from typing import List, Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import _abi
from .args import MoeArgs


class _GateView:
    def __init__(self, layer: "MoeLayer"):
        self._layer = layer

    @property
    def weight(self) -> torch.Tensor:
        return self._layer.gate_weight


class MoeLayer(nn.Module):
    def __init__(self, experts: List[nn.Module], gate_weight: nn.Parameter, moe_args: MoeArgs):
        super().__init__()
        assert len(experts) > 0
        self.experts = nn.ModuleList(experts)
        self.gate_weight = gate_weight  # [E, dim]
        self.args = moe_args

    @property
    def gate(self) -> _GateView:
        return _GateView(self)

    def forward(self, inputs: torch.Tensor, ws: Optional["_abi.Workspace"] = None) -> torch.Tensor:
        """`inputs` = ffn_norm(h) [T, dim] (already normed, like the reference's MoeLayer.forward)."""
        T, dim = inputs.shape
        ws = ws or _abi.Workspace(_abi.workspace_bytes(T, dim, 1, 1, 128, self.experts[0].hidden_dim, 0, 1), inputs.device)
        gate_logits = torch.empty(T, self.args.num_experts, dtype=inputs.dtype, device=inputs.device)
        _abi.linear_residual(inputs, self.gate_weight, None, gate_logits, ws)
        weights, selected_experts = torch.topk(gate_logits, self.args.num_experts_per_tok)
        weights = F.softmax(weights, dim=1, dtype=torch.float).to(inputs.dtype)
        results = torch.zeros_like(inputs)
        for i, expert in enumerate(self.experts):
            batch_idx, nth_expert = torch.where(selected_experts == i)
            if batch_idx.numel() == 0:
                continue
            y = expert.run(inputs[batch_idx].contiguous(), None, 0.0, None, ws)
            results[batch_idx] += weights[batch_idx, nth_expert, None] * y
        return results
