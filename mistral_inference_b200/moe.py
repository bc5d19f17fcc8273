"""Mixture-of-experts layer (API mirror of mistral_inference/moe.py:16-32).

Batch-1 decode runs the router and the selected experts INSIDE the decode megakernel (csrc/decode_megakernel.cuh).  This module
is the path for prefill and batch > 1: the routing bookkeeping (top-k on the bf16 router logits, fp32 softmax over the k selected,
ascending-expert bf16 `+=` accumulation, moe.py:25-31) is host-orchestrated like the reference, with all GEMMs (router, gate/up +
SiLU*mul, down) running in libmb200.

Expert sharding: with `expert_shard = (g, G)` this rank owns the experts `e % G == g`, holds no weights of the
others, evaluates the router redundantly (deterministic), runs its local experts on the tokens routed to them and contributes
its weighted partial output to ONE all-reduce(sum) of `[T, dim]` per MoE layer.  With top-2 routing every token has exactly two
non-zero contributions, each already rounded to bf16 (`w * y`), so any reduction order yields `bf16(a + b)`: bit-identical to
the reference's ordered `+=` (moe.py:31).
"""
from typing import Callable, Dict, Iterable, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from . import _abi
from .args import MoeArgs


def all_reduce_partial(results: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of the per-rank partial MoE outputs.  NCCL reduces bf16 natively; other backends (gloo in the CPU tests) go through
    fp32 on the wire, which rounds the two-term sum once, exactly like a bf16 add."""
    import torch.distributed as dist

    if dist.get_backend(group) == "nccl":
        dist.all_reduce(results, op=dist.ReduceOp.SUM, group=group)
        return results
    wide = results.float()
    dist.all_reduce(wide, op=dist.ReduceOp.SUM, group=group)
    return wide.to(results.dtype)


def route_and_combine(inputs: torch.Tensor, gate_logits: torch.Tensor, num_experts_per_tok: int, local_experts: Iterable[int],
                      expert_fn: Callable[[int, torch.Tensor], torch.Tensor],
                      reduce_fn: Optional[Callable[[torch.Tensor], torch.Tensor]] = None) -> torch.Tensor:
    """moe.py:25-32 restricted to `local_experts` (all of them when unsharded), then the cross-rank sum if `reduce_fn` is given.
    Pure host logic over tensors: the CPU tests drive it with the oracle's FeedForward as `expert_fn`."""
    weights, selected_experts = torch.topk(gate_logits, num_experts_per_tok)
    weights = F.softmax(weights, dim=1, dtype=torch.float).to(inputs.dtype)
    results = torch.zeros_like(inputs)
    for e in sorted(local_experts):  # ascending expert index: the order the reference's `+=` runs in
        batch_idx, nth_expert = torch.where(selected_experts == e)
        if batch_idx.numel() == 0:
            continue
        y = expert_fn(e, inputs[batch_idx].contiguous())
        results[batch_idx] += weights[batch_idx, nth_expert, None] * y
    return reduce_fn(results) if reduce_fn is not None else results


class _GateView:
    def __init__(self, layer: "MoeLayer"):
        self._layer = layer

    @property
    def weight(self) -> torch.Tensor:
        return self._layer.gate_weight


class MoeLayer(nn.Module):
    def __init__(self, experts: Dict[int, nn.Module], gate_weight: nn.Parameter, moe_args: MoeArgs,
                 expert_shard: Tuple[int, int] = (0, 1), expert_group=None):
        """`experts`: the LOCAL experts keyed by global expert id (all of them when unsharded; a list is accepted too)."""
        super().__init__()
        if not isinstance(experts, dict):
            experts = dict(enumerate(experts))
        assert len(experts) > 0
        # keyed by the GLOBAL expert id so that the reference's key names `experts.{e}.w1.weight` stay valid on every rank
        self.experts = nn.ModuleDict({str(e): m for e, m in sorted(experts.items())})
        self.gate_weight = gate_weight  # [E, dim]
        self.args = moe_args
        self.expert_shard = expert_shard
        self.expert_group = expert_group

    @property
    def gate(self) -> _GateView:
        return _GateView(self)

    @property
    def local_expert_ids(self) -> List[int]:
        return sorted(int(e) for e in self.experts.keys())

    @property
    def sharded(self) -> bool:
        return self.expert_shard[1] > 1

    def forward(self, inputs: torch.Tensor, ws: Optional["_abi.Workspace"] = None) -> torch.Tensor:
        """`inputs` = ffn_norm(h) [T, dim] (already normed, like the reference's MoeLayer.forward)."""
        T, dim = inputs.shape
        first = self.experts[str(self.local_expert_ids[0])]
        ws = ws or _abi.Workspace(_abi.workspace_bytes(T, dim, 1, 1, 128, first.hidden_dim, 0, 1), inputs.device)
        gate_logits = torch.empty(T, self.args.num_experts, dtype=inputs.dtype, device=inputs.device)
        _abi.linear_residual(inputs, self.gate_weight, None, gate_logits, ws)
        return route_and_combine(
            inputs, gate_logits, self.args.num_experts_per_tok, self.local_expert_ids,
            lambda e, x: self.experts[str(e)].run(x, None, 0.0, None, ws),
            (lambda r: all_reduce_partial(r, self.expert_group)) if self.sharded else None)
