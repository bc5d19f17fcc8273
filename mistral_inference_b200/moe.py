"""Mixture-of-experts layer (API mirror of mistral_inference/moe.py:16-32).

Batch-1 decode runs the router and the selected experts INSIDE the decode megakernel (csrc/decode_megakernel.cuh).  Everything
else -- prefill and batched decode -- goes through two C-ABI calls per layer with no host synchronisation (csrc/moe.cuh):

  mb200_moe_route        gate GEMV, top-k on the bf16 router logits, fp32 softmax over the k selected (moe.py:25-27), a
                         deterministic expert-sorted row plan, gather of the token rows
  mb200_moe_grouped_ffn  grouped tcgen05 GEMMs over the experts (gate/up + SiLU*mul, down) and the combine:
                         out = h + sum_j bf16(w_j * y_j) taken in ascending expert index, rounded to bf16 at every step exactly
                         like the reference's `results[idx] += w * expert(x)` loop (moe.py:28-31)

Expert parallelism: with `expert_shard = (g, G)` this rank owns the experts `e % G == g` and holds no weights of
the others.  Every rank evaluates the router and the row plan redundantly (deterministic, identical on all ranks), runs the
grouped GEMMs for its own experts, and the DOWN PROJECTION'S EPILOGUE stores each weighted output row into the row buffer of
every rank over NVLink (peer memory mapped with CUDA IPC, `ExpertComm`): the exchange is an all-gather of rows fused into the
GEMM, not a collective call after it.  After a flag handshake every rank combines all k rows of every token in the reference's
order, so the result is bit-identical to the unsharded model for any k -- no reduction order is left to a library.
"""
import ctypes
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import _abi
from .args import MoeArgs

MOE_BLOCK_TOKENS = 16384  # prefill goes through the experts in blocks of this many tokens (bounds the row buffers: ~2 GB at 8x22B)


class _GateView:
    def __init__(self, layer: "MoeLayer"):
        self._layer = layer

    @property
    def weight(self) -> torch.Tensor:
        return self._layer.gate_weight


class MoeBuffers:
    """Row buffers of one MoE call for up to `T` tokens (shared by all layers of a model; sizes from mb200_moe_sizes)."""

    def __init__(self, T: int, dim: int, hidden: int, E: int, k: int, device: torch.device, dtype: torch.dtype, yw_ptr: Optional[int] = None):
        self.T = T
        self.tile_rows, self.rows_cap, self.plan_words = _abi.moe_sizes(T, E, k)
        i32 = dict(dtype=torch.int32, device=device)
        self.sel = torch.empty(T * k, **i32)
        self.slot = torch.empty(T * k, **i32)
        self.plan = torch.zeros(self.plan_words, **i32)
        self.wts = torch.empty(T * k, dtype=dtype, device=device)
        self.row_w = torch.zeros(self.rows_cap, dtype=dtype, device=device)
        self.xs = torch.zeros(self.rows_cap, dim, dtype=dtype, device=device)  # zero once: padded rows stay finite
        self.g = torch.zeros(self.rows_cap, hidden, dtype=dtype, device=device)
        # weighted expert output rows: a local tensor, or (expert parallel) this rank's IPC-exported region that peers write too
        self.yw = torch.zeros(self.rows_cap, dim, dtype=dtype, device=device) if yw_ptr is None else None
        self.yw_ptr = self.yw.data_ptr() if yw_ptr is None else yw_ptr


class ExpertComm:
    """NVLink peer memory of an expert-parallel group: per rank ONE cudaMalloc'ed region exported with CUDA IPC, holding two row
    buffers (consecutive MoE layers alternate, so a rank that runs ahead never overwrites rows a peer is still combining) and
    the handshake flags.  Handles travel through torch.distributed (`all_gather_object`)."""

    FLAG_BYTES = 4096  # [2 parities][n_ranks] uint32, padded

    def __init__(self, rank: int, world: int, group, rows_cap: int, dim: int, device: torch.device):
        import torch.distributed as dist

        self.rank, self.world, self.rows_cap, self.dim = rank, world, rows_cap, dim
        self.buf_bytes = (rows_cap * dim * 2 + 255) & ~255
        self.total = self.FLAG_BYTES + 2 * self.buf_bytes
        self.base = _abi.comm_alloc(self.total)
        handle = _abi.comm_export(self.base)
        handles: List[Optional[bytes]] = [None] * world
        dist.all_gather_object(handles, handle, group=group)
        self.peers: List[int] = []  # mapped base pointers of the other ranks, in rank order
        for r in range(world):
            if r != rank:
                self.peers.append(_abi.comm_open(handles[r]))
        self.state = torch.zeros(2, 64, dtype=torch.int32, device=device)  # per parity: [0] epoch, [32] done counter (separate lines)
        self.calls = 0
        self._structs: Dict[int, "_abi.MoeCommStruct"] = {}
        dist.barrier(group=group)  # nobody writes a peer before every rank has mapped everything

    def yw_ptr(self, parity: int) -> int:
        """Device pointer of this rank's row buffer of `parity` (inside the IPC-exported allocation)."""
        return self.base + self.FLAG_BYTES + parity * self.buf_bytes

    def struct(self, parity: int) -> "_abi.MoeCommStruct":
        s = self._structs.get(parity)
        if s is None:
            s = _abi.MoeCommStruct()
            s.n_ranks, s.my_rank = self.world, self.rank
            for i, pb in enumerate(self.peers):
                s.peer_yw[i] = pb + self.FLAG_BYTES + parity * self.buf_bytes
                s.peer_flags[i] = pb + parity * 4 * 64
            s.my_flags = self.base + parity * 4 * 64
            s.epoch = self.state[parity].data_ptr()
            s.done_counter = self.state[parity].data_ptr() + 32 * 4
            self._structs[parity] = s
        return s

    def close(self) -> None:
        for pb in self.peers:
            _abi.comm_close(pb)
        self.peers = []
        if self.base:
            _abi.comm_free(self.base)
            self.base = 0


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
        self.layer_parity = 0  # set by the model: consecutive MoE layers alternate the exchange buffer
        self._ptrs = None

    @property
    def gate(self) -> _GateView:
        return _GateView(self)

    @property
    def local_expert_ids(self) -> List[int]:
        return sorted(int(e) for e in self.experts.keys())

    @property
    def sharded(self) -> bool:
        return self.expert_shard[1] > 1

    def _weight_tables(self):
        """HOST arrays of E device pointers (NULL for experts of other ranks), rebuilt when a weight moved."""
        E = self.args.num_experts
        key = tuple((e, self.experts[str(e)].w13.data_ptr(), self.experts[str(e)].w2_weight.data_ptr()) for e in self.local_expert_ids)
        if self._ptrs is None or self._ptrs[0] != key:
            w13 = (ctypes.c_void_p * E)()
            w2 = (ctypes.c_void_p * E)()
            for e in self.local_expert_ids:
                w13[e] = self.experts[str(e)].w13.data_ptr()
                w2[e] = self.experts[str(e)].w2_weight.data_ptr()
            self._ptrs = (key, w13, w2)
        return self._ptrs[1], self._ptrs[2]

    def run(self, hn: torch.Tensor, residual: Optional[torch.Tensor], ws: "_abi.Workspace") -> torch.Tensor:
        """`hn` = ffn_norm(h) [T, dim]; returns residual + moe(hn) (or moe(hn) when residual is None)."""
        T, dim = hn.shape
        first = self.experts[str(self.local_expert_ids[0])]
        E, k = self.args.num_experts, self.args.num_experts_per_tok
        out = torch.empty_like(hn)
        w13, w2 = self._weight_tables()
        g, G = self.expert_shard
        for r0 in range(0, T, MOE_BLOCK_TOKENS):
            r1 = min(T, r0 + MOE_BLOCK_TOKENS)
            n = r1 - r0
            comm = ws.expert_comm(self, n, dim) if self.sharded else None
            b = ws.moe_buffers(n, dim, first.hidden_dim, E, k, hn.dtype, comm, self.layer_parity)
            assert comm is None or b.rows_cap <= comm.rows_cap
            _abi.moe_route(hn[r0:r1], self.gate_weight, E, k, g, G, b)
            _abi.moe_grouped_ffn(b, w13, w2, residual[r0:r1] if residual is not None else None, out[r0:r1], n, dim, first.hidden_dim, E, k,
                                 comm.struct(self.layer_parity) if comm is not None else None, ws)
        return out

    def forward(self, inputs: torch.Tensor, ws: Optional["_abi.Workspace"] = None) -> torch.Tensor:
        """`inputs` = ffn_norm(h) [T, dim] (already normed, like the reference's MoeLayer.forward); returns the expert mixture."""
        T, dim = inputs.shape
        first = self.experts[str(self.local_expert_ids[0])]
        ws = ws or _abi.Workspace(_abi.workspace_bytes(T, dim, 1, 1, 128, first.hidden_dim, 0, 1), inputs.device)
        return self.run(inputs, None, ws)
