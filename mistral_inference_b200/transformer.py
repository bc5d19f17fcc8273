"""Transformer model shell (API mirror of mistral_inference/transformer.py).

`Transformer.from_folder / forward / forward_partial / load_state_dict` keep the reference's signatures and
on-disk contract (params.json + consolidated.safetensors | consolidated.00.pth, reference state-dict keys);
the layer loop calls the fused libmb200 kernels.  bf16 on CUDA only; there is no PyTorch/CPU fallback.
"""
import json
import logging
import math
import os
from pathlib import Path
from typing import Any, Dict, List, Mapping, Optional, Tuple, Union

import torch
from torch import nn

from . import _abi
from .args import TransformerArgs
from .cache import BufferCache, CacheInputMetadata
from .rope import precompute_freqs_cis
from .transformer_layers import RMSNorm, TransformerBlock

ROPE_TABLE_LEN = 128_000  # transformer.py:116
_NVTX = os.environ.get("MB200_NVTX", "0") == "1"


class _nvtx:
    """NVTX range around a phase of the forward (MB200_NVTX=1; visible in nsys / ncu --nvtx): the layer loop, the decode step."""

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _NVTX:
            torch.cuda.nvtx.range_pop()


class _OutputView:
    def __init__(self, model: "Transformer"):
        self._m = model

    @property
    def weight(self) -> torch.Tensor:
        return self._m.output_weight


class Transformer(nn.Module):
    def __init__(self, args: TransformerArgs, pipeline_rank: int = 0, num_pipeline_ranks: int = 1, softmax_fp32: bool = True,
                 expert_parallel: Optional[Tuple[int, int]] = None, expert_group: Any = None):
        """Same signature as the reference (transformer.py:34-40) plus `expert_parallel = (rank, world)`: MoE experts sharded
        `e % world == rank` over the ranks of `expert_group` (default process group), everything else replicated, one
        all-reduce of [T, dim] per MoE layer."""
        super().__init__()
        self.args = args
        self.expert_parallel = expert_parallel or (0, 1)
        assert 0 <= self.expert_parallel[0] < self.expert_parallel[1], self.expert_parallel
        assert self.expert_parallel[1] == 1 or (args.moe is not None and num_pipeline_ranks == 1), \
            "expert sharding needs a MoE model and excludes pipeline ranks"
        self.vocab_size = args.vocab_size
        self.n_layers = args.n_layers
        self._rope_table: Optional[torch.Tensor] = None
        assert self.vocab_size > 0
        assert pipeline_rank < num_pipeline_ranks, (pipeline_rank, num_pipeline_ranks)
        self.pipeline_rank = pipeline_rank
        self.num_pipeline_ranks = num_pipeline_ranks
        self.softmax_fp32 = softmax_fp32

        self.tok_embeddings: Optional[nn.Embedding] = None
        self.norm: Optional[RMSNorm] = None
        self.output_weight: Optional[nn.Parameter] = None
        if pipeline_rank == 0:
            self.tok_embeddings = nn.Embedding(args.vocab_size, args.dim)
            self.tok_embeddings.weight.requires_grad_(False)
        if pipeline_rank == num_pipeline_ranks - 1:
            self.norm = RMSNorm(args.dim, eps=args.norm_eps)
            self.output_weight = nn.Parameter(torch.empty(args.vocab_size, args.dim), requires_grad=False)
        # contiguous layer ranges per pipeline rank, keyed by GLOBAL layer id (transformer.py:94-98)
        num_layers_per_rank = math.ceil(self.n_layers / self.num_pipeline_ranks)
        offset = self.pipeline_rank * num_layers_per_rank
        end = min(self.n_layers, offset + num_layers_per_rank)
        self.layers = nn.ModuleDict({
            str(i): TransformerBlock(dim=args.dim, hidden_dim=args.hidden_dim, n_heads=args.n_heads, n_kv_heads=args.n_kv_heads,
                                     head_dim=args.head_dim, norm_eps=args.norm_eps, lora=args.lora, moe=args.moe,
                                     expert_shard=self.expert_parallel, expert_group=expert_group)
            for i in range(offset, end)
        })
        self.n_local_layers = len(self.layers)
        for j, blk in enumerate(self.layers.values()):  # consecutive MoE layers alternate the expert-parallel exchange buffer
            if hasattr(blk.feed_forward, "layer_parity"):
                blk.feed_forward.layer_parity = j & 1
        self._ws: Optional[_abi.Workspace] = None
        self._ws_tokens = 0
        self.last_argmax: Optional[torch.Tensor] = None  # device token id(s) written by the last fused-argmax decode step
        self._last_static_logits = 0                     # data_ptr of the logits buffer that argmax belongs to

    # ------------------------------------------------------------------ properties
    @property
    def dtype(self) -> torch.dtype:
        return next(self.parameters()).dtype

    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    @property
    def output(self) -> _OutputView:
        return _OutputView(self)

    @property
    def freqs_cis(self) -> torch.Tensor:
        """complex64 [128000, hd/2] on the model device (transformer.py:108-120)."""
        return torch.view_as_complex(self.rope_table)

    @property
    def rope_table(self) -> torch.Tensor:
        """fp32 [128000, hd/2, 2] (cos, sin): the same bits as the reference's table, built on the CPU."""
        if self._rope_table is None:
            theta = self.args.rope_theta or 1000000.0
            self._rope_table = torch.view_as_real(precompute_freqs_cis(self.args.head_dim, ROPE_TABLE_LEN, theta)).contiguous()
        if self._rope_table.device != self.device:
            self._rope_table = self._rope_table.to(device=self.device)
        return self._rope_table

    def workspace(self, num_tokens: int) -> _abi.Workspace:
        if self._ws is None or self._ws_tokens < num_tokens or self._ws.buf.device != self.device:
            a = self.args
            need = _abi.workspace_bytes(num_tokens, a.dim, a.n_heads, a.n_kv_heads, a.head_dim, a.hidden_dim, a.vocab_size,
                                        max(a.max_batch_size, 1))
            self._ws = _abi.Workspace(need, self.device)  # decode states remember the pointer they captured (see _decode_state)
            self._ws_tokens = num_tokens
        return self._ws

    # ------------------------------------------------------------------ forward
    def _check_runnable(self) -> None:
        if self.device.type != "cuda" or self.dtype != torch.bfloat16:
            raise _abi.Mb200Error(f"the libmb200 hot path runs bf16 on CUDA only (got {self.dtype} on {self.device}); no fallback exists")

    @torch.inference_mode()
    def forward_partial(self, input_ids: torch.Tensor, seqlens: List[int], cache: Optional[BufferCache] = None,
                        images: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
        """Local forward pass (transformer.py:163-219): hidden states of this stage; the last stage returns
        the normalised final embeddings."""
        assert not images, "vision inputs are outside the accelerated hot path"
        self._check_runnable()
        assert len(seqlens) <= self.args.max_batch_size, f"Max batch size is {self.args.max_batch_size}, got batch size of {len(seqlens)}"
        (num_toks,) = input_ids.shape
        assert sum(seqlens) == num_toks, (sum(seqlens), num_toks)
        ws = self.workspace(num_toks)

        input_metadata: Optional[List[CacheInputMetadata]] = None
        if cache is not None:
            self._check_positions(cache, seqlens)
            input_metadata = cache.get_input_metadata(seqlens)
            positions = input_metadata[0].positions
        else:
            positions = torch.cat([torch.arange(0, s, dtype=torch.int32) for s in seqlens]).to(self.device)

        if self.pipeline_rank == 0:
            assert self.tok_embeddings is not None
            h = self.tok_embeddings(input_ids)
        else:
            h = torch.empty(num_toks, self.args.dim, device=self.device, dtype=self.dtype)
            torch.distributed.recv(h, src=self.pipeline_rank - 1)

        rope = self.rope_table
        for local_layer_id, layer in enumerate(self.layers.values()):
            view = cache.get_view(local_layer_id, input_metadata[local_layer_id]) if cache is not None else None
            h = layer(h, rope, positions, view, ws)

        if cache is not None:
            cache.update_seqlens(seqlens)
        if self.pipeline_rank < self.num_pipeline_ranks - 1:
            torch.distributed.send(h, dst=self.pipeline_rank + 1)
            return h
        assert self.norm is not None
        return self.norm(h)

    @torch.inference_mode()
    def forward(self, input_ids: torch.Tensor, seqlens: List[int], cache: Optional[BufferCache] = None,
                images: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
        """transformer.py:221-242.  [T, vocab] logits, fp32 when softmax_fp32."""
        assert not images, "vision inputs are outside the accelerated hot path"
        self._check_runnable()
        last = self.pipeline_rank == self.num_pipeline_ranks - 1
        if last and self.num_pipeline_ranks == 1:
            if self._graph_decode_ok(seqlens, cache):
                outs32 = self.decode_static(input_ids, cache).clone()
                return outs32 if self.softmax_fp32 else outs32.to(self.dtype)
            # single stage: final norm + lm head + .float() are one fused call (no [T, dim] normed round trip)
            h = self._hidden_no_norm(input_ids, seqlens, cache)
            if cache is not None:
                cache.update_seqlens(seqlens)
            outs32 = torch.empty(h.shape[0], self.vocab_size, device=h.device, dtype=torch.float32)
            assert self.norm is not None and self.output_weight is not None
            _abi.lm_head(h, self.norm.weight, self.output_weight, outs32, self.args.norm_eps, self.workspace(h.shape[0]))
            return outs32 if self.softmax_fp32 else outs32.to(self.dtype)
        h = self.forward_partial(input_ids, seqlens, cache=cache)
        if not last:
            outs = torch.empty(h.shape[0], self.vocab_size, device=h.device, dtype=h.dtype)
        else:
            assert self.output_weight is not None
            outs = torch.empty(h.shape[0], self.vocab_size, device=h.device, dtype=h.dtype)
            _abi.linear_residual(h, self.output_weight, None, outs, self.workspace(h.shape[0]))
        if self.num_pipeline_ranks > 1:
            torch.distributed.broadcast(outs, src=self.num_pipeline_ranks - 1)
        return outs.float() if self.softmax_fp32 else outs

    def _hidden_no_norm(self, input_ids: torch.Tensor, seqlens: List[int], cache: Optional[BufferCache],
                        input_metadata: Optional[List[CacheInputMetadata]] = None) -> torch.Tensor:
        assert len(seqlens) <= self.args.max_batch_size, f"Max batch size is {self.args.max_batch_size}, got batch size of {len(seqlens)}"
        (num_toks,) = input_ids.shape
        assert sum(seqlens) == num_toks, (sum(seqlens), num_toks)
        ws = self.workspace(num_toks)
        if cache is not None:
            if input_metadata is None:
                self._check_positions(cache, seqlens)
                input_metadata = cache.get_input_metadata(seqlens)
            positions = input_metadata[0].positions
        else:
            positions = torch.cat([torch.arange(0, s, dtype=torch.int32) for s in seqlens]).to(self.device)
        assert self.tok_embeddings is not None
        h = self.tok_embeddings(input_ids)
        rope = self.rope_table
        with _nvtx(f"mb200.layers[T={num_toks}]"):
            for local_layer_id, layer in enumerate(self.layers.values()):
                view = cache.get_view(local_layer_id, input_metadata[local_layer_id]) if cache is not None else None
                h = layer(h, rope, positions, view, ws)
        return h

    def _check_positions(self, cache: BufferCache, seqlens: List[int]) -> None:
        """The reference indexes freqs_cis[positions] and raises past the table (transformer.py:199); the kernels would read out of bounds."""
        host = cache._kv_seqlens_host or [0] * len(seqlens)
        last = max(p + s for p, s in zip(host, seqlens)) if seqlens else 0
        if last > ROPE_TABLE_LEN:
            raise IndexError(f"position {last - 1} is out of bounds for the rope table of {ROPE_TABLE_LEN} positions")

    # ------------------------------------------------------------------ CUDA-graph decode
    def _graph_decode_ok(self, seqlens: List[int], cache: Optional[BufferCache]) -> bool:
        if cache is None or self.num_pipeline_ranks != 1:
            return False
        if os.environ.get("MB200_DECODE_GRAPH", "1") == "0":
            return False
        host = cache._kv_seqlens_host
        return (host is not None and len(host) == len(seqlens) and host[0] != 0 and all(s == 1 for s in seqlens)
                and len(set(cache.cache_sizes)) <= 8)

    def _megakernel_ok(self, B: int) -> bool:
        return (B == 1 and self.num_pipeline_ranks == 1 and self.expert_parallel[1] == 1 and self.args.n_kv_heads <= 8
                and os.environ.get("MB200_MEGAKERNEL", "1") != "0"
                and (self.args.moe is None or (self.args.moe.num_experts <= 32 and self.args.moe.num_experts_per_tok <= 4)))

    def _decode_state(self, cache: BufferCache, key: Any) -> Dict[str, Any]:
        """Per-(cache, kind) decode state (descriptor tables, static buffers, captured graph).  It lives ON THE CACHE OBJECT, so
        it is freed with the cache (generate() allocates a cache per call); an entry is rebuilt when this model's workspace was
        reallocated since (a captured graph holds the old pointer)."""
        states = cache.__dict__.setdefault("_mb200_decode_states", {})
        ws_ptr = self._ws.ptr if self._ws is not None else 0
        st = states.get((id(self), key))
        if st is None or st["ws_ptr"] != ws_ptr or st["device"] != self.device:
            st = {"ws_ptr": ws_ptr, "device": self.device}
            states[(id(self), key)] = st
        return st

    @torch.inference_mode()
    def _decode_megakernel(self, tokens: torch.Tensor, cache: BufferCache) -> torch.Tensor:
        """Batch-1 decode step as ONE persistent cooperative kernel (csrc/decode_megakernel.cuh)."""
        import numpy as np

        a = self.args
        ws = self.workspace(1)
        st = self._decode_state(cache, "mk")
        if "layers" not in st:
            blocks = list(self.layers.values())
            desc = np.zeros((len(blocks), 8), dtype=np.uint64)
            moe = self.args.moe
            E = moe.num_experts if moe is not None else 0
            gate_tab, w13_tab, w2_tab = [], [], []
            for i, blk in enumerate(blocks):
                ff = blk.feed_forward
                if moe is None:
                    w13p, w2p = ff.w13.data_ptr(), ff.w2_weight.data_ptr()
                else:  # the kernel reads the expert tables instead
                    w13p = w2p = 0
                    gate_tab.append(ff.gate_weight.data_ptr())
                    w13_tab += [ff.experts[str(e)].w13.data_ptr() for e in range(E)]
                    w2_tab += [ff.experts[str(e)].w2_weight.data_ptr() for e in range(E)]
                desc[i] = [blk.attention.wqkv.data_ptr(), blk.attention.wo_weight.data_ptr(), w13p, w2p,
                           blk.attention_norm.weight.data_ptr(), blk.ffn_norm.weight.data_ptr(), cache.cache_k[i].data_ptr(),
                           cache.cache_v[i].data_ptr()]
            tab = lambda v: torch.tensor(v, dtype=torch.int64, device=self.device) if v else None  # noqa: E731
            st.update({"E": E, "k": moe.num_experts_per_tok if moe is not None else 0,
                       "moe_gate": tab(gate_tab), "moe_w13": tab(w13_tab), "moe_w2": tab(w2_tab),
                       "layers": torch.from_numpy(desc.view(np.int64)).to(self.device),
                       "windows": torch.tensor(cache.cache_sizes, dtype=torch.int32, device=self.device),
                       "token": torch.zeros(1, dtype=torch.long, device=self.device),
                       "next": torch.zeros(1, dtype=torch.long, device=self.device),
                       "logits": torch.empty(1, self.vocab_size, dtype=torch.float32, device=self.device)})
        if cache._kv_seqlens_host is None:
            cache.init_kvseqlens(1)
        pos = cache._kv_seqlens_host[0]
        if pos >= ROPE_TABLE_LEN:
            raise IndexError(f"position {pos} is out of bounds for the rope table of {ROPE_TABLE_LEN} positions")
        # `tokens` may be the previous step's fused argmax (st["next"]): then nothing is copied and the greedy loop is one
        # kernel launch per token
        tok = tokens.reshape(1)
        if tok.data_ptr() != st["next"].data_ptr():
            st["token"].copy_(tok, non_blocking=True)
            tok = st["token"]
        self.last_argmax = st["next"]
        with _nvtx("mb200.decode_megakernel"):
            _abi.decode_step(st["layers"], st["windows"], self.n_local_layers, self.tok_embeddings.weight, self.norm.weight, self.output_weight,
                             self.rope_table, tok, pos, 0, st["logits"], st["next"], a.dim, a.hidden_dim, a.n_heads, a.n_kv_heads, a.head_dim,
                             self.vocab_size, a.norm_eps, ws, st["E"], st["k"], st["moe_gate"], st["moe_w13"], st["moe_w2"])
        cache.update_seqlens([1])
        self._last_static_logits = st["logits"].data_ptr()
        return st["logits"]

    @torch.inference_mode()
    def decode_static(self, tokens: torch.Tensor, cache: BufferCache) -> torch.Tensor:
        """One decode step for every sequence of `cache` (one new token each).  Batch 1: the persistent megakernel.  Batch > 1:
        the per-layer kernels replayed from a CUDA graph.  The step state lives on the DEVICE: `mb200_decode_meta` derives
        positions / ring rows / kv lengths from a device-side position vector and advances it inside the graph, so a replay
        needs no host write at all (a pinned staging buffer rewritten by the host while earlier copies are still queued was the
        round-1 design and a race).  Returns the STATIC fp32 logits buffer [B, V] (overwritten by the next step).  The first call
        per (cache, batch) runs eagerly (warm-up: loads modules, sets function attributes), the second captures."""
        B = tokens.shape[0]
        if self._megakernel_ok(B):
            return self._decode_megakernel(tokens, cache)
        seqlens = [1] * B
        self.workspace(B)
        st = self._decode_state(cache, ("graph", B))
        host = cache._kv_seqlens_host
        assert host is not None and len(host) == B, "decode_static needs a prefilled cache of this batch size"
        if max(host) >= ROPE_TABLE_LEN:
            raise IndexError(f"position {max(host)} is out of bounds for the rope table of {ROPE_TABLE_LEN} positions")
        distinct = sorted(set(cache.cache_sizes))
        if "meta" not in st:
            st.update({"graph": None, "warmed": False, "expected": None,
                       "seqpos": torch.zeros(B, dtype=torch.int32, device=self.device),
                       "tokens": torch.zeros(B, dtype=torch.long, device=self.device),
                       "meta": torch.zeros(3 * B + 1 + 2 * B * len(distinct), dtype=torch.int32, device=self.device),
                       "logits": torch.empty(B, self.vocab_size, dtype=torch.float32, device=self.device),
                       "next": torch.zeros(B, dtype=torch.long, device=self.device)})
        if st["expected"] != host:  # another forward() advanced the cache since the last step (or this is the first one)
            st["seqpos"].copy_(torch.tensor(host, dtype=torch.int32))  # pageable source: staged by the runtime, no reuse hazard
        st["tokens"].copy_(tokens, non_blocking=True)  # device source (previous pick): D2D; host source: staged by the runtime
        layout = {"T": B, "B": B, "prefill": False, "first_prefill": False, "max_seqlen": 1, "windows": distinct}
        md = cache.metadata_from_block(st["meta"], layout, seqlens)

        def run() -> None:
            _abi.decode_meta(st["seqpos"], st["meta"], distinct)
            h = self._hidden_no_norm(st["tokens"], seqlens, cache, md)
            _abi.lm_head(h, self.norm.weight, self.output_weight, st["logits"], self.args.norm_eps, self.workspace(B))
            _abi.argmax_rows(st["logits"], st["next"])  # greedy pick on the device (generate.py:156): feeds the next step

        if st["graph"] is None and not st["warmed"]:
            run()
            st["warmed"] = True
        elif st["graph"] is None:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                run()
            st["graph"] = g
            g.replay()
        else:
            st["graph"].replay()
        cache.update_seqlens(seqlens)
        st["expected"] = list(cache._kv_seqlens_host)
        self.last_argmax = st["next"]
        self._last_static_logits = st["logits"].data_ptr()
        return st["logits"]

    # ------------------------------------------------------------------ generate() support
    def last_argmax_valid_for(self, logits: torch.Tensor) -> bool:
        """True when `last_argmax` is the decode kernel's own argmax of exactly this logits buffer."""
        return self.last_argmax is not None and logits.data_ptr() == self._last_static_logits

    @torch.inference_mode()
    def next_token_logits(self, tokens: torch.Tensor, cache: BufferCache) -> torch.Tensor:
        """fp32 logits [B, V] of one decode step for every sequence.  On the single-stage CUDA path this is the step's static
        buffer (no clone; valid until the next step), otherwise `forward`."""
        B = tokens.shape[0]
        if self.pipeline_rank == self.num_pipeline_ranks - 1 == 0 and self._graph_decode_ok([1] * B, cache):
            self._check_runnable()
            return self.decode_static(tokens, cache)
        self._last_static_logits = 0
        out = self.forward(tokens, [1] * B, cache)
        return out if out.dtype == torch.float32 else out.float()

    @torch.inference_mode()
    def forward_logprobs(self, input_ids: torch.Tensor, seqlens: List[int], cache: Optional[BufferCache],
                         targets: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """One prompt chunk for generate(): returns (lp [T] fp32 with lp[t] = log_softmax(logits[t])[targets[t]] where
        targets[t] >= 0, logits [B, V] fp32 of each sequence's last token).  Replaces forward + log_softmax over [T, V] +
        per-token gathers (generate.py:97-118): the lm head runs over blocks of rows, each followed by the fused
        log-softmax + gather kernel, so the full [T, V] logits never exist."""
        self._last_static_logits = 0
        T, V = input_ids.shape[0], self.vocab_size
        last_idx = torch.tensor(seqlens, device=input_ids.device).cumsum(0) - 1
        lp = torch.zeros(T, dtype=torch.float32, device=input_ids.device)
        if self.num_pipeline_ranks > 1:  # reference-compatible pipeline mode: logits arrive by broadcast (transformer.py:236-237)
            logits = self.forward(input_ids, seqlens, cache).float().contiguous()
            _abi.logprob_gather(logits, targets, out=lp)
            return lp, logits.index_select(0, last_idx)
        self._check_runnable()
        h = self._hidden_no_norm(input_ids, seqlens, cache)
        if cache is not None:
            cache.update_seqlens(seqlens)
        assert self.norm is not None and self.output_weight is not None
        rows = max(128, (256 << 20) // (4 * V))
        ws = self.workspace(max(T, 1))
        block = torch.empty(min(rows, T), V, dtype=torch.float32, device=h.device)
        for r0 in range(0, T, rows):
            r1 = min(T, r0 + rows)
            _abi.lm_head(h[r0:r1], self.norm.weight, self.output_weight, block[: r1 - r0], self.args.norm_eps, ws)
            _abi.logprob_gather(block[: r1 - r0], targets[r0:r1], out=lp[r0:r1])
        last_logits = torch.empty(len(seqlens), V, dtype=torch.float32, device=h.device)
        _abi.lm_head(h.index_select(0, last_idx), self.norm.weight, self.output_weight, last_logits, self.args.norm_eps, ws)
        return lp, last_logits

    # ------------------------------------------------------------------ weights
    def _assign(self, k: str, v: torch.Tensor) -> bool:
        """Copies reference-keyed tensor `v` into the packed parameters.  Returns False when the key belongs to
        another pipeline rank."""
        def put(dst: torch.Tensor) -> None:
            assert dst.shape == v.shape, f"{k}: shape {tuple(v.shape)} != expected {tuple(dst.shape)}"
            dst.copy_(v)

        if k == "tok_embeddings.weight":
            if self.tok_embeddings is None:
                return False
            put(self.tok_embeddings.weight)
        elif k == "norm.weight":
            if self.norm is None:
                return False
            put(self.norm.weight)
        elif k == "output.weight":
            if self.output_weight is None:
                return False
            put(self.output_weight)
        elif k.startswith("layers."):
            _, lid, rest = k.split(".", 2)
            if lid not in self.layers:
                return False
            blk: TransformerBlock = self.layers[lid]  # type: ignore[assignment]
            att = blk.attention
            if rest == "attention.wq.weight":
                put(att.wqkv[: att.q_dim])
            elif rest == "attention.wk.weight":
                put(att.wqkv[att.q_dim: att.q_dim + att.kv_dim])
            elif rest == "attention.wv.weight":
                put(att.wqkv[att.q_dim + att.kv_dim:])
            elif rest == "attention.wo.weight":
                put(att.wo_weight)
            elif rest == "attention_norm.weight":
                put(blk.attention_norm.weight)
            elif rest == "ffn_norm.weight":
                put(blk.ffn_norm.weight)
            elif rest == "feed_forward.gate.weight":
                put(blk.feed_forward.gate_weight)
            else:
                parts = rest.split(".")
                if parts[0] != "feed_forward":
                    raise ValueError(f"Unexpected key {k}")
                ff = blk.feed_forward
                if parts[1] == "experts":
                    if parts[2] not in ff.experts:
                        return False  # an expert owned by another expert-parallel rank
                    ff = ff.experts[parts[2]]
                    parts = parts[2:]
                name = parts[1]
                if name == "w1":
                    put(ff.w13.view(ff.hidden_dim, 2, ff.dim)[:, 0])
                elif name == "w3":
                    put(ff.w13.view(ff.hidden_dim, 2, ff.dim)[:, 1])
                elif name == "w2":
                    put(ff.w2_weight)
                else:
                    raise ValueError(f"Unexpected key {k}")
        else:
            raise ValueError(f"Unexpected key {k}")
        return True

    def _owns_key(self, k: str) -> bool:
        """False for checkpoint tensors that belong to another pipeline / expert-parallel rank (decided from the key alone, so a
        loader can skip reading them)."""
        if not k.startswith("layers."):
            return True
        _, lid, rest = k.split(".", 2)
        if lid not in self.layers:
            return False
        parts = rest.split(".")
        if len(parts) > 2 and parts[0] == "feed_forward" and parts[1] == "experts":
            return parts[2] in self.layers[lid].feed_forward.experts
        return True

    def load_state_dict(self, state_dict: Mapping[str, Any], strict: bool = True, assign: bool = False) -> None:  # type: ignore[override]
        """Takes a REFERENCE-keyed state dict (transformer.py:244-295), filters by pipeline rank and packs."""
        del assign  # tensors are copied into the packed buffers
        loaded = set()
        with torch.no_grad():
            for k, v in state_dict.items():
                if self._assign(k, v):
                    loaded.add(k)
                else:
                    logging.debug("Skipping parameter %s at pipeline rank %d", k, self.pipeline_rank)
        if strict:
            missing = set(self.reference_keys()) - loaded
            assert not missing, f"missing keys: {sorted(missing)[:8]}"

    def reference_keys(self) -> List[str]:
        return list(self.state_dict().keys())

    def state_dict(self, *args: Any, **kwargs: Any) -> Dict[str, torch.Tensor]:  # type: ignore[override]
        """Reference-keyed views of the packed parameters."""
        out: Dict[str, torch.Tensor] = {}
        if self.tok_embeddings is not None:
            out["tok_embeddings.weight"] = self.tok_embeddings.weight
        for lid, blk in self.layers.items():
            p = f"layers.{lid}."
            att = blk.attention
            out[p + "attention.wq.weight"] = att.wq.weight
            out[p + "attention.wk.weight"] = att.wk.weight
            out[p + "attention.wv.weight"] = att.wv.weight
            out[p + "attention.wo.weight"] = att.wo_weight
            out[p + "attention_norm.weight"] = blk.attention_norm.weight
            out[p + "ffn_norm.weight"] = blk.ffn_norm.weight
            ff = blk.feed_forward
            if hasattr(ff, "experts"):
                out[p + "feed_forward.gate.weight"] = ff.gate_weight
                for e, ex in ff.experts.items():  # keyed by the global expert id; the local ones only when sharded
                    for n in ("w1", "w2", "w3"):
                        out[p + f"feed_forward.experts.{e}.{n}.weight"] = getattr(ex, n).weight
            else:
                for n in ("w1", "w2", "w3"):
                    out[p + f"feed_forward.{n}.weight"] = getattr(ff, n).weight
        if self.norm is not None:
            out["norm.weight"] = self.norm.weight
            out["output.weight"] = self.output_weight
        return out

    # ------------------------------------------------------------------ LoRA, merged path (lora.py:92-155 with args.lora is None)
    def load_lora(self, lora_path: Union[Path, str], scaling: float = 2.0) -> None:
        """Loads a LoRA checkpoint and MERGES it into the packed weights (lora.py:93-101,120-139): the forward path is unchanged."""
        import safetensors.torch

        lora_path = Path(lora_path)
        assert lora_path.is_file(), f"{lora_path} does not exist or is not a file"
        self._load_lora_state_dict(safetensors.torch.load_file(str(lora_path)), scaling=scaling)

    def _load_lora_state_dict(self, lora_state_dict: Dict[str, torch.Tensor], scaling: float = 2.0) -> None:
        """weight <- weight + (lora_B @ lora_A) * scaling for every Linear of this rank except the output layer, with the same
        torch ops and dtype as the reference (lora.py:129-137).  Un-merged adapters (args.lora set) are outside the hot path."""
        lora_dtypes = set(p.dtype for p in lora_state_dict.values())
        assert len(lora_dtypes) == 1, f"LoRA weights have multiple different dtypes {lora_dtypes}. All weights need to have the same dtype"
        lora_dtype = lora_dtypes.pop()
        assert lora_dtype == self.dtype, f"LoRA weights dtype differs from model's dtype {lora_dtype} != {self.dtype}"
        assert all("lora" in key for key in lora_state_dict.keys())
        assert self.args.lora is None, "un-merged LoRA adapters are outside the accelerated hot path: merge them (args.lora = None)"
        lora_state_dict = {k: v.to(self.device) for k, v in lora_state_dict.items()}
        with torch.no_grad():
            for key, weight in self.state_dict().items():
                if not key.endswith(".weight") or key == "output.weight" or not key.startswith("layers."):
                    continue
                name = key[: -len(".weight")]
                if (name + ".lora_B.weight") in lora_state_dict:
                    merged = weight + (lora_state_dict[name + ".lora_B.weight"] @ lora_state_dict[name + ".lora_A.weight"]) * scaling
                    assert self._assign(key, merged)

    @staticmethod
    def empty(args: TransformerArgs, device: Union[torch.device, str] = "cuda", dtype: torch.dtype = torch.bfloat16, **kwargs: Any) -> "Transformer":
        """A model with UNINITIALISED parameters allocated once, directly in `dtype` on `device` (shapes are laid out on `meta`
        first).  `Transformer(args)` under `torch.device("cuda")` would allocate fp32 and run the Embedding initialiser there:
        2-3x the model's bf16 bytes at the peak (Mixtral-8x7B: 187 GB)."""
        dev = torch.device(device)
        if dev.type == "cuda" and dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        with torch.device("meta"):
            m = Transformer(args, **kwargs)
        return m.to(dtype=dtype).to_empty(device=dev)

    @staticmethod
    def from_folder(folder: Union[Path, str], max_batch_size: int = 1, num_pipeline_ranks: int = 1,
                    device: Union[torch.device, str] = "cuda", dtype: Optional[torch.dtype] = None,
                    softmax_fp32: bool = True, expert_parallel: Optional[Tuple[int, int]] = None, expert_group: Any = None) -> "Transformer":
        """transformer.py:297-338.  Tensors stream from disk straight into the packed device buffers; with `expert_parallel`
        the experts of other ranks are skipped (never read into device memory)."""
        with open(Path(folder) / "params.json", "r") as f:
            model_args = TransformerArgs.from_dict(json.load(f))
        model_args.max_batch_size = max_batch_size
        pipeline_rank = torch.distributed.get_rank() if num_pipeline_ranks > 1 else 0

        pt_model_file = Path(folder) / "consolidated.00.pth"
        safetensors_model_file = Path(folder) / "consolidated.safetensors"
        assert pt_model_file.exists() or safetensors_model_file.exists(), f"Make sure either {pt_model_file} or {safetensors_model_file} exists"
        assert not (pt_model_file.exists() and safetensors_model_file.exists()), f"Both {pt_model_file} and {safetensors_model_file} cannot exist"

        dev = torch.device(device)
        if dev.type == "cuda" and dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())

        def build(ck_dtype: torch.dtype) -> "Transformer":
            # shapes on `meta`, storage allocated ONCE, directly in the target dtype on the target device (the reference builds
            # on meta and assigns, transformer.py:321-331; a fp32 build followed by .to(bf16) would need 3x the model's bytes)
            return Transformer.empty(model_args, dev, dtype or ck_dtype, pipeline_rank=pipeline_rank, num_pipeline_ranks=num_pipeline_ranks,
                                     softmax_fp32=softmax_fp32, expert_parallel=expert_parallel, expert_group=expert_group)

        if pt_model_file.exists():
            loaded = torch.load(str(pt_model_file), mmap=True)
            model = build(next(iter(loaded.values())).dtype)
            model.load_state_dict(loaded, strict=True)
        else:
            import safetensors

            with safetensors.safe_open(str(safetensors_model_file), framework="pt", device="cpu") as f:
                keys = list(f.keys())
                probe = "norm.weight" if "norm.weight" in keys else keys[0]
                model = build(f.get_tensor(probe).dtype)
                loaded_keys = set()
                with torch.no_grad():
                    for k in keys:
                        if model._owns_key(k) and model._assign(k, f.get_tensor(k)):
                            loaded_keys.add(k)
                missing = set(model.reference_keys()) - loaded_keys
                assert not missing, f"missing keys: {sorted(missing)[:8]}"
        return model.eval()
