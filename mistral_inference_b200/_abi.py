"""ctypes binding of libmb200.so (C ABI: include/mistral_b200.h).

This is the only place Python touches the native library.  There is NO fallback: if the library is
missing or a call fails, an exception is raised (the product path never routes through PyTorch
reference math or the CPU oracle).
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p
from pathlib import Path
from typing import Optional

import torch

_LIB_PATH = Path(os.environ.get("MB200_LIB_PATH") or Path(__file__).resolve().parent / "libmb200.so")  # override: A/B builds of experiments
_lib: Optional[ctypes.CDLL] = None

ABI_VERSION = 1
SKINNY_MAX_T = 4
WORKSPACE_HEADER_BYTES = 64 * 1024

# name -> (restype, argtypes); mirrors include/mistral_b200.h declaration by declaration
_SIGNATURES = {
    "mb200_abi_version": (c_int, []),
    "mb200_last_error": (c_char_p, []),
    "mb200_device_info": (c_int, [ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "mb200_rmsnorm": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_float, c_void_p]),
    "mb200_attn_qkv": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_float, c_void_p, c_size_t, c_void_p]),
    "mb200_kv_ring_write": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    "mb200_attn_decode": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64,
                                  c_int64, c_void_p, c_size_t, c_void_p]),
    "mb200_attn_prefill": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                   c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p]),
    "mb200_linear_residual": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_size_t, c_void_p]),
    "mb200_ffn_gateup": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_float, c_void_p, c_size_t,
                                 c_void_p]),
    "mb200_lm_head": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_float, c_void_p, c_size_t,
                              c_void_p]),
    "mb200_decode_meta": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "mb200_argmax_rows": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "mb200_logprob_gather": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "mb200_sample_top_p": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_float, c_float, c_void_p]),
    "mb200_moe_sizes": (c_int, [c_int64, c_int64, c_int64, ctypes.POINTER(c_int64), ctypes.POINTER(c_int64), ctypes.POINTER(c_int64)]),
    "mb200_moe_route": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p]),
    "mb200_moe_grouped_ffn": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                      c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mb200_comm_alloc": (c_int, [c_size_t, ctypes.POINTER(c_void_p)]),
    "mb200_comm_free": (c_int, [c_void_p]),
    "mb200_comm_export": (c_int, [c_void_p, c_void_p]),
    "mb200_comm_open": (c_int, [c_void_p, ctypes.POINTER(c_void_p)]),
    "mb200_comm_close": (c_int, [c_void_p]),
    "mb200_workspace_bytes": (c_size_t, [c_int64] * 8),
    "mb200_decode_step": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                  c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_float, c_int64, c_int64, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mb200_debug_set_decode_timeline": (c_int, [c_void_p]),
    "mb200_debug_set_barrier_timeline": (c_int, [c_void_p]),
    "mb200_test_gemm_naive": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
}


class Mb200Error(RuntimeError):
    pass


def library_path() -> Path:
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    """Loads libmb200.so once.  Raises (never falls back) when it is missing or has the wrong ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise Mb200Error(
            f"{_LIB_PATH} not found: build it with `python -m mistral_inference_b200.build` "
            "(there is no PyTorch/CPU fallback for the hot path)")
    handle = ctypes.CDLL(str(_LIB_PATH), mode=os.RTLD_LOCAL if hasattr(os, "RTLD_LOCAL") else 0)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError if a declared symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    if handle.mb200_abi_version() != ABI_VERSION:
        raise Mb200Error(f"libmb200 ABI {handle.mb200_abi_version()} != expected {ABI_VERSION}")
    _lib = handle
    return handle


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().mb200_last_error()
        raise Mb200Error(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "libmb200 takes contiguous CUDA tensors"
    return t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class Workspace:
    """Caller-owned scratch handed to every entry point (zero-filled once: the first 64 KiB hold self-resetting
    counters, see MB200_WORKSPACE_HEADER_BYTES).  Also owns the model-wide mixture-of-experts row buffers and, for an
    expert-parallel model, the NVLink peer memory of the group (mistral_inference_b200/moe.py)."""

    def __init__(self, nbytes: int, device: torch.device):
        self.buf = torch.zeros(max(int(nbytes), WORKSPACE_HEADER_BYTES + 256), dtype=torch.uint8, device=device)
        self._moe: dict = {}
        self._comm = None

    def moe_buffers(self, T: int, dim: int, hidden: int, E: int, k: int, dtype: torch.dtype, comm=None, parity: int = 0):
        from .moe import MoeBuffers

        key = (T, dim, hidden, E, k, parity if comm is not None else 0, id(comm))
        b = self._moe.get(key)
        if b is None:
            if len(self._moe) >= 8:  # prompt chunks of many different lengths: keep the pool small
                self._moe.clear()
            yw_ptr = comm.yw_ptr(parity) if comm is not None else None
            b = self._moe[key] = MoeBuffers(T, dim, hidden, E, k, self.buf.device, dtype, yw_ptr=yw_ptr)
        return b

    def expert_comm(self, layer, T: int, dim: int):
        """The group's peer memory, (re)created collectively when a call needs more rows than it holds."""
        from .moe import ExpertComm

        _, rows_cap, _ = moe_sizes(T, layer.args.num_experts, layer.args.num_experts_per_tok)
        if self._comm is None or self._comm.rows_cap < rows_cap or self._comm.dim != dim:
            if self._comm is not None:
                torch.cuda.synchronize()
                self._comm.close()
                self._moe.clear()
            g, G = layer.expert_shard
            self._comm = ExpertComm(g, G, layer.expert_group, rows_cap, dim, self.buf.device)
        return self._comm

    @property
    def ptr(self) -> int:
        return self.buf.data_ptr()

    @property
    def nbytes(self) -> int:
        return self.buf.numel()


def workspace_bytes(T: int, dim: int, n_heads: int, n_kv_heads: int, head_dim: int, hidden: int, vocab: int, max_batch: int) -> int:
    return int(lib().mb200_workspace_bytes(T, dim, n_heads, n_kv_heads, head_dim, hidden, vocab, max_batch))


def device_info():
    sm, smem = c_int(0), c_int(0)
    _check(lib().mb200_device_info(ctypes.byref(sm), ctypes.byref(smem)), "mb200_device_info")
    return sm.value, smem.value


# ----------------------------------------------------------------------------- thin typed wrappers
def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    T, dim = x.shape
    out = torch.empty_like(x) if out is None else out
    _check(lib().mb200_rmsnorm(_ptr(x), _ptr(w), _ptr(out), T, dim, eps, _stream()), "mb200_rmsnorm")
    return out


def attn_qkv(x, norm_w, wqkv, rope, positions, q_out, k_out, v_out, cache_k, cache_v, cache_rows, n_heads, n_kv_heads, head_dim, eps,
             ws: Workspace) -> None:
    T, dim = x.shape
    _check(lib().mb200_attn_qkv(_ptr(x), _ptr(norm_w), _ptr(wqkv), _ptr(rope), _ptr(positions), _ptr(q_out), _ptr(k_out), _ptr(v_out),
                                _ptr(cache_k), _ptr(cache_v), _ptr(cache_rows), T, dim, n_heads, n_kv_heads, head_dim, eps, ws.ptr,
                                ws.nbytes, _stream()), "mb200_attn_qkv")


def kv_ring_write(k_new, v_new, cache_k, cache_v, cache_rows, n_kv_heads, head_dim) -> None:
    _check(lib().mb200_kv_ring_write(_ptr(k_new), _ptr(v_new), _ptr(cache_k), _ptr(cache_v), _ptr(cache_rows), k_new.shape[0],
                                     n_kv_heads, head_dim, _stream()), "mb200_kv_ring_write")


def attn_decode(q, cache_k, cache_v, kv_len, out, n_heads, n_kv_heads, head_dim, n_splits, ws: Workspace) -> None:
    B = q.shape[0]
    W = cache_k.shape[1]
    _check(lib().mb200_attn_decode(_ptr(q), _ptr(cache_k), _ptr(cache_v), _ptr(kv_len), _ptr(out), B, W, n_heads, n_kv_heads, head_dim,
                                   n_splits, ws.ptr, ws.nbytes, _stream()), "mb200_attn_decode")


def attn_prefill(q, k_new, v_new, cache_k, cache_v, q_start, seqpos, out, B, max_seqlen, W, n_heads, n_kv_heads, head_dim,
                 causal: bool, first_prefill: bool = False) -> None:
    _check(lib().mb200_attn_prefill(_ptr(q), _ptr(k_new), _ptr(v_new), _ptr(cache_k), _ptr(cache_v), _ptr(q_start), _ptr(seqpos),
                                    _ptr(out), q.shape[0], B, max_seqlen, W, n_heads, n_kv_heads, head_dim, (2 if first_prefill else 1) if causal else 0,
                                    _stream()), "mb200_attn_prefill")


def linear_residual(x, w, residual, out, ws: Workspace) -> None:
    T, K = x.shape
    N = w.shape[0]
    _check(lib().mb200_linear_residual(_ptr(x), _ptr(w), _ptr(residual), _ptr(out), T, N, K, ws.ptr, ws.nbytes, _stream()),
           "mb200_linear_residual")


def ffn_gateup(x, norm_w, w13, g_out, eps, ws: Workspace) -> None:
    T, dim = x.shape
    hidden = w13.shape[0] // 2
    _check(lib().mb200_ffn_gateup(_ptr(x), _ptr(norm_w), _ptr(w13), _ptr(g_out), T, dim, hidden, eps, ws.ptr, ws.nbytes, _stream()),
           "mb200_ffn_gateup")


def lm_head(x, norm_w, w_out, logits, eps, ws: Workspace) -> None:
    T, dim = x.shape
    _check(lib().mb200_lm_head(_ptr(x), _ptr(norm_w), _ptr(w_out), _ptr(logits), T, dim, w_out.shape[0], eps, ws.ptr, ws.nbytes,
                               _stream()), "mb200_lm_head")


def argmax_rows(logits: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    T, V = logits.shape
    assert logits.dtype == torch.float32
    out = torch.empty(T, dtype=torch.long, device=logits.device) if out is None else out
    _check(lib().mb200_argmax_rows(_ptr(logits), _ptr(out), T, V, _stream()), "mb200_argmax_rows")
    return out


def logprob_gather(logits: torch.Tensor, target: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """log_softmax(logits, -1)[t, target[t]] (rows with target < 0 are left untouched)."""
    T, V = logits.shape
    assert logits.dtype == torch.float32 and target.dtype == torch.long and target.shape == (T,)
    out = torch.zeros(T, dtype=torch.float32, device=logits.device) if out is None else out
    _check(lib().mb200_logprob_gather(_ptr(logits), _ptr(target), _ptr(out), T, V, _stream()), "mb200_logprob_gather")
    return out


def sample_top_p(logits: torch.Tensor, uniform: torch.Tensor, temperature: float, top_p: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    T, V = logits.shape
    assert logits.dtype == torch.float32 and uniform.dtype == torch.float32 and uniform.shape == (T,)
    out = torch.empty(T, dtype=torch.long, device=logits.device) if out is None else out
    _check(lib().mb200_sample_top_p(_ptr(logits), _ptr(uniform), _ptr(out), T, V, temperature, top_p, _stream()), "mb200_sample_top_p")
    return out


class MoeCommStruct(ctypes.Structure):
    """mb200_moe_comm (include/mistral_b200.h)."""
    _fields_ = [("n_ranks", ctypes.c_int32), ("my_rank", ctypes.c_int32), ("peer_yw", c_void_p * 8), ("my_flags", c_void_p),
                ("peer_flags", c_void_p * 8), ("epoch", c_void_p), ("done_counter", c_void_p)]


def moe_sizes(T: int, E: int, k: int):
    tr, rc, pw = c_int64(0), c_int64(0), c_int64(0)
    _check(lib().mb200_moe_sizes(T, E, k, ctypes.byref(tr), ctypes.byref(rc), ctypes.byref(pw)), "mb200_moe_sizes")
    return tr.value, rc.value, pw.value


def moe_route(hn: torch.Tensor, gate_w: torch.Tensor, E: int, k: int, shard_rank: int, shard_world: int, b) -> None:
    """Router + row plan + gather into the buffers `b` (moe.MoeBuffers)."""
    T, dim = hn.shape
    _check(lib().mb200_moe_route(_ptr(hn), _ptr(gate_w), T, dim, E, k, shard_rank, shard_world, _ptr(b.sel), _ptr(b.wts), _ptr(b.slot), _ptr(b.plan),
                                 _ptr(b.xs), _ptr(b.row_w), _stream()), "mb200_moe_route")


def moe_grouped_ffn(b, w13_host, w2_host, residual: Optional[torch.Tensor], out: torch.Tensor, T: int, dim: int, hidden: int, E: int, k: int,
                    comm: Optional[MoeCommStruct], ws: "Workspace") -> None:
    _check(lib().mb200_moe_grouped_ffn(_ptr(b.xs), ctypes.cast(w13_host, c_void_p), ctypes.cast(w2_host, c_void_p), _ptr(b.plan), _ptr(b.row_w),
                                       _ptr(b.slot), _ptr(residual), _ptr(b.g), b.yw_ptr, _ptr(out), T, dim, hidden, E, k,
                                       ctypes.cast(ctypes.pointer(comm), c_void_p) if comm is not None else None, ws.ptr, ws.nbytes, _stream()),
           "mb200_moe_grouped_ffn")


def comm_alloc(nbytes: int) -> int:
    p = c_void_p(0)
    _check(lib().mb200_comm_alloc(nbytes, ctypes.byref(p)), "mb200_comm_alloc")
    return int(p.value)


def comm_free(ptr: int) -> None:
    _check(lib().mb200_comm_free(ptr), "mb200_comm_free")


def comm_export(ptr: int) -> bytes:
    h = ctypes.create_string_buffer(64)
    _check(lib().mb200_comm_export(ptr, ctypes.cast(h, c_void_p)), "mb200_comm_export")
    return h.raw


def comm_open(handle: bytes) -> int:
    h = ctypes.create_string_buffer(handle, 64)
    p = c_void_p(0)
    _check(lib().mb200_comm_open(ctypes.cast(h, c_void_p), ctypes.byref(p)), "mb200_comm_open")
    return int(p.value)


def comm_close(ptr: int) -> None:
    _check(lib().mb200_comm_close(ptr), "mb200_comm_close")


def decode_meta(seqpos_dev: torch.Tensor, meta_dev: torch.Tensor, windows) -> None:
    """Device-side metadata of a one-token step for every sequence; advances `seqpos_dev` (include/mistral_b200.h)."""
    B = seqpos_dev.shape[0]
    arr = (ctypes.c_int32 * len(windows))(*[int(w) for w in windows])
    assert seqpos_dev.dtype == torch.int32 and meta_dev.dtype == torch.int32 and meta_dev.numel() >= 3 * B + 1 + 2 * B * len(windows)
    _check(lib().mb200_decode_meta(_ptr(seqpos_dev), _ptr(meta_dev), B, ctypes.cast(arr, c_void_p), len(windows), _stream()), "mb200_decode_meta")


def decode_step(layers_dev, windows_dev, n_layers, emb, final_norm, w_out, rope, token_dev, pos, batch_row, logits, next_token, dim, hidden,
                n_heads, n_kv_heads, head_dim, vocab, eps, ws: Workspace, n_experts: int = 0, top_k: int = 0, moe_gate=None, moe_w13=None,
                moe_w2=None) -> None:
    _check(lib().mb200_decode_step(_ptr(layers_dev), _ptr(windows_dev), n_layers, _ptr(emb), _ptr(final_norm), _ptr(w_out), _ptr(rope),
                                   _ptr(token_dev), pos, batch_row, _ptr(logits), _ptr(next_token), dim, hidden, n_heads, n_kv_heads, head_dim,
                                   vocab, eps, n_experts, top_k, _ptr(moe_gate), _ptr(moe_w13), _ptr(moe_w2), ws.ptr, ws.nbytes, _stream()),
           "mb200_decode_step")


def set_decode_timeline(buf: Optional[torch.Tensor]) -> None:
    _check(lib().mb200_debug_set_decode_timeline(_ptr(buf)), "mb200_debug_set_decode_timeline")


def set_barrier_timeline(buf: Optional[torch.Tensor]) -> None:
    _check(lib().mb200_debug_set_barrier_timeline(_ptr(buf)), "mb200_debug_set_barrier_timeline")


def test_gemm_naive(a, w) -> torch.Tensor:
    T, K = a.shape
    c = torch.empty(T, w.shape[0], dtype=torch.float32, device=a.device)
    _check(lib().mb200_test_gemm_naive(_ptr(a), _ptr(w), _ptr(c), T, w.shape[0], K, _stream()), "mb200_test_gemm_naive")
    return c
