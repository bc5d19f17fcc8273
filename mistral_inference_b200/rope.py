"""RoPE table (mirror of mistral_inference/rope.py:6-10).

The table is built on the CPU in fp32 exactly as the reference does (transformer.py:108-120 builds it on
the default device, then moves it) and uploaded, so the kernels multiply by the very same cos/sin bits
instead of calling sincosf on the device.  The rotation itself
(rope.py:13-23) is the epilogue of the fused QKV kernel (csrc/epilogue.cuh, EPI_QKV_ROPE).
"""
import torch


def precompute_freqs_cis(dim: int, end: int, theta: float) -> torch.Tensor:
    """complex64 [end, dim/2]: polar(1, t * theta^(-2i/dim))."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim))
    t = torch.arange(end, device=freqs.device)
    freqs = torch.outer(t, freqs).float()
    return torch.polar(torch.ones_like(freqs), freqs)
