"""Batched fast Walsh-Hadamard transform on the GPU: drop-in for the reference's
``src/fwht/cuda/fwht.py`` (``FWHTFunction``) plus the ``FWHT`` module the cpp/python
flavours expose (``src/fwht/cpp/fwht.py:21-30``).

``FWHTFunction.apply(x)``: ``x`` is a 2-D CUDA tensor ``(rows, D)`` with ``D`` a power of
two; returns a NEW tensor ``H_D``-transformed along dim 1 (natural order, unnormalised),
the input is left untouched (contract of ``fwht_cuda.cpp:11``).  Backward is the same
transform applied to the incoming gradient (``cuda/fwht.py:14-16``).
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def fwht_(x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """Raw call: transform the rows of ``x`` into ``out`` (``out`` may be ``x``)."""
    # same three checks, same messages, as fwht_cuda.cpp:6-10
    if x.device.type != "cuda":
        raise RuntimeError("X must be a CUDA tensor")
    if x.dim() != 2:
        raise RuntimeError("X must be two-dimensional")
    n = x.size(-1)
    if n < 1 or (n & (n - 1)) != 0:
        raise RuntimeError("n must be a power of 2")
    if x.dtype not in (torch.float32, torch.float64):
        raise RuntimeError(f"whvi_b200 FWHT supports float32 and float64 (got {x.dtype})")
    x = x.contiguous()
    if out is None:
        out = torch.empty_like(x)
    elif out.shape != x.shape or out.dtype != x.dtype or not out.is_contiguous() or out.device != x.device:
        raise RuntimeError("out must be a contiguous tensor of x's dtype and shape on the same device")
    if x.numel() == 0:
        return out
    from . import functional as _F
    name = "whvi_fwht_f32" if x.dtype == torch.float32 else "whvi_fwht_f64"   # same dispatch as fwht_cuda_kernel.cu:170
    with torch.cuda.device(x.device), _F._Timed(name):
        rc = getattr(_lib.lib(), name)(x.data_ptr(), out.data_ptr(), x.size(0), n, _stream_ptr(x.device))
    _lib.check(rc, name)
    return out


class FWHTFunction(Function):
    """Python frontend for the batched FWHT on the GPU (B200-native kernel)."""

    @staticmethod
    def forward(ctx, x):
        return fwht_(x)

    @staticmethod
    def backward(ctx, grad_output):
        return FWHTFunction.apply(grad_output)


class FWHT(nn.Module):
    """``nn.Module`` wrapper, as ``src/fwht/cpp/fwht.py:21-30``."""

    def forward(self, x):
        return FWHTFunction.apply(x)
