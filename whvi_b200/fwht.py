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
from . import functional as _F


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)   # the current stream's handle without a Stream object
_cur_device = getattr(torch._C, "_cuda_getDevice", None) or torch.cuda.current_device


def _stream_ptr(device: torch.device) -> int:
    if _raw_stream is not None:
        return _raw_stream(device.index if device.index is not None else _cur_device())
    return torch.cuda.current_stream(device).cuda_stream


_FN = {}   # dtype -> (callable(in_ptr, out_ptr, rows, D, stream) -> status, name): looked up once, not per call


def _entry(dtype):
    fn = _FN.get(dtype)
    if fn is None:
        # same dispatch as fwht_cuda_kernel.cu:170, plus bf16 activations (fp32 butterflies, one rounding at the store).
        # _lib.lib() hands out the entry point behind the CPython shim (csrc_host/fastcall.c) when that is built: at the sizes
        # of the reference's published benchmark a call is launch-latency bound and ctypes' marshalling was 3.5 of its 9.5 us
        name = {torch.float32: "whvi_fwht_f32", torch.float64: "whvi_fwht_f64", torch.bfloat16: "whvi_fwht_bf16"}[dtype]
        fn = _FN[dtype] = (getattr(_lib.lib(), name), name)
    return fn


def fwht_(x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """Raw call: transform the rows of ``x`` into ``out`` (``out`` may be ``x``).

    The host side is kept to the reference's own per-call work (three checks, one output allocation, one native call;
    ``fwht_cuda.cpp:5-14``): at the sizes the reference's benchmark uses (batch 512, D = 2^6..2^11,
    ``benchmarks/walsh_plot.py:43-54``) a call is launch-latency bound, so every Python-level context manager or lookup
    on this path shows up 1:1 in the per-call time."""
    # same three checks, same messages, as fwht_cuda.cpp:6-10
    if not x.is_cuda:
        raise RuntimeError("X must be a CUDA tensor")
    if x.dim() != 2:
        raise RuntimeError("X must be two-dimensional")
    n = x.size(-1)
    if n < 1 or (n & (n - 1)) != 0:
        raise RuntimeError("n must be a power of 2")
    if x.dtype not in (torch.float32, torch.float64, torch.bfloat16):
        raise RuntimeError(f"whvi_b200 FWHT supports float32, float64 and bfloat16 (got {x.dtype})")
    if not x.is_contiguous():
        x = x.contiguous()
    if out is None:
        out = torch.empty_like(x)
    elif out.shape != x.shape or out.dtype != x.dtype or not out.is_contiguous() or out.device != x.device:
        raise RuntimeError("out must be a contiguous tensor of x's dtype and shape on the same device")
    if x.numel() == 0:
        return out
    fn, name = _entry(x.dtype)
    dev = x.device
    if _F.EVENT_SINK is None and _cur_device() == dev.index:   # the common case: no bookkeeping objects, no device switch
        _F.LAUNCH_COUNTS[name] = _F.LAUNCH_COUNTS.get(name, 0) + 1
        rc = fn(x.data_ptr(), out.data_ptr(), x.size(0), n, _stream_ptr(dev))
    else:
        with torch.cuda.device(dev), _F._Timed(name):
            rc = fn(x.data_ptr(), out.data_ptr(), x.size(0), n, _stream_ptr(dev))
    if rc:
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
