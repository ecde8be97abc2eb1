"""Small helpers kept for API compatibility with the reference's ``src/utils.py``."""
from __future__ import annotations

import torch

from . import functional as WF


def matmul_diag_left(D_diagonal, A):
    """diag(D_diagonal) @ A without building the diagonal matrix."""
    return D_diagonal.unsqueeze(-1) * A


def matmul_diag_right(A, D_diagonal):
    """A @ diag(D_diagonal)."""
    return A * D_diagonal


def is_pow_of_2(x):
    return bool(x) and not (x & (x - 1))


def kl_diag_normal(mu1, sd1, mu2, sd2):
    """General 4-argument KL of the reference (``src/utils.py:49-71``); the layers use the
    fused kernel (``functional.kl_gaussian``) for the (mu, softplus(rho)) vs (0, lambda) case."""
    d = mu1.numel()
    diff = mu2 - mu1
    return 0.5 * (torch.log(sd2).sum() - torch.log(sd1).sum() - d + (sd1 / sd2).sum() + diff @ (diff / sd2))


def build_H(D, device):
    """Dense Walsh-Hadamard matrix, computed by transforming the identity with the FWHT kernel."""
    from .fwht import fwht_
    assert is_pow_of_2(D)
    return fwht_(torch.eye(D, device=device))


kl_gaussian = WF.kl_gaussian


class DevicePrefetcher:
    """Iterate over host batches (tuples of pinned tensors) with the host->device copy of
    batch i+1 running on a side stream while batch i is being computed on."""

    def __init__(self, batches, device):
        self.batches = iter(batches)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._next = None
        self._preload()

    def _preload(self):
        try:
            host = next(self.batches)
        except StopIteration:
            self._next = None
            return
        with torch.cuda.stream(self.stream):
            dev = tuple(t.to(self.device, non_blocking=True) for t in host)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._next = (dev, ev)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        dev, ev = self._next
        torch.cuda.current_stream(self.device).wait_event(ev)
        for t in dev:
            t.record_stream(torch.cuda.current_stream(self.device))
        self._preload()
        return dev
