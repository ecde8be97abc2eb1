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


class Cosine(torch.nn.Module):
    """cos(x) as a module: the interface of the reference's one custom activation (``src/activations.py:5-13``).  None of the
    reference's experiments uses it (``src/evaluation.py:37`` builds ReLU networks), so unlike ``nn.ReLU`` -- which
    ``WHVINetwork`` folds into the layer kernels -- it stays an ordinary elementwise op; inside a ``WHVINetwork`` it sees the
    MC samples folded into the batch like any foreign module."""

    def forward(self, x):
        return x.cos()


_PEER_SLOTS: dict = {}   # (device, group id, shapes) -> symmetric double buffers (allocating + rendezvous is collective and slow: once)


def _peer_slots(device, group, shapes_dtypes):
    """Two sets of symmetric (NVLink peer-mapped) device buffers for ``DevicePrefetcher(shard_over_ranks=True)``, or None when
    symmetric memory is not available.  Collective over ``group`` on first use for a given set of shapes."""
    import torch.distributed as dist
    key = (torch.device(device).index, id(group), shapes_dtypes)
    if key in _PEER_SLOTS:
        return _PEER_SLOTS[key]
    slots = None
    try:
        import torch.distributed._symmetric_memory as symm_mem
        grp = group if group is not None else dist.group.WORLD
        world = dist.get_world_size(grp)
        slots = []
        for _ in range(2):
            local, peers, handles = [], [], []
            for shape, dtype in shapes_dtypes:
                t = symm_mem.empty(shape, dtype=dtype, device=device)
                h = symm_mem.rendezvous(t, grp)
                local.append(t)
                peers.append([h.get_buffer(q, shape, dtype) for q in range(world)])
                handles.append(h)
            slots.append((tuple(local), peers, handles))
    except Exception:
        slots = None
    _PEER_SLOTS[key] = slots
    return slots


class DevicePrefetcher:
    """Iterate over host batches (tuples of pinned tensors) with the host->device copy of
    batch i+1 running on a side stream while batch i is being computed on.  Two fixed sets of
    device buffers are reused (no allocator traffic); a batch handed out stays valid until the
    next one is requested."""

    def __init__(self, batches, device, group=None, shard_over_ranks=False):
        """``shard_over_ranks`` (MC-sample-sharded jobs, where every rank needs the WHOLE minibatch):
        each rank copies only its 1/world slice of the rows over its own PCIe link and hands it to the other ranks over
        NVLink on the side stream, so the job reads the minibatch from host memory once per step instead of once per rank.
        The hand-over uses no SMs when symmetric memory is available: every rank's buffers are peer-mapped and the slice is
        written into each of them by the copy engines (``copy_`` device -> peer), bracketed by two signal-pad barriers (all
        ranks done with the slot / all slices landed); the compute kernels it overlaps with keep every SM.  Fallback: an NCCL
        all-gather (an SM-resident kernel that competes with them: measured 0.5 ms per step at 8 GPUs)."""
        import torch.distributed as dist
        self.batches = iter(batches)
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if (shard_over_ranks and dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        # high priority: the slice all-gather is an SM-resident NCCL kernel, and the layer kernels it overlaps with fill every
        # SM (the backward is one persistent CTA per SM) -- with equal priority it only starts when a whole compute kernel has
        # drained, i.e. it does not overlap at all; with priority its few CTAs take the first slots that free up
        self.stream = torch.cuda.Stream(device=self.device, priority=-1)
        # the (cached, peer-mapped) buffers may still be in use by work enqueued before this object existed
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        self.bufs = [None, None]
        self.views = [None, None]   # what is handed out: the leading rows of bufs that the batch fills
        self.ready = [None, None]   # copy finished (recorded on the side stream)
        self.done = [None, None]    # consumer finished with the slot (recorded on its stream)
        self.i = 0                  # index of the next batch to hand out
        self.loaded = 0             # number of batches whose copy has been issued
        self._preload()

    def _preload(self):
        try:
            host = next(self.batches)
        except StopIteration:
            return
        k = self.loaded % 2
        # the slot's buffers are sized by the largest batch seen; a shorter batch (a DataLoader's last one
        # with drop_last=False) is copied into -- and handed out as -- the leading rows of the buffers
        fits = self.bufs[k] is not None and len(self.bufs[k]) == len(host) and all(
            d.shape[1:] == h.shape[1:] and d.dtype == h.dtype and d.size(0) >= h.size(0) for d, h in zip(self.bufs[k], host))
        peer = None
        if self.world > 1 and all(h.size(0) % self.world == 0 and h.size(0) > 0 for h in host):
            slots = _peer_slots(self.device, self.group, tuple((tuple(h.shape), h.dtype) for h in host))
            if slots is not None:
                peer = slots[k]
                self.bufs[k] = peer[0]
                fits = True
        if not fits:
            self.bufs[k] = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in host)
        views = tuple(d[:h.size(0)] for d, h in zip(self.bufs[k], host))
        with torch.cuda.stream(self.stream):
            if self.done[k] is not None:
                self.stream.wait_event(self.done[k])
            if peer is not None:
                _, peers, handles = peer
                handles[0].barrier()                      # every rank's consumer has finished with this slot
                for t, (d, h) in enumerate(zip(views, host)):
                    n = h.size(0) // self.world
                    rows = slice(self.rank * n, (self.rank + 1) * n)
                    d[rows].copy_(h[rows], non_blocking=True)            # my slice: host -> my buffer (PCIe)
                    for q in range(self.world):
                        if q != self.rank:
                            peers[t][q][rows].copy_(d[rows], non_blocking=True)   # -> rank q's buffer (NVLink, copy engine)
                handles[0].barrier()                      # all slices have landed everywhere
                host = ()
            for d, h in zip(views, host):
                if self.world > 1 and h.size(0) % self.world == 0 and h.size(0) > 0:
                    import torch.distributed as dist
                    n = h.size(0) // self.world
                    mine = d[self.rank * n:(self.rank + 1) * n]
                    mine.copy_(h[self.rank * n:(self.rank + 1) * n], non_blocking=True)
                    dist.all_gather_into_tensor(d, mine, group=self.group)  # in place, on the side stream
                else:
                    d.copy_(h, non_blocking=True)
            self.views[k] = views
            self.ready[k] = torch.cuda.Event()
            self.ready[k].record(self.stream)
        self.loaded += 1

    def __iter__(self):
        return self

    def __next__(self):
        cur = torch.cuda.current_stream(self.device)
        if self.i > 0:  # the consumer has issued all its work on the previous batch
            p = (self.i - 1) % 2
            self.done[p] = torch.cuda.Event()
            self.done[p].record(cur)
        if self.i >= self.loaded:
            raise StopIteration
        k = self.i % 2
        cur.wait_event(self.ready[k])
        self.i += 1
        out = self.views[k]
        self._preload()
        return out
