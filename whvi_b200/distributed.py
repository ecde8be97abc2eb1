"""Multi-GPU plumbing for the WHVI hot path: one process per GPU (``torchrun``), MC samples
sharded across ranks, ONE flat all-reduce of the parameter gradients per step.

Every (sample, row) pair is independent given replicated parameters, so the data
path needs no collective; the only exchange is the sum of the small parameter gradients
({d mu, d rho, d s1, d s2, d bias} of every layer plus the likelihood's sigma): 4.D floats per
square block, ~49 K floats (196 KB) for the 3 x 4096 network -- latency-bound, one NCCL call.

Scaling convention: each rank computes ``loss_r = (mnll_r + kl) / world`` on its own sample
shard (``mnll_r`` already divides by its local sample count), so the SUM over ranks of the
gradients equals the single-process gradient of ``mean_s mnll + kl``.
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_samples(total_samples: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(first_sample, n_local) for this rank; samples are split as evenly as possible."""
    if total_samples < 0 or world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad shard arguments")
    base, extra = divmod(total_samples, world_size)
    n_local = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, n_local


def shard_rows(total_rows: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(first_row, n_local): evaluation can shard minibatch rows instead of samples."""
    return shard_samples(total_rows, rank, world_size)


class FlatGradAllReduce:
    """Sum the gradients of ``params`` across ranks through one contiguous fp32 buffer."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        self._flat = None

    def __call__(self, group=None) -> None:
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        if not self.params:
            return
        dev = self.params[0].device
        if self._flat is None or self._flat.device != dev:
            self._flat = torch.empty(self.numel, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self._flat[off:off + n].zero_()
            else:
                self._flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                p.grad = self._flat[off:off + n].reshape(p.shape).clone()
            else:
                p.grad.copy_(self._flat[off:off + n].reshape(p.shape))
            off += n


def reduce_predictive_moments(sum_y: torch.Tensor, sum_y2: torch.Tensor, n_local: int, group=None):
    """Evaluation sharded by MC sample: combine per-rank sum_s y and sum_s y^2 into the
    predictive mean and (biased) variance over all samples."""
    count = torch.tensor([float(n_local)], device=sum_y.device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in (sum_y, sum_y2, count):
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    mean = sum_y / count
    var = sum_y2 / count - mean * mean
    return mean, var


class PeerMomentExchange:
    """Reduce-scatter of the evaluation path's (sum y, sum y^2) over NVLink peer memory, without a
    collective kernel: every rank's reduction kernel (``functional.mc_moments_into``) stores the
    partial sums of the rows owned by rank q straight into q's staging buffer (symmetric memory,
    NVLink-mapped), a signal-pad barrier on a side stream publishes them, and q adds up the
    ``world`` slots for its rows.  No SMs are taken from the transforms by a communication kernel
    and the partial sums never make a second trip through local HBM.

    ``rows`` per input chunk must divide by the world size; ``slots`` chunks may be in flight.
    """

    def __init__(self, rows: int, D: int, device, group=None, slots: int = 2):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if rows % self.world:
            raise RuntimeError("rows per chunk must be a multiple of the world size")
        self.rows, self.D, self.n_mine, self.slots = rows, D, rows // self.world, slots
        self.shape = (slots, self.world, 2, self.n_mine, D)   # [slot][source rank][sum y | sum y^2][row][col]
        self.local = symm_mem.empty(self.shape, dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(self.local, self.group)
        self.peers = [self.handle.get_buffer(q, self.shape, torch.float32) for q in range(self.world)]
        self.stream = torch.cuda.Stream(device=device)
        self.handle.barrier()

    def destinations(self, slot: int):
        """``scatter_to`` argument of ``predictive_moments``: rows of owner q -> q's slot for this rank."""
        return [(q * self.n_mine, (q + 1) * self.n_mine, self.peers[q][slot, self.rank, 0], self.peers[q][slot, self.rank, 1])
                for q in range(self.world)]

    def publish(self) -> torch.cuda.Event:
        """Call after the scattering kernels have been enqueued on the current stream: a barrier on
        the side stream (so the current stream can go on with the next chunk); returns the event to
        wait for before reading ``totals``."""
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            self.handle.barrier()
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return ev

    def totals(self, slot: int):
        """(sum y, sum y^2) over all ranks' samples for the rows this rank owns: (n_mine, D) each."""
        t = self.local[slot].sum(dim=0)
        return t[0], t[1]
