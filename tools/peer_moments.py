"""EXPERIMENT (not part of the product package): reduce-scatter of the evaluation path's (sum y, sum y^2)
over NVLink peer memory without a collective kernel.  Measured in round 1 at 5.7x on 8 GPUs against 6.9x
for NCCL's reduce-scatter (eight small launches + a barrier wait per chunk), so NCCL is what
``tools/bench_eval.py`` uses by default; kept here for the A/B (``--exchange peer``)."""
from __future__ import annotations

import torch
import torch.distributed as dist


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
