"""CPU, gloo, world_size 2: the host-side multi-GPU logic (sample sharding, the single flat
gradient all-reduce, the loss scaling convention, predictive-moment reduction)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from whvi_b200.distributed import FlatGradAllReduce, rank_loss, reduce_predictive_moments, shard_rows, shard_samples
from whvi_b200.optim import FlatParams


def test_shard_samples_partition():
    for total in (0, 1, 7, 64, 128, 129):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                first, n = shard_samples(total, r, world)
                seen.extend(range(first, first + n))
            assert seen == list(range(total))
    assert shard_samples(128, 3, 8) == (48, 16)
    assert shard_rows(10, 1, 4) == (3, 3)
    with pytest.raises(ValueError):
        shard_samples(4, 4, 4)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)  # replicated parameters
        net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
        x = torch.randn(6, 5)
        S = 8
        first, n_local = shard_samples(S, rank, world)
        gen = torch.Generator().manual_seed(123)
        noise = torch.randn(S, 6, 3, generator=gen)          # the "MC samples"
        kl = sum((p ** 2).sum() for p in net.parameters())   # replicated regulariser, like the KL term
        mnll_local = ((net(x).unsqueeze(0) + noise[first:first + n_local]) ** 2).mean()
        ((mnll_local + kl) / world).backward()
        unused = torch.nn.Parameter(torch.zeros(4))           # a parameter without a gradient on this rank
        reducer = FlatGradAllReduce(list(net.parameters()) + [unused])
        reducer()
        grads = [p.grad.clone() for p in net.parameters()]
        # single-process reference: all samples at once
        net.zero_grad()
        (((net(x).unsqueeze(0) + noise) ** 2).mean() + sum((p ** 2).sum() for p in net.parameters())).backward()
        ok = all(torch.allclose(g, p.grad, atol=1e-6) for g, p in zip(grads, net.parameters()))
        ok = ok and unused.grad is not None and float(unused.grad.abs().sum()) == 0.0
        # predictive moments over samples sharded across ranks
        ys = noise[first:first + n_local]
        mean, var = reduce_predictive_moments(ys.sum(0), (ys ** 2).sum(0), n_local)
        ok = ok and torch.allclose(mean, noise.mean(0), atol=1e-6)
        ok = ok and torch.allclose(var, noise.var(0, unbiased=False), atol=1e-5)
        # uneven shards (S = 7 over 2 ranks) with rank_loss, gradients exchanged through FlatParams (views of one
        # buffer, a single all-reduce, no copies): still the single-process gradient
        S7 = 7
        first7, n7 = shard_samples(S7, rank, world)
        net.zero_grad(set_to_none=True)
        flat = FlatParams(net.parameters())
        kl = sum((p ** 2).sum() for p in net.parameters())
        mnll7 = ((net(x).unsqueeze(0) + noise[first7:first7 + n7]) ** 2).mean()
        rank_loss(mnll7, kl, n7, S7, world).backward()
        ok = ok and flat.attached()
        flat.all_reduce()
        got = [p.grad.clone() for p in net.parameters()]
        flat.zero_grad()
        ok = ok and flat.attached() and float(flat.grad.abs().sum()) == 0.0
        (((net(x).unsqueeze(0) + noise[:S7]) ** 2).mean() + sum((p ** 2).sum() for p in net.parameters())).backward()
        ok = ok and all(torch.allclose(g, p.grad, atol=1e-6) for g, p in zip(got, net.parameters()))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_flat_allreduce_matches_single_process_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(out.get(r) for r in range(world)), dict(out)
