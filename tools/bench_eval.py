"""BASELINE config 5 on 1..8 GPUs: MC predictive mean/variance of WHVILinear(32768, 32768), 256 MC samples,
MC samples sharded over the ranks; per input chunk the partial (sum y, sum y^2) are
reduce-scattered over NVLink (each rank finishes mean/variance for its 1/N of the inputs), and the
input side is sharded too: each rank prepares t2 = H(s2 x) for 1/N of the chunk's inputs and the slices
are all-gathered one chunk ahead.  Both collectives overlap the current chunk's compute.

    python tools/bench_eval.py [--inputs 8192]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        tools/bench_eval.py [--inputs 8192]

Prints one JSON line (rank 0): (sample, input) pairs per second over all ranks, strong scaling
(the sample count stays 256), device time, max over ranks.  `--inputs` is a bounded sample of the
1M inputs of the config; they are generated on the device chunk by chunk (identically on every rank).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import whvi_b200 as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--inputs", type=int, default=8192)
    ap.add_argument("--log2d", type=int, default=15)
    ap.add_argument("--samples", type=int, default=256)
    ap.add_argument("--chunk-inputs", type=int, default=512)
    ap.add_argument("--chunk-samples", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=2, help="untimed input chunks")
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "peer"],
                    help="nccl: local sums + ncclReduceScatter on a side stream (default: measured faster at 8 GPUs); "
                         "peer: the moments kernel stores partial sums straight into the owning rank's memory over "
                         "NVLink (symmetric memory) and a signal-pad barrier publishes them")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    D, S, cb = 1 << args.log2d, args.samples, args.chunk_inputs
    assert S % world == 0
    lo, hi = rank * (S // world), (rank + 1) * (S // world)
    torch.manual_seed(0)                                   # replicated parameters
    layer = W.WHVISquarePow2Matrix(D, lambda_=1.0).to(dev)
    gen_x = torch.Generator(device=dev)
    gen_eps = torch.Generator(device=dev).manual_seed(1000 + rank)   # this rank's samples
    assert cb % world == 0
    buckets = [torch.empty(2, cb, D, device=dev) for _ in range(2)]  # (sum y, sum y^2), double-buffered
    mine = [torch.empty(2, cb // world, D, device=dev) for _ in range(2)]  # this rank's share after the reduce-scatter
    comm = torch.cuda.Stream(device=dev)
    peer = None
    if world > 1 and args.exchange == "peer":
        from tools.peer_moments import PeerMomentExchange
        peer = PeerMomentExchange(cb, D, dev)
    checksum = torch.zeros((), device=dev)

    def finish(part, work_done_event):
        torch.cuda.current_stream().wait_event(work_done_event)
        if peer is not None:
            part = peer.totals(part)                       # `part` is the slot index
        mean = part[0] / S
        var = part[1] / S - mean * mean
        checksum.add_((mean.abs().mean() + 0.0 * var.mean()) / world)

    from whvi_b200.fwht import fwht_
    t2_full = [torch.empty(cb, D, device=dev) for _ in range(2)]
    t2_ready = [torch.cuda.Event(), torch.cuda.Event()]
    n_mine = cb // world

    def prepare(c):
        """t2 of input chunk c: this rank's row slice on the main stream, the all-gather on `comm`."""
        gen_x.manual_seed(c * world + rank)                # this rank's slice of the synthetic inputs
        x = torch.randn(n_mine, D, device=dev, generator=gen_x)
        full = t2_full[c % 2]
        sl = full[rank * n_mine:(rank + 1) * n_mine]
        fwht_(x * layer.s2.detach(), out=sl)
        if world > 1:
            comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(comm):
                dist.all_gather_into_tensor(full, sl)
                t2_ready[c % 2].record(comm)
        else:
            t2_ready[c % 2].record()

    def run(n_chunks, first_chunk):
        pending = None
        prepare(first_chunk)
        for c in range(first_chunk, first_chunk + n_chunks):
            if c + 1 < first_chunk + n_chunks:
                prepare(c + 1)                             # one chunk ahead: overlaps this chunk's kernels
            torch.cuda.current_stream().wait_event(t2_ready[c % 2])
            bucket = buckets[c % 2]
            if peer is not None:
                # the previous chunk's totals must have been read before anyone may overwrite that slot
                # two chunks later: finish it before this chunk's barrier is entered
                if pending is not None:
                    finish(*pending)
                    pending = None
                with torch.no_grad():
                    layer.predictive_moments(None, S, chunk_samples=args.chunk_samples, sample_range=(lo, hi),
                                             out=(bucket[0], bucket[1]), generator=gen_eps, t2=t2_full[c % 2],
                                             scatter_to=peer.destinations(c % 2))
                pending = (c % 2, peer.publish())
                continue
            with torch.no_grad():
                layer.predictive_moments(None, S, chunk_samples=args.chunk_samples, sample_range=(lo, hi),
                                         out=(bucket[0], bucket[1]), generator=gen_eps, t2=t2_full[c % 2])
            ev = torch.cuda.Event()
            part = bucket
            if world > 1:
                part = mine[c % 2]
                comm.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(comm):              # overlaps the next chunk's kernels
                    dist.reduce_scatter_tensor(part[0], bucket[0])
                    dist.reduce_scatter_tensor(part[1], bucket[1])
                    ev.record(comm)
            else:
                ev.record()
            if pending is not None:
                finish(*pending)
            pending = (part, ev)
        if pending is not None:
            finish(*pending)

    run(args.warmup, 0)
    n_chunks = args.inputs // cb
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run(n_chunks, args.warmup)
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if world > 1:
        dist.all_reduce(checksum)
    if rank == 0:
        pairs = n_chunks * cb * S
        print(json.dumps({"metric": "MC predictive (sample, input) pairs/s", "value": pairs / (ms.item() * 1e-3), "unit": "rows/s",
                          "n_gpus": world, "ms": ms.item(), "higher_is_better": True, "scaling": "strong", "dtype": "f32",
                          "data": "synthetic",
                          "config": {"workload": f"WHVILinear({D},{D}) MC predictive mean/var, {n_chunks * cb} inputs "
                                                 f"(bounded sample of 1M) x {S} MC samples", "parallelism": f"mc-sample-shard x{world}",
                                     "chunk_inputs": cb, "chunk_samples": args.chunk_samples,
                                     "collective": "all-gather of t2 slices (one chunk ahead) + reduce-scatter of (sum y, sum y^2) per input chunk, overlapped",
                                     "exchange": args.exchange if world > 1 else "none"},
                          "equiv_algorithmic_gbs": 8.0 * D * pairs / ms.item() / 1e6, "checksum": float(checksum)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
