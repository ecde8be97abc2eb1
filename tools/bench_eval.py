"""BASELINE config 5 on 1..8 GPUs: MC predictive mean/variance of WHVILinear(32768, 32768) over synthetic inputs,
256 MC samples, the MC samples sharded over the ranks.

Per input chunk and rank: ONE launch of the fused moments kernel (whvi_layer_moments_f32: the forward over this rank's
samples with sum y / sum y^2 kept in tensor memory -- no prediction ever reaches HBM), then the partial sums are
reduce-scattered over NVLink on a side stream so that each rank finishes mean/variance for its 1/N of the chunk's
inputs while the next chunk computes.  The input side is sharded too: each rank prepares t2 = H(s2 x) for 1/N of the
chunk's inputs and the slices are all-gathered one chunk ahead.  The noise is ONE (S, D) draw for all inputs, as in the
reference's eval_model (src/networks.py:101-115: one forward pass of S samples over the whole test batch).

Inputs and noise are generated from fixed seeds in blocks that do not depend on the number of ranks, so the printed
checksum must agree for N = 1, 2, 4, 8 (up to fp32 summation order).

    python tools/bench_eval.py [--inputs 18944]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_eval.py

Also imported by bench.py (`run_eval`), which puts the result under the `eval` key of its JSON line.
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
from whvi_b200 import functional as WF  # noqa: E402
from whvi_b200.fwht import fwht_  # noqa: E402

SUB_BLOCKS = 8   # input rows of a chunk are generated in 8 seeded sub-blocks: the same data for 1, 2, 4 or 8 ranks


def run_eval(dev, rank, world, inputs=0, log2d=15, samples=256, chunk_inputs=592, warmup_chunks=2):
    """Returns (on every rank) the result dict; timing = CUDA events, barrier on both sides, max over ranks."""
    D, S, cb = 1 << log2d, samples, chunk_inputs
    assert S % world == 0 and cb % SUB_BLOCKS == 0 and SUB_BLOCKS % world == 0
    n_chunks = max(1, (inputs or 32 * cb) // cb)
    lo, hi = rank * (S // world), (rank + 1) * (S // world)
    torch.manual_seed(0)                                   # replicated parameters
    layer = W.WHVISquarePow2Matrix(D, lambda_=1.0).to(dev)
    with torch.no_grad():
        layer.g_mu.copy_(torch.randn(D, generator=torch.Generator().manual_seed(7)).to(dev))
        eps = torch.randn(S, D, generator=torch.Generator().manual_seed(11))[lo:hi].to(dev)
        g = WF.reparam(layer.g_mu, layer.g_rho, eps)       # (S/N, D): this rank's weight samples, for ALL inputs
        s1, s2 = layer.s1.detach(), layer.s2.detach()
    sub = cb // SUB_BLOCKS
    n_mine = cb // world
    gen_x = torch.Generator(device=dev)
    buckets = [torch.empty(2, cb, D, device=dev) for _ in range(2)]        # (sum y, sum y^2) partials, double-buffered
    mine = [torch.empty(2, n_mine, D, device=dev) for _ in range(2)]       # this rank's share after the reduce-scatter
    t2_full = [torch.empty(cb, D, device=dev) for _ in range(2)]
    t2_ready = [torch.cuda.Event(), torch.cuda.Event()]
    comm = torch.cuda.Stream(device=dev)
    checksum = torch.zeros(2, device=dev, dtype=torch.float64)

    def prepare(c):
        """t2 of input chunk c: this rank's row slice on the main stream, the all-gather on `comm`."""
        full = t2_full[c % 2]
        sl = full[rank * n_mine:(rank + 1) * n_mine]
        x = torch.empty(n_mine, D, device=dev)
        for j in range(SUB_BLOCKS // world):               # rank-count independent data
            gen_x.manual_seed(1000 + c * SUB_BLOCKS + rank * (SUB_BLOCKS // world) + j)
            x[j * sub:(j + 1) * sub].normal_(generator=gen_x)
        fwht_(x * s2, out=sl)
        if world > 1:
            comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(comm):
                dist.all_gather_into_tensor(full, sl)
                t2_ready[c % 2].record(comm)
        else:
            t2_ready[c % 2].record()

    def finish(part, ev):
        torch.cuda.current_stream().wait_event(ev)
        mean = part[0] / S
        var = part[1] / S - mean * mean
        checksum[0] += mean.abs().sum(dtype=torch.float64)
        checksum[1] += var.sum(dtype=torch.float64)

    def run(first_chunk, count):
        pending = None
        prepare(first_chunk)
        for c in range(first_chunk, first_chunk + count):
            if c + 1 < first_chunk + count:
                prepare(c + 1)                             # one chunk ahead: overlaps this chunk's kernel
            torch.cuda.current_stream().wait_event(t2_ready[c % 2])
            bucket = buckets[c % 2]
            WF.layer_moments_raw(t2_full[c % 2], g, s1, s2, None, bucket[0], bucket[1], from_t2=True)
            ev = torch.cuda.Event()
            part = bucket
            if world > 1:
                part = mine[c % 2]
                comm.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(comm):              # overlaps the next chunk's kernel
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

    with torch.no_grad():
        run(0, warmup_chunks)
        checksum.zero_()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run(warmup_chunks, n_chunks)
        b.record()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([a.elapsed_time(b)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(checksum)
    pairs = n_chunks * cb * S
    n_el = float(n_chunks * cb * D)
    return {"metric": "MC predictive (sample, input) pairs/s", "value": pairs / (ms.item() * 1e-3), "unit": "rows/s",
            "n_gpus": world, "ms": ms.item(), "higher_is_better": True, "scaling": "strong", "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"config5-large-eval: WHVILinear({D},{D}) MC predictive mean/var, {n_chunks * cb} inputs "
                                   f"(bounded sample of 1M) x {S} MC samples", "parallelism": f"mc-sample-shard x{world}",
                       "chunk_inputs": cb, "samples_per_rank": S // world,
                       "kernel": "layer_moments_kernel (forward + sum y, sum y^2 in tensor memory, one launch per chunk per rank)",
                       "collective": "none" if world == 1 else "all-gather of t2 slices (one chunk ahead) + reduce-scatter of "
                                                              "(sum y, sum y^2) per input chunk, both overlapped on a side stream"},
            "equiv_algorithmic_gbs": 8.0 * D * pairs / ms.item() / 1e6,
            "checksum": {"mean_abs_pred_mean": float(checksum[0]) / n_el, "mean_pred_var": float(checksum[1]) / n_el,
                         "note": "fixed seeds, rank-count independent data: must agree across n_gpus"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--inputs", type=int, default=0)
    ap.add_argument("--log2d", type=int, default=15)
    ap.add_argument("--samples", type=int, default=256)
    ap.add_argument("--chunk-inputs", type=int, default=592)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    res = run_eval(dev, rank, world, inputs=args.inputs, log2d=args.log2d, samples=args.samples, chunk_inputs=args.chunk_inputs)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
