"""BASELINE config 5 on 1..8 GPUs: MC predictive mean/variance of WHVILinear(32768, 32768) over synthetic inputs,
256 MC samples, sharded over the ranks in two dimensions ("sharding MC samples and minibatch rows"):
    rank = (sample group a, row group b),  sample groups G_s = min(2, N),  row groups G_r = N / G_s
Per input chunk a rank computes t2 = H(s2 x) for its row group in one pass (whvi_fwht_scaled_f32) and runs the fused moments
kernel (whvi_layer_moments_f32: the forward over its S / G_s samples with sum y / sum y^2 kept in tensor memory -- no
prediction ever reaches HBM).  The two ranks of a sample-group pair share their rows and each FINISHES half of them:

  --exchange peer (default): no collective kernel.  A rank first runs the rows its PARTNER finishes, with the kernel's output
      pointers aimed at the partner's staging buffer (symmetric memory, NVLink-mapped): the partial sums travel as the
      kernel's own stores.  After one signal-pad barrier it runs its own rows starting from what the partner stored
      (whvi_layer_moments_add_f32) and the sums land in the result.  No SM is set aside, the sums make no second trip
      through HBM.  Three staging slots keep the partner's next store off the slot being read.
  --exchange nccl: reduce-scatter of the partial sums on a side stream, next to a kernel that leaves a few SMs free for it
      (WHVI_LAYER_RESERVE_SMS; the pair communicator is capped at as many CTAs).  With so few channels the reduce-scatter
      itself becomes the limit at 8 GPUs (6.66x against 7.72x for the peer exchange).

Why not samples over all N ranks: the exchange is 8*D bytes per input per rank whatever N is (256 KB at D = 2^15), so with
256/8 = 32 samples per rank it would be a quarter of the compute time; with two sample groups it is 1/8 of that and pairwise.
The noise is ONE (S, D) draw for all inputs, as in the reference's eval_model (src/networks.py:101-115: one forward pass of S
samples over the whole test batch).  The inputs are resident in HBM before the timed region (a fixed function of the global
row index) and the product is (sum_s y, sum_s y^2) for every input -- what WHVINetwork.predictive_sums returns; the checksum
(mean |predictive mean|, mean predictive variance) is taken after the timed region and must agree for N = 1, 2, 4, 8 up to
fp32 summation order.

    python tools/bench_eval.py [--inputs 170496] [--exchange peer|nccl]
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

RESERVE_SMS = 2   # left to the pairwise reduce-scatter when there is one (N >= 2); its communicator is capped at as many CTAs
_PAIR_GROUPS = {}
_CAP = [RESERVE_SMS]   # CTAs the pair communicator may use = SMs the moments kernel leaves free


def _pair_group(world, g_s, rank):
    """The process group of the ranks that share this rank's rows (same row group, different sample groups)."""
    if g_s == 1:
        return None
    key = (world, g_s)
    if key not in _PAIR_GROUPS:   # every rank creates every group, in the same order
        opts = None
        try:   # the exchange must fit the SMs the moments kernel leaves free: cap this communicator's CTAs
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = RESERVE_SMS
            opts.config.min_ctas = 1
        except Exception:   # older builds: the NCCL_MAX_CTAS environment variable (set by the callers) is the cap
            opts = None
            _CAP[0] = int(os.environ.get("NCCL_MAX_CTAS", "4"))
        _PAIR_GROUPS[key] = [dist.new_group(list(range(b * g_s, (b + 1) * g_s)), pg_options=opts) for b in range(world // g_s)]
    return _PAIR_GROUPS[key][rank // g_s]


def synthetic_rows(row0, rows, D, dev):
    """Inputs as a fixed function of (global row, column): the same data whatever the chunking and the rank count."""
    r = torch.arange(row0, row0 + rows, device=dev, dtype=torch.float32).unsqueeze(1) * 12.9898
    c = torch.arange(D, device=dev, dtype=torch.float32).unsqueeze(0) * 78.233
    return torch.add(r, c).sin_()


def run_eval(dev, rank, world, inputs=0, log2d=15, samples=256, warmup_chunks=2, tiles=0, exchange="peer"):
    """Returns (on every rank) the result dict; timing = CUDA events, barrier on both sides, max over ranks."""
    D, S = 1 << log2d, samples
    g_s = 2 if world >= 2 else 1
    g_r = world // g_s
    a_idx, b_idx = rank % g_s, rank // g_s
    assert world == g_s * g_r and S % g_s == 0
    group = _pair_group(world, g_s, rank)
    # how the two ranks of a sample-group pair combine their partial sums:
    #   "peer": no collective kernel -- a rank first runs the rows its PARTNER will finish, with the fused kernel's output
    #           pointers aimed at the partner's staging buffer (symmetric memory, NVLink-mapped: the partial sums travel as
    #           the kernel's own stores), then its own rows starting from what the partner stored (whvi_layer_moments_add_f32);
    #           one signal-pad barrier per chunk orders the two; no SMs are set aside
    #   "nccl": reduce-scatter on a side stream next to a kernel that leaves a few SMs free for it
    peer = stage = handle = peer_stage = None
    if g_s > 1 and exchange == "peer":
        try:   # symmetric staging buffer: 3 slots x (sum y, sum y^2) x the rows this rank finishes per chunk
            import torch.distributed._symmetric_memory as symm_mem
            shape = (3, 2, (tiles or 4 * g_s) * 148 // g_s, D)
            stage = symm_mem.empty(shape, dtype=torch.float32, device=dev)
            handle = symm_mem.rendezvous(stage, group)
            peer_stage = handle.get_buffer(1 - a_idx, shape, torch.float32)
            handle.barrier()
            peer = symm_mem
        except Exception as exc:   # no symmetric memory on this box (the same on every rank): the NCCL exchange
            if rank == 0:
                print(f"[bench_eval] peer exchange unavailable ({type(exc).__name__}: {exc}); using --exchange nccl", file=sys.stderr)
            peer = None
    reserve = _CAP[0] if (g_s > 1 and peer is None) else 0
    # 4 * G_s tiles per CTA of the persistent grid per chunk: a rank of a sample-group pair runs half the samples per row, so
    # it takes twice the rows per chunk to keep the launches as long as at N = 1
    rows_mine = (tiles or 4 * g_s) * (148 - reserve)
    cb = rows_mine * g_r                                  # inputs per chunk over all row groups
    # default: 170496 inputs = a whole number of chunks for N = 1 (592 rows), 2 (1184), 4 (2368) and 8 (4736 rows per chunk) --
    # 172864 for the nccl exchange, whose kernels leave SMs free: every N evaluates the same input set (strong scaling)
    n_chunks = max(1, (inputs or (170496 if reserve == 0 else 172864)) // cb)
    lo, hi = a_idx * (S // g_s), (a_idx + 1) * (S // g_s)
    torch.manual_seed(0)                                   # replicated parameters
    layer = W.WHVISquarePow2Matrix(D, lambda_=1.0).to(dev)
    with torch.no_grad():
        layer.g_mu.copy_(torch.randn(D, generator=torch.Generator().manual_seed(7)).to(dev))
        eps = torch.randn(S, D, generator=torch.Generator().manual_seed(11))[lo:hi].to(dev)
        g = WF.reparam(layer.g_mu, layer.g_rho, eps)       # (S / G_s, D): this rank's weight samples, for ALL inputs
        s1, s2 = layer.s1.detach(), layer.s2.detach()
    n_fin = rows_mine // g_s                               # rows this rank finishes after the exchange
    total_chunks = warmup_chunks + n_chunks
    # the product: (sum_s y, sum_s y^2) for every input this rank finishes -- what `WHVINetwork.predictive_sums` returns; the
    # fused kernel (N = 1) or the pair's reduce-scatter (N >= 2) writes straight into it, nothing else touches the sums
    result = torch.empty(total_chunks, 2, n_fin, D, device=dev)
    buckets = [torch.empty(2, rows_mine, D, device=dev) for _ in range(2)] if (g_s > 1 and peer is None) else None   # nccl: partial sums
    assert peer is None or tuple(stage.shape) == (3, 2, n_fin, D)
    t2buf = [torch.empty(rows_mine, D, device=dev) for _ in range(2)]
    comm = torch.cuda.Stream(device=dev)
    # the inputs of this rank's row group, resident in HBM before the timed region starts (bench contract: `value` is
    # measured with inputs already on the device); warm-up chunks first
    x_all = torch.empty(total_chunks, rows_mine, D, device=dev)
    for c in range(total_chunks):
        x_all[c].copy_(synthetic_rows(c * cb + b_idx * rows_mine, rows_mine, D, dev))

    def prepare(c):
        """t2 = H(s2 x) of this rank's rows of chunk c, one pass (the sample-group partner computes the same rows
        redundantly: about a percent of the chunk's work, and it saves a collective)."""
        WF.fwht_scaled_(x_all[c], s2, out=t2buf[c % 2])

    def run_nccl(first_chunk, count):
        rs_done = [None, None]   # per bucket: its reduce-scatter has finished (recorded on `comm`)
        main = torch.cuda.current_stream()
        prepare(first_chunk)
        for c in range(first_chunk, first_chunk + count):
            if g_s == 1:
                WF.layer_moments_raw(t2buf[c % 2], g, s1, s2, None, result[c, 0], result[c, 1], from_t2=True)
            else:
                bucket = buckets[c % 2]
                if rs_done[c % 2] is not None:             # chunk c - 2's exchange read this bucket
                    main.wait_event(rs_done[c % 2])
                WF.layer_moments_raw(t2buf[c % 2], g, s1, s2, None, bucket[0], bucket[1], from_t2=True, reserve_sms=reserve)
                comm.wait_stream(main)
                with torch.cuda.stream(comm):              # overlaps the next chunk's kernel (which leaves SMs free for it)
                    dist.reduce_scatter_tensor(result[c, 0], bucket[0], group=group)
                    dist.reduce_scatter_tensor(result[c, 1], bucket[1], group=group)
                    rs_done[c % 2] = torch.cuda.Event()
                    rs_done[c % 2].record(comm)
            if c + 1 < first_chunk + count:
                prepare(c + 1)
        for ev in rs_done:
            if ev is not None:
                main.wait_event(ev)

    def run_peer(first_chunk, count):
        """Per chunk k, on ONE stream:  t2(k+1);  theirs(k+1) -> partner's staging[(k+1) % 3];  barrier;  mine(k) = staging[k % 3]
        + my sums -> result[k].  The barrier (after theirs(k+1) on both ranks) guarantees the partner's theirs(k) has landed
        before mine(k) reads it; three staging slots keep the partner's theirs(k+3) -- issued only after it has passed the
        barrier I reach after mine(k) -- off the slot mine(k) is reading."""
        other = slice((1 - a_idx) * n_fin, (2 - a_idx) * n_fin)   # rows the partner finishes
        own = slice(a_idx * n_fin, (a_idx + 1) * n_fin)

        def theirs(c):
            WF.layer_moments_raw(t2buf[c % 2][other], g, s1, s2, None, peer_stage[c % 3, 0], peer_stage[c % 3, 1], from_t2=True)

        def mine(c):
            WF.layer_moments_raw(t2buf[c % 2][own], g, s1, s2, None, result[c, 0], result[c, 1], from_t2=True,
                                 init=(stage[c % 3, 0], stage[c % 3, 1]))

        last = first_chunk + count - 1
        prepare(first_chunk)
        theirs(first_chunk)
        for c in range(first_chunk, last + 1):
            if c < last:
                prepare(c + 1)      # t2buf[(c + 1) % 2]: its last reader, mine(c - 1), is already enqueued
                theirs(c + 1)
            handle.barrier()
            mine(c)
        handle.barrier()            # nobody leaves while its partner may still be writing into it

    run = run_peer if peer is not None else run_nccl

    with torch.no_grad():
        run(0, warmup_chunks)
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
    checksum = torch.zeros(2, device=dev, dtype=torch.float64)
    # checksum of the product, after the timed region: mean |predictive mean| and mean predictive variance over the timed
    # chunks, sum var = sum(sum y^2) / S - sum((sum y)^2) / S^2
    timed_res = result[warmup_chunks:]
    checksum[0] = torch.linalg.vector_norm(timed_res[:, 0], ord=1, dtype=torch.float64) / S
    checksum[1] = timed_res[:, 1].sum(dtype=torch.float64) / S - torch.linalg.vector_norm(timed_res[:, 0], ord=2, dtype=torch.float64) ** 2 / (S * S)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(checksum)
    pairs = n_chunks * cb * S
    n_el = float(n_chunks * cb * D)
    return {"metric": "MC predictive (sample, input) pairs/s", "value": pairs / (ms.item() * 1e-3), "unit": "rows/s",
            "n_gpus": world, "ms": ms.item(), "higher_is_better": True, "scaling": "strong",
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"config5-large-eval: WHVILinear({D},{D}) MC predictive mean/var, {n_chunks * cb} inputs "
                                   f"(bounded sample of 1M) x {S} MC samples", "parallelism": f"{g_s} sample groups x {g_r} row groups",
                       "chunk_inputs": cb, "samples_per_rank": S // g_s, "rows_per_rank_per_chunk": rows_mine,
                       "kernel": "layer_moments_kernel (forward + sum y, sum y^2 in tensor memory, one launch per chunk per rank)",
                       "inputs_resident": "the rank's inputs are in HBM before the timed region; per chunk: t2 = H(s2 x) in one pass "
                                          "(whvi_fwht_scaled_f32), one fused launch whose sums (N = 1) or whose pair exchange (N >= 2) "
                                          "land in the (inputs, 2, D) result; the checksum is taken after the timed region",
                       "collective": "none" if g_s == 1 else (
                           "none (no collective kernel): inside each sample-group pair the fused kernel stores the partner's rows "
                           "straight into the partner's NVLink-mapped staging buffer and starts its own rows from what the partner "
                           "stored; one signal-pad barrier per chunk" if peer is not None else
                           f"pairwise reduce-scatter of (sum y, sum y^2) inside each sample-group pair, overlapped on a side stream "
                           f"({reserve} SMs left free for it)")},
            "equiv_algorithmic_gbs": 8.0 * D * pairs / ms.item() / 1e6,
            "checksum": {"mean_abs_pred_mean": float(checksum[0]) / n_el, "mean_pred_var": float(checksum[1]) / n_el,
                         "inputs": n_chunks * cb,
                         "note": "fixed seeds, data a function of the global row: agrees across n_gpus when `inputs` does"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--inputs", type=int, default=0)
    ap.add_argument("--log2d", type=int, default=15)
    ap.add_argument("--samples", type=int, default=256)
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--tiles", type=int, default=0, help="tiles (rows) per CTA of the persistent grid per chunk (0: 4 x sample groups)")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # the pair exchange must fit the SMs the moments kernel leaves free (RESERVE_SMS); bench.py sets the same
        os.environ.setdefault("NCCL_MAX_CTAS", str(RESERVE_SMS))
        dist.init_process_group("nccl", device_id=dev)
    res = run_eval(dev, rank, world, inputs=args.inputs, log2d=args.log2d, samples=args.samples, tiles=args.tiles, exchange=args.exchange)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
