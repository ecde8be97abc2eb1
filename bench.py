#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the WHVI hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    (N > 1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...)

Workload (BASELINE.json configs[3], the largest one that fits a single GPU): a wide WHVI
MLP of three WHVILinear(4096, 4096) layers with ReLU between them, minibatch B = 8192,
S = 128 MC samples => 2^20 (sample, row) pairs per layer per step.  One "step" = forward +
ELBO (Gaussian MNLL + KL) + backward of that model over the whole batch, plus the
parameter-gradient all-reduce (N > 1) and an Adam step.  MC samples are sharded across the
N GPUs (strong scaling: S stays 128) and processed in chunks of `--chunk` samples per
kernel launch so that activations stay at 2 GiB per tensor.

metric  = "WHVILinear fwd+bwd MC-sample rows/s": S*B*3 layer-rows / step time (a row = one
          (sample, minibatch-row) pair pushed through one WHVILinear forward AND backward).
value   = inputs resident in HBM;   e2e = same step, but x/y start in pinned host memory and
          the loss is read back to the host every step (through the public Python API).
Also reported: the batched-FWHT GB/s sweep (`fwht`), the roofline of the dominant kernel
(the fused layer backward, 12*D algorithmic bytes per row), and the CPU baseline.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "WHVILinear fwd+bwd MC-sample rows/s"
UNIT = "rows/s"
D_MODEL, B_BATCH, S_TOTAL, N_LAYERS = 4096, 8192, 128, 3
WORKLOAD = "config4-wide-mlp: 3x WHVILinear(4096,4096)+ReLU, B=8192, S=128 MC samples, fwd+bwd"


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        except Exception:
            pass
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


# --------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.rows[0][1]),
                "reasons": reasons, "samples": len(self.rows)}


# --------------------------------------------------------------------------- CPU baselines
def cpu_layer_baseline(threads: int, budget_s: float = 12.0):
    """Reference torch-CPU layer path AS WRITTEN (oracle/ref_torch.py, a restatement of
    src/weights.py:66-93) on a bounded sample of the workload: whole fwd+bwd passes of one
    WHVILinear(4096,4096) on 512 of the 8192 rows, one MC sample each, until the budget."""
    import torch
    from oracle import ref_torch
    torch.set_num_threads(threads)
    rows, done, t_used = 512, 0, 0.0
    ref_torch.layer_fwd_bwd_seconds(D_MODEL, 32, 1)  # builds/caches nothing big at D=4096; warms the allocator
    while t_used < budget_s:
        t_used += ref_torch.layer_fwd_bwd_seconds(D_MODEL, rows, 1)
        done += 1
    return {"value": rows * done / t_used, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{done} MC sample(s) x {rows} of {B_BATCH} rows x 1 of {N_LAYERS} layers, D={D_MODEL}, "
                      f"fwd+bwd, torch CPU restatement of the reference layer as written (oracle/ref_torch.py); "
                      f"{t_used:.1f} s of CPU work"}


def cpu_fwht_baseline(threads: int):
    """The reference's own C++ CPU FWHT (oracle/_ref/fwht_cpp.so, compiled unmodified from
    src/fwht/cpp/fwht.cpp) if it travelled, else the C oracle port; D = 1024, 2^18 elements."""
    import numpy as np
    import torch
    from oracle import ref_torch
    torch.set_num_threads(threads)
    D, rows = 1024, 256
    mod = ref_torch.fwht_cpp_module()
    x = torch.randn(rows, D)
    if mod is not None:
        mod.forward(x[:8])
        t0 = time.perf_counter()
        mod.forward(x)
        dt = time.perf_counter() - t0
        kind = "reference"
    else:
        from oracle import oracle as O
        a = x.numpy()
        O.fwht(a[:8])
        t0 = time.perf_counter()
        O.fwht(a)
        dt = time.perf_counter() - t0
        kind = "port"
    return {"gbs": 8.0 * rows * D / dt / 1e9, "kind": kind, "D": D, "rows": rows, "cores": threads, "seconds": dt}


def gpu_fwht_reference_baseline(fwht_ours, dev, log2n: int = 28):
    """The reference's own CUDA FWHT (src/fwht/cuda, recompiled for sm_100a into oracle/_ref/fwht_cuda.so
    with the torch-API renames of SURVEY F3 only) timed on this GPU next to this repo's kernel, same
    inputs, through its public entry (which clones its input, fwht_cuda.cpp:11).  D <= 2^12 only: its
    launch shape is invalid beyond.  Baseline leg: the one other place bench.py may execute
    oracle/."""
    import torch
    from oracle import ref_torch
    mod = ref_torch.fwht_cuda_module()
    if mod is None:
        return None
    n = 1 << log2n
    x = torch.randn(n, device=dev)
    out = []
    for k in (6, 10, 12):
        Dk = 1 << k
        xv = x.view(n // Dk, Dk)
        try:
            got = mod.fwht(xv)
            ours = fwht_ours(xv)
            diff = float((got - ours).abs().max() / ours.abs().max())
            del got, ours
            for _ in range(2):
                mod.fwht(xv)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                mod.fwht(xv)
            b.record()
            b.synchronize()
            ms = a.elapsed_time(b) / 5
            out.append({"D": Dk, "ms": round(ms, 3), "gbs": round(8.0 * n / ms / 1e6, 1), "max_rel_diff_vs_ours": diff})
        except Exception as e:  # a baseline must never take the bench down
            out.append({"D": Dk, "error": str(e)[:120]})
    return {"kind": "reference CUDA kernel recompiled for sm_100a (oracle/_ref/fwht_cuda.so)", "elements": n, "sweep": out}


# --------------------------------------------------------------------------- reference arm
def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    import torch
    from oracle import ref_torch
    torch.set_num_threads(threads)
    rows = 512
    for _ in range(args.warmup):
        ref_torch.layer_fwd_bwd_seconds(D_MODEL, rows, 1)
    t = 0.0
    for _ in range(args.steps):
        t += ref_torch.layer_fwd_bwd_seconds(D_MODEL, rows, 1)
    value = rows * args.steps / t
    sample = (f"each step = 1 MC sample x {rows} of {B_BATCH} rows x 1 of {N_LAYERS} layers, D={D_MODEL}, fwd+bwd, "
              f"torch CPU restatement of the reference layer path as written (src/weights.py:66-93)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import whvi_b200 as W
    from whvi_b200 import functional as WF
    from whvi_b200 import fwht_
    from whvi_b200.distributed import FlatGradAllReduce, shard_samples

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _, S_local = shard_samples(S_TOTAL, rank, world)
    assert S_TOTAL % world == 0
    chunk = min(args.chunk, S_local)
    assert S_local % chunk == 0
    n_chunks = S_local // chunk
    D, B = D_MODEL, B_BATCH

    torch.manual_seed(0)  # parameters identical on every rank
    model = W.WHVIRegression([W.WHVILinear(D, D), torch.nn.ReLU(), W.WHVILinear(D, D), torch.nn.ReLU(),
                              W.WHVILinear(D, D)], train_samples=chunk).to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    grad_allreduce = FlatGradAllReduce(model.parameters())
    gen = torch.Generator().manual_seed(1)  # the same minibatch on every rank (MC-sample sharding)
    x_host = torch.randn(B, D, generator=gen).pin_memory()
    y_host = torch.randn(B, D, generator=gen).pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)
    torch.manual_seed(100 + rank)  # each rank draws its own eps shard
    scale = 1.0 / (n_chunks * world)

    def step(x, y):
        total = None
        for _ in range(n_chunks):
            loss = model.loss(x, y, n=B) * scale
            loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
        grad_allreduce()  # one flat NCCL all-reduce of all parameter gradients (no-op at N = 1)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return total

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- warm-up, then the device-resident timed region -----------------------------------
    for _ in range(max(args.warmup, 3)):
        step(x_dev, y_dev)
    WF.LAUNCH_COUNTS.clear()
    WF.EVENT_SINK = {"whvi_layer_bwd_fused_f32": [], "whvi_layer_bwd_scaled_f32": [], "whvi_layer_fwd_fused_f32": [],
                     "whvi_layer_loss_f32": []}
    with ClockSampler(local_rank) as clk:
        ms_total = timed(lambda: step(x_dev, y_dev), args.steps)
    sink, WF.EVENT_SINK = WF.EVENT_SINK, None
    launches = sum(WF.LAUNCH_COUNTS.values())
    ms_step = ms_total / args.steps
    rows_per_step = S_TOTAL * B * N_LAYERS
    value = rows_per_step / (ms_step * 1e-3)

    # ---- end to end: pinned host inputs in, loss out, every step ---------------------------
    # Every step's x/y are copied from pinned host memory (a fresh H2D copy per step) and the
    # loss is read back to the host; the copy of step i+1 runs on a side stream while step i
    # computes (whvi_b200.utils.DevicePrefetcher), as a data loader would do it.
    from whvi_b200.utils import DevicePrefetcher

    # N > 1: every rank needs the full minibatch (the ranks shard MC samples, not rows).  Each rank
    # copies 1/N of the rows from pinned host memory over its own PCIe link and the slices are
    # all-gathered over NVLink, all on the prefetch stream while the previous step computes: the
    # whole job reads the minibatch from host memory once per step.
    def e2e_run(steps):
        losses = []
        for x, y in DevicePrefetcher(((x_host, y_host) for _ in range(steps)), dev, shard_over_ranks=True):
            losses.append(float(step(x, y).item()))
        return losses

    e2e_run(2)
    ms_e2e = timed(lambda: e2e_run(args.steps), 1) / args.steps
    e2e_value = rows_per_step / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel (fused backward) from events inside the timed region
    peak, peak_src = measured_peaks()
    bwd_ms = [a.elapsed_time(b) for a, b in sink["whvi_layer_bwd_fused_f32"] + sink["whvi_layer_bwd_scaled_f32"]]
    fwd_ms = [a.elapsed_time(b) for a, b in sink["whvi_layer_fwd_fused_f32"]]
    loss_ms = [a.elapsed_time(b) for a, b in sink["whvi_layer_loss_f32"]]
    rows_per_launch = chunk * B
    bwd_avg = sum(bwd_ms) / len(bwd_ms)
    fwd_avg = sum(fwd_ms) / len(fwd_ms)
    bwd_gbs = 12.0 * D * rows_per_launch / (bwd_avg * 1e-3) / 1e9
    fwd_gbs = 8.0 * D * rows_per_launch / (fwd_avg * 1e-3) / 1e9
    traffic = None  # DRAM bytes per launch of that kernel, from the committed ncu --set full capture
    tp = ROOT / "profiles" / "r01_bwd_traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text())["dram_bytes_per_row"] * rows_per_launch
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "layer_bwd_tma_kernel (+ layer_bwd_reduce_{slabs,fold}_kernel)", "achieved": bwd_gbs, "peak": peak,
                "unit": "GB/s", "frac": bwd_gbs / peak, "frac_of_8TBs_nominal": bwd_gbs / 8000.0, "traffic": traffic,
                "traffic_source": "profiles/r01_bwd_traffic.json (ncu dram__bytes_read+write per row x rows per launch)",
                "algorithmic_bytes_per_launch": 12.0 * D * rows_per_launch,
                "peak_source": peak_src, "algorithmic_bytes_per_row": 12 * D, "rows_per_launch": rows_per_launch,
                "avg_launch_ms": bwd_avg, "launches_timed": len(bwd_ms),
                "share_of_step": sum(bwd_ms) / ms_total,
                "fwd_kernel": {"achieved": fwd_gbs, "frac": fwd_gbs / peak, "avg_launch_ms": fwd_avg,
                               "algorithmic_bytes_per_row": 8 * D, "share_of_step": sum(fwd_ms) / ms_total}}
    if loss_ms:  # fused last layer (forward + MNLL residual + backward in one pass: x in, dx out)
        loss_avg = sum(loss_ms) / len(loss_ms)
        loss_gbs = 8.0 * D * rows_per_launch / (loss_avg * 1e-3) / 1e9
        roofline["loss_kernel"] = {"kernel": "layer_loss_kernel (+ layer_bwd_reduce_{slabs,fold}_kernel)", "achieved": loss_gbs,
                                   "frac": loss_gbs / peak, "avg_launch_ms": loss_avg, "algorithmic_bytes_per_row": 8 * D,
                                   "share_of_step": sum(loss_ms) / ms_total,
                                   "note": "replaces a forward (8 B/elt) + backward (12 B/elt) pair of the last layer"}

    # whole step against the same roofline: SURVEY 8d's 20*D bytes per (sample, row, layer) over all ranks
    step_bytes = 20.0 * D * S_TOTAL * B * N_LAYERS
    step_gbs = step_bytes / (ms_step * 1e-3) / 1e9
    roofline["whole_step"] = {"algorithmic_bytes_per_step": step_bytes, "achieved": step_gbs, "unit": "GB/s",
                              "frac": step_gbs / (peak * world), "n_gpus": world,
                              "note": "fwd 8*D + bwd 12*D bytes per (sample, row, layer); the fused last layer and the "
                                      "shared first-layer input move fewer bytes than this yardstick"}

    out = None
    if rank == 0:
        # ---- FWHT GB/s sweep (second half of BASELINE.json's metric), 2^28 elements = 1 GiB in + out
        n = 1 << 28
        xf, yf = torch.randn(n, device=dev), torch.empty(n, device=dev)
        fw = []
        for k in range(6, 16):
            Dk = 1 << k
            xv, yv = xf.view(n // Dk, Dk), yf.view(n // Dk, Dk)
            for _ in range(3):
                fwht_(xv, out=yv)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                fwht_(xv, out=yv)
            b.record()
            b.synchronize()
            gbs = 8.0 * n * 10 / (a.elapsed_time(b) * 1e-3) / 1e9
            fw.append({"D": Dk, "gbs": round(gbs, 1), "frac_of_measured": round(gbs / peak, 4),
                       "frac_of_8TBs_nominal": round(gbs / 8000.0, 4)})
        del xf, yf
        threads = os.cpu_count() or 1
        cpu = cpu_layer_baseline(threads) if not args.no_cpu_baseline else None
        cpu_f = cpu_fwht_baseline(threads) if not args.no_cpu_baseline else None
        gpu_f = gpu_fwht_reference_baseline(fwht_, dev) if not args.no_cpu_baseline else None
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": WORKLOAD, "D": D, "B": B, "S": S_TOTAL, "layers": N_LAYERS,
                          "rows_counted": "S*B*layers per step", "parallelism": f"mc-sample-shard x{world}",
                          "samples_per_launch": chunk, "l2": f"inputs larger than L2 ({chunk * B * D * 4 / 2**30:.0f} GiB activations per launch)",
                          "step": "fwd + MNLL + KL + bwd + grad all-reduce (N>1) + Adam",
                          "fusion": "ReLU folded into the layer kernels; last layer fwd+MNLL+bwd in one kernel"},
               "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e,
                       "h2d_bytes_per_step": int(x_host.numel() * 4 + y_host.numel() * 4), "d2h_bytes_per_step": 4},
               "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
               "fwht": {"elements": n, "unit": "GB/s", "sweep": fw, "cpu_baseline": cpu_f, "ref_cuda_baseline": gpu_f},
               "clocks": clk.summary()}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chunk", type=int, default=32, help="MC samples per kernel launch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
