"""The other BASELINE.json configurations (C1, C2, C3, C5) on one GPU -- parity-test
sized workloads, reported for completeness next to bench.py's headline (C4).

    python tools/bench_configs.py [--out gpurun_out/configs.json] [--c5-inputs 16384]

C1  README toy regression: WHVIRegression[WHVILinear(3,16,lambda_=2), ReLU, WHVILinear(16,1)], batch 64
C2  batched FWHT fwd+bwd (autograd) sweep, D = 2^6..2^15, 2^20..2^28 elements
C3  UCI-shaped MLP [13->128, 128->128, 128->1], batch 4096, 64 MC samples, full Adam step
C5  WHVILinear(32768,32768) MC predictive evaluation, 256 MC samples, predictive mean/variance
    accumulated on the fly over sample chunks (bounded number of inputs; rows/s is per (s,b) pair)
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import whvi_b200 as W  # noqa: E402
from whvi_b200 import FWHTFunction  # noqa: E402
from whvi_b200 import functional as F  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / iters


def c1():
    torch.manual_seed(0)
    x = torch.randn(200, 3, device=dev)
    y = torch.reshape(x[:, 0] + x[:, 1] ** 2 - 0.3 * x[:, 2] ** 3, (-1, 1))
    model = W.WHVIRegression([W.WHVILinear(3, 16, lambda_=2.0), torch.nn.ReLU(), W.WHVILinear(16, 1)]).to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)

    def step():
        loss = model.loss(x[:64], y[:64], n=150)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)

    ms = timed(step, 5, 50)
    # the same step captured as one CUDA graph (this model is launch-bound)
    from whvi_b200.graphs import GraphedTrainStep
    opt_g = torch.optim.Adam(model.parameters(), lr=torch.tensor(1e-3, device=dev), capturable=True)
    gstep = GraphedTrainStep(model, opt_g, x[:64], y[:64], n=150)
    ms_g = timed(lambda: gstep(x[:64], y[:64]), 5, 200)
    return {"config": "C1 README toy, batch 64, S=1", "ms_per_step": ms, "steps_per_s": 1e3 / ms,
            "cuda_graph_ms_per_step": ms_g, "cuda_graph_steps_per_s": 1e3 / ms_g}


def c2():
    out = []
    for log2n in (20, 24, 28):
        n = 1 << log2n
        for k in (6, 10, 13, 15):
            D = 1 << k
            x = torch.randn(n // D, D, device=dev, requires_grad=True)
            dy = torch.randn(n // D, D, device=dev)

            def fb():
                y = FWHTFunction.apply(x)
                y.backward(dy)
                x.grad = None

            ms = timed(fb, 3, 10)
            out.append({"elements": n, "D": D, "fwd_bwd_ms": ms, "gbs": 16.0 * n / ms / 1e6})
    return {"config": "C2 FWHT fwd+bwd through autograd (16 B/elt)", "sweep": out}


def c3():
    torch.manual_seed(0)
    B, S = 4096, 64
    model = W.WHVIRegression([W.WHVILinear(13, 128, lambda_=3.0), torch.nn.ReLU(), W.WHVILinear(128, 128, lambda_=3.0),
                              torch.nn.ReLU(), W.WHVILinear(128, 1, lambda_=3.0)], train_samples=S).to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    x, y = torch.randn(B, 13, device=dev), torch.randn(B, 1, device=dev)

    def step():
        loss = model.loss(x, y, n=B)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)

    ms = timed(step, 5, 20)
    from whvi_b200.graphs import GraphedTrainStep
    opt_g = torch.optim.Adam(model.parameters(), lr=torch.tensor(1e-3, device=dev), capturable=True)
    gstep = GraphedTrainStep(model, opt_g, x, y, n=B)
    ms_g = timed(lambda: gstep(x, y), 5, 50)
    return {"config": "C3 UCI-shaped MLP 13-128-128-1, B=4096, S=64, full Adam step", "ms_per_step": ms,
            "mc_rows_per_s": S * B / (ms * 1e-3), "cuda_graph_ms_per_step": ms_g,
            "cuda_graph_mc_rows_per_s": S * B / (ms_g * 1e-3)}


def c5(n_inputs: int):
    """MC predictive mean/variance, D = 2^15, S = 256: per input chunk t2 = H(s2 x) once, then per
    sample chunk one FROM_T2 forward (one transform per (s, b) pair) and one moments pass."""
    D, S, chunk_b, chunk_s = 1 << 15, 256, 256, 32
    torch.manual_seed(0)
    layer = W.WHVISquarePow2Matrix(D, lambda_=1.0).to(dev)
    mean_abs = torch.zeros((), device=dev)
    t0 = time.perf_counter()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    with torch.no_grad():
        for b0 in range(0, n_inputs, chunk_b):
            x = torch.randn(chunk_b, D, device=dev)          # inputs generated on the device, never materialised in full
            s_y, s_y2, n = layer.predictive_moments(x, S, chunk_samples=chunk_s)
            mean = s_y / n
            var = s_y2 / n - mean * mean
            mean_abs += mean.abs().mean() + var.mean() * 0.0
    b.record()
    b.synchronize()
    ms = a.elapsed_time(b)
    rows = n_inputs * S
    return {"config": f"C5 WHVILinear(32768,32768) MC predictive mean/var, {n_inputs} inputs x {S} samples (bounded sample of 1M inputs)",
            "ms": ms, "mc_rows_per_s": rows / (ms * 1e-3), "fwd_algorithmic_gbs": 8.0 * D * rows / ms / 1e6,
            "note": "first transform hoisted out of the sample loop (1 transform per (s,b) pair); 8*D B/row is the "
                    "unhoisted layer's algorithmic traffic, kept as the common yardstick",
            "wall_s": time.perf_counter() - t0, "checksum": float(mean_abs)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--c5-inputs", type=int, default=8192)
    ap.add_argument("--only", default="", help="comma-separated subset, e.g. C1,C3")
    args = ap.parse_args()
    runs = {"C1": c1, "C2": c2, "C3": c3, "C5": lambda: c5(args.c5_inputs)}
    only = [k for k in args.only.split(",") if k] or list(runs)
    res = {k: runs[k]() for k in only}
    print(json.dumps(res, indent=1))
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
