"""Fused-layer bandwidth sweep: forward (8 B/elt), backward (12 B/elt) per D.
    python tools/bench_layer.py [--log2n 28] [--out gpurun_out/layer_sweep.json]"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from whvi_b200 import functional as F  # noqa: E402
from tools.bench_fwht import time_op  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--dims", default="16,128,1024,2048,4096,8192")
    ap.add_argument("--samples", type=int, default=16)
    ap.add_argument("--out", default="")
    ap.add_argument("--mode", default="plain", choices=["plain", "shared", "target", "loss", "fromt2"],
                    help="plain: per-sample x; shared: one (B,D) x block for all samples (first layer); "
                         "target: fused MNLL residual (last layer)")
    args = ap.parse_args()
    n = 1 << args.log2n
    dev = torch.device("cuda:0")
    xflat = torch.randn(n, device=dev)
    dyflat = torch.randn(n, device=dev)
    yflat = torch.empty(n, device=dev)
    med, _ = time_op(lambda: yflat.copy_(xflat))
    copy_gbs = 8.0 * n / med / 1e6
    print(f"torch copy: {copy_gbs:.0f} GB/s")
    res = {"elements": n, "copy_gbs": copy_gbs, "sweep": []}
    for D in [int(d) for d in args.dims.split(",")]:
        S = args.samples
        B = n // (S * D)
        x, dy, y = xflat.view(S, B, D), dyflat.view(S, B, D), yflat.view(S, B, D)
        g = torch.randn(S, D, device=dev)
        s1, s2 = torch.randn(D, device=dev), torch.randn(D, device=dev)
        if args.mode == "shared":
            xs = x[0].contiguous()
            f_med, _ = time_op(lambda: F.layer_forward_raw(xs, g, s1, s2, out=y, relu_out=True))
            b_med, _ = time_op(lambda: F.layer_backward_raw(xs, dy, g, s1, s2, want_dx=False), warmup=3, iters=10)
        elif args.mode == "fromt2":  # shared input with the first transform hoisted; "bwd" column = moments pass
            xs = x[0].contiguous()
            f_med, _ = time_op(lambda: F.layer_forward_raw(xs, g, s1, s2, out=y, from_t2=True))
            sy, sy2 = torch.zeros(B, D, device=dev), torch.zeros(B, D, device=dev)
            b_med, _ = time_op(lambda: F.mc_moments_(y, sy, sy2), warmup=3, iters=10)
        elif args.mode == "target":
            tgt = torch.randn(B, D, device=dev)
            coef = torch.tensor(0.5, device=dev)
            f_med, _ = time_op(lambda: F.layer_forward_raw(x, g, s1, s2, out=y, target=tgt))
            b_med, _ = time_op(lambda: F.layer_backward_raw(x, dy, g, s1, s2, want_dx=True, relu_in=True, target=tgt,
                                                            coef=coef), warmup=3, iters=10)
        elif args.mode == "loss":  # fused last layer: the "forward" column is 0, "bwd" is the whole fused pass
            tgt = torch.randn(B, D, device=dev)
            f_med = 1e-9
            b_med, _ = time_op(lambda: F.layer_loss_raw(x, g, s1, s2, None, tgt, want_dx=True, relu_in=True), warmup=3, iters=10)
        else:
            f_med, _ = time_op(lambda: F.layer_forward_raw(x, g, s1, s2, out=y))
            b_med = float("inf")
            if D <= 8192:
                b_med, _ = time_op(lambda: F.layer_backward_raw(x, dy, g, s1, s2, want_dx=True), warmup=3, iters=10)
        rows = S * B
        rec = {"D": D, "S": S, "B": B, "fwd_ms": f_med, "bwd_ms": b_med, "fwd_gbs": 8.0 * n / f_med / 1e6,
               "bwd_gbs": 12.0 * n / b_med / 1e6, "fwdbwd_rows_per_s": rows / ((f_med + b_med) * 1e-3),
               "fwdbwd_gbs": 20.0 * n / (f_med + b_med) / 1e6}
        res["sweep"].append(rec)
        print(f"D={D:5d} S={S} B={B:8d}  fwd {f_med:7.3f} ms {rec['fwd_gbs']:6.0f} GB/s ({rec['fwd_gbs'] / copy_gbs * 100:5.1f}%)"
              f"  bwd {b_med:8.3f} ms {rec['bwd_gbs']:6.0f} GB/s ({rec['bwd_gbs'] / copy_gbs * 100:5.1f}%)"
              f"  fwd+bwd {rec['fwdbwd_rows_per_s']:.3e} rows/s {rec['fwdbwd_gbs']:6.0f} GB/s")
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
