"""FWHT bandwidth sweep (BASELINE config 2): D = 2^6..2^15 at a fixed element count.
Run on the GPU box:  python tools/bench_fwht.py [--log2n 28] [--out gpurun_out/fwht_sweep.json]"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from whvi_b200 import fwht_  # noqa: E402


def time_op(fn, warmup=5, iters=20):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--kmin", type=int, default=2)
    ap.add_argument("--kmax", type=int, default=15)
    ap.add_argument("--out", default="")
    ap.add_argument("--baselines", action="store_true",
                    help="also time torch's dense fp32 H @ x (what the reference layer uses below D = 2^12, "
                         "src/fwht/python/fwht.py:17-32) on the same inputs; the reference's own CUDA kernel is "
                         "timed by bench.py's baseline leg")
    args = ap.parse_args()
    n = 1 << args.log2n
    dev = torch.device("cuda:0")
    x = torch.randn(n, device=dev)
    y = torch.empty_like(x)
    nbytes = 8.0 * n
    med, best = time_op(lambda: y.copy_(x))
    res = {"elements": n, "copy_gbs_median": nbytes / med / 1e6, "copy_gbs_best": nbytes / best / 1e6, "sweep": []}
    print(f"torch copy {n} floats: median {res['copy_gbs_median']:.0f} GB/s best {res['copy_gbs_best']:.0f} GB/s")
    for k in range(args.kmin, args.kmax + 1):
        D = 1 << k
        xv, yv = x.view(n // D, D), y.view(n // D, D)
        med, best = time_op(lambda: fwht_(xv, out=yv))
        gbs = nbytes / med / 1e6
        row = {"D": D, "rows": n // D, "ms_median": med, "ms_best": best, "gbs_median": gbs,
               "gbs_best": nbytes / best / 1e6}
        if args.baselines and 2 <= k <= 12:
            if k >= 6:
                sgn = torch.tensor([[1.0, 1.0], [1.0, -1.0]], device=dev)
                H = sgn
                while H.size(0) < D:
                    H = torch.kron(H, sgn)
                prev = torch.backends.cuda.matmul.allow_tf32
                torch.backends.cuda.matmul.allow_tf32 = False
                m3, _ = time_op(lambda: torch.matmul(xv, H, out=yv), 2, 5)
                torch.backends.cuda.matmul.allow_tf32 = prev
                row["torch_matmul_fp32_ms"] = m3
                row["torch_matmul_fp32_gbs"] = nbytes / m3 / 1e6
                del H
            print("      baselines:", {a: round(b, 3) for a, b in row.items() if a.startswith("torch_")})
        res["sweep"].append(row)
        print(f"D=2^{k:<2d} rows={n // D:>9d}  {med:8.3f} ms  {gbs:7.0f} GB/s  (best {nbytes / best / 1e6:7.0f})"
              f"  {gbs / res['copy_gbs_median'] * 100:5.1f}% of copy")
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
