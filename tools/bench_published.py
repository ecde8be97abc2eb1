"""The one benchmark the reference publishes: `benchmarks/walsh_plot.py:43-54` -- the batched
FWHT on the GPU, batch 512 x D = 2^6..2^11 fp32, time per call over 1000 calls of `FWHTFunction.apply` (15-17 us on
an RTX 2070S, unsynchronised).  Measured here on the same GPU for
  * this repo's `whvi_b200.FWHTFunction.apply` (autograd Function -> ctypes -> whvi_fwht_f32),
  * the raw `whvi_b200.fwht_` call (no autograd node),
  * the reference's own CUDA extension recompiled for sm_100a (oracle/_ref/fwht_cuda.so) through its own autograd
    wrapper restated here exactly as src/fwht/cuda/fwht.py:5-16 (forward = fwht_cuda.fwht(x)),
both the reference's way (loop of 1000 calls, wall clock, NO synchronisation inside the loop: launch rate) and
synchronised (one `torch.cuda.synchronize()` after the loop: sustained per-call time), plus CUDA-event device time.
    python tools/bench_published.py [--out gpurun_out/published.json]"""
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import whvi_b200 as W  # noqa: E402
from oracle import ref_torch  # noqa: E402  (baseline leg: the recompiled reference extension)

CALLS = 1000


def per_call(fn, x, sync_inside=False):
    for _ in range(50):
        fn(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter_ns()
    for _ in range(CALLS):
        fn(x)
    t_unsync = (time.perf_counter_ns() - t0) / CALLS / 1e3      # the reference's number: launch rate
    torch.cuda.synchronize()
    t_sync = (time.perf_counter_ns() - t0) / CALLS / 1e3        # all work done
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(CALLS):
        fn(x)
    b.record()
    b.synchronize()
    return {"us_per_call_unsynchronised": round(t_unsync, 2), "us_per_call_synchronised": round(t_sync, 2),
            "us_per_call_device_events": round(a.elapsed_time(b) * 1e3 / CALLS, 2)}


def main():
    out_path = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else ""
    dev = torch.device("cuda:0")
    ref = ref_torch.fwht_cuda_module()

    class RefFWHTFunction(torch.autograd.Function):  # src/fwht/cuda/fwht.py:5-16, verbatim behaviour
        @staticmethod
        def forward(ctx, x):
            return ref.fwht(x)

        @staticmethod
        def backward(ctx, grad_output):
            return RefFWHTFunction.apply(grad_output)

    res = {"batch": 512, "calls": CALLS, "published_rtx2070s_us": {64: 15, 128: 17, 256: 17, 512: 16, 1024: 15, 2048: 17}, "sweep": []}
    for k in range(6, 12):
        D = 1 << k
        x = torch.randn(512, D, device=dev)
        rec = {"D": D, "ours_FWHTFunction": per_call(W.FWHTFunction.apply, x), "ours_raw_fwht_": per_call(W.fwht_, x)}
        if ref is not None:
            rec["reference_ext_sm100a_FWHTFunction"] = per_call(RefFWHTFunction.apply, x)
            rec["reference_ext_sm100a_raw"] = per_call(ref.fwht, x)
            rec["max_abs_diff"] = float((W.fwht_(x) - ref.fwht(x)).abs().max())
        res["sweep"].append(rec)
        print(json.dumps(rec))
    if out_path:
        Path(out_path).write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
