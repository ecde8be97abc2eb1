"""Lab: the per-chunk glue of tools/bench_eval.py (prepare = s2 scaling + FWHT, finish = checksum reductions) timed alone."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from whvi_b200.fwht import fwht_  # noqa: E402

dev = torch.device("cuda:0")
D, rows, S = 1 << 15, 592, 256
x = torch.randn(8, rows, D, device=dev)
s2 = torch.randn(D, device=dev)
scaled, t2 = torch.empty(rows, D, device=dev), torch.empty(rows, D, device=dev)
part = torch.randn(2, rows, D, device=dev)
checksum = torch.zeros(2, device=dev, dtype=torch.float64)


def timed(fn, reps=20):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        fn(i)
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / reps


def f_mul(i): torch.mul(x[i % 8], s2, out=scaled)
def f_fwht(i): fwht_(scaled, out=t2)
def f_n1(i): checksum[0].add_(torch.linalg.vector_norm(part[0], ord=1, dtype=torch.float64) / S)
def f_n2(i): checksum[1].add_(torch.linalg.vector_norm(part[0], ord=2, dtype=torch.float64) ** 2 / (S * S))
def f_sum(i): checksum[1].add_(part[1].sum(dtype=torch.float64) / S)
def f_sum32(i): checksum[1].add_(part[1].sum() / S)
def f_abs32(i): checksum[0].add_(part[0].abs().sum() / S)
def f_sq32(i): checksum[1].add_(torch.linalg.vector_norm(part[0]) ** 2)

for name, fn in (("mul", f_mul), ("fwht", f_fwht), ("norm1_f64", f_n1), ("norm2_f64", f_n2), ("sum_f64", f_sum), ("sum_f32", f_sum32),
                 ("abs_sum_f32", f_abs32), ("norm2_f32", f_sq32)):
    print(f"{name}: {timed(fn) * 1e3:.1f} us")
