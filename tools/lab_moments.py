"""Lab: the fused moments kernel alone (no prepare / finish glue), per samples-per-launch and rows-per-launch.
    python tools/lab_moments.py"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from whvi_b200 import functional as WF  # noqa: E402

dev = torch.device("cuda:0")
D = 1 << 15
torch.manual_seed(0)
s1, s2 = torch.randn(D, device=dev), torch.randn(D, device=dev)
for rows in (592, 2368):
    t2 = torch.randn(rows, D, device=dev)
    out = torch.empty(2, rows, D, device=dev)
    for S in (32, 64, 128, 256):
        g = torch.randn(S, D, device=dev)
        for _ in range(2):
            WF.layer_moments_raw(t2, g, s1, s2, None, out[0], out[1], from_t2=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        a.record()
        for _ in range(reps):
            WF.layer_moments_raw(t2, g, s1, s2, None, out[0], out[1], from_t2=True)
        b.record()
        b.synchronize()
        ms = a.elapsed_time(b) / reps
        print(f"rows={rows} S={S}: {ms:.3f} ms/launch, {rows * S / ms / 1e3:.3e} pairs/s, {ms * 1e3 / (rows / 148) / S:.2f} us per (tile, sample) per CTA")
