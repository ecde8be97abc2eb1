"""ncu target: the fused MC-moments kernel of the config-5 evaluation (D = 2^15, 592 inputs x 32 samples, FROM_T2)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from whvi_b200 import functional as WF  # noqa: E402
dev = torch.device("cuda:0")
D, rows, S = 1 << 15, 592, 32
torch.manual_seed(0)
s1, s2, g = torch.randn(D, device=dev), torch.randn(D, device=dev), torch.randn(S, D, device=dev)
x = torch.randn(rows, D, device=dev)
out = torch.empty(2, rows, D, device=dev)
for _ in range(2):
    t2 = WF.fwht_scaled_(x, s2)
    WF.layer_moments_raw(t2, g, s1, s2, None, out[0], out[1], from_t2=True)
torch.cuda.synchronize()
print("ok")
