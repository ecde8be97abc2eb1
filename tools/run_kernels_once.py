"""One launch of every hot kernel at representative sizes (ncu --set full target; SURVEY 8d:
FWHT at D = 2^10, 2^13, 2^15; fused forward / backward / loss layer at D = 4096; the MC-evaluation
pair at D = 2^15).  2^27 elements per operand (512 MB > L2).
    python tools/run_kernels_once.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from whvi_b200 import functional as F  # noqa: E402
from whvi_b200 import fwht_  # noqa: E402

dev = torch.device("cuda:0")
n = 1 << 27
x, y, dy = (torch.randn(n, device=dev) for _ in range(3))
for k in (10, 13, 15):
    D = 1 << k
    fwht_(x.view(n // D, D), out=y.view(n // D, D))
D, S = 4096, 16
B = n // (S * D)
g, s1, s2 = torch.randn(S, D, device=dev), torch.randn(D, device=dev), torch.randn(D, device=dev)
tgt = torch.randn(B, D, device=dev)
F.layer_forward_raw(x.view(S, B, D), g, s1, s2, out=y.view(S, B, D), relu_out=True)
F.layer_backward_raw(x.view(S, B, D), dy.view(S, B, D), g, s1, s2, want_dx=True, relu_in=True)
F.layer_loss_raw(x.view(S, B, D), g, s1, s2, None, tgt, want_dx=True, relu_in=True)
D = 1 << 15
B = n // (S * D)
g, s1, s2 = torch.randn(S, D, device=dev), torch.randn(D, device=dev), torch.randn(D, device=dev)
F.layer_forward_raw(x.view(S, B, D)[0].contiguous(), g, s1, s2, out=y.view(S, B, D), from_t2=True)
sy, sy2 = torch.empty(B, D, device=dev), torch.empty(B, D, device=dev)
F.mc_moments_(y.view(S, B, D), sy, sy2, accumulate=False)
torch.cuda.synchronize()
print("ok")
