"""Run the fused layer forward and backward a few times for one shape (ncu target).
    python tools/run_layer_once.py [D] [S] [B] [iters]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from whvi_b200 import functional as F  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
S = int(sys.argv[2]) if len(sys.argv) > 2 else 16
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda:0")
x = torch.randn(S, B, D, device=dev)
dy = torch.randn(S, B, D, device=dev)
g = torch.randn(S, D, device=dev)
s1, s2 = torch.randn(D, device=dev), torch.randn(D, device=dev)
y = torch.empty_like(x)
for _ in range(iters):
    F.layer_forward_raw(x, g, s1, s2, out=y)
    F.layer_backward_raw(x, dy, g, s1, s2, want_dx=True)
torch.cuda.synchronize()
print("ok")
