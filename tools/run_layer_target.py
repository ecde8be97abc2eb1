"""ncu target: fused last-layer variants (MNLL target) of the layer kernels."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from whvi_b200 import functional as F  # noqa: E402
D, S, B = 4096, 16, 4096
dev = torch.device("cuda:0")
x = torch.randn(S, B, D, device=dev); dy = torch.randn(S, B, D, device=dev); g = torch.randn(S, D, device=dev)
s1, s2 = torch.randn(D, device=dev), torch.randn(D, device=dev)
tgt = torch.randn(B, D, device=dev); coef = torch.tensor(0.5, device=dev); y = torch.empty_like(x)
for _ in range(2):
    F.layer_forward_raw(x, g, s1, s2, out=y, target=tgt)
    F.layer_backward_raw(x, dy, g, s1, s2, want_dx=True, relu_in=True, target=tgt, coef=coef)
torch.cuda.synchronize()
print("ok")
