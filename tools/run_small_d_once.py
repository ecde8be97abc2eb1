"""ncu target: forward + backward at D = 16, 128, 1024, 8192 (2^26 elements each)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from whvi_b200 import functional as F  # noqa: E402
dev = torch.device("cuda:0")
for D in (16, 128, 1024, 8192):
    S, B = 8, (1 << 26) // (8 * D)
    x = torch.randn(S, B, D, device=dev); dy = torch.randn(S, B, D, device=dev); g = torch.randn(S, D, device=dev)
    s1, s2 = torch.randn(D, device=dev), torch.randn(D, device=dev)
    y = torch.empty_like(x)
    for _ in range(2):
        F.layer_forward_raw(x, g, s1, s2, out=y, relu_out=True)
        F.layer_backward_raw(x, dy, g, s1, s2, want_dx=True, relu_in=True)
torch.cuda.synchronize()
print("ok")
