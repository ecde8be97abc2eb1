"""ncu target: the three kernels of the config-4 training step at D = 4096 (forward, TMEM backward, fused loss layer),
32 samples x 8192 rows per launch like bench.py."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from whvi_b200 import functional as F  # noqa: E402
D, S, B = 4096, 8, 4096
dev = torch.device("cuda:0")
x = torch.randn(S, B, D, device=dev); dy = torch.randn(S, B, D, device=dev); g = torch.randn(S, D, device=dev)
s1, s2 = torch.randn(D, device=dev), torch.randn(D, device=dev)
tgt = torch.randn(B, D, device=dev); y = torch.empty_like(x)
for _ in range(2):
    F.layer_forward_raw(x, g, s1, s2, out=y, relu_out=True)
    F.layer_backward_raw(x, dy, g, s1, s2, want_dx=True, relu_in=True)
    F.layer_loss_raw(x, g, s1, s2, None, tgt, want_dx=True, relu_in=True)
torch.cuda.synchronize()
print("ok")
