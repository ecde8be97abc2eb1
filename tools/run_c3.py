"""BASELINE config 3 (UCI-shaped MLP 13-128-128-1, B=4096, S=64): a few full training steps, for profiling."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import whvi_b200 as W  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
B, S = 4096, 64
model = W.WHVIRegression([W.WHVILinear(13, 128, lambda_=3.0), torch.nn.ReLU(), W.WHVILinear(128, 128, lambda_=3.0),
                          torch.nn.ReLU(), W.WHVILinear(128, 1, lambda_=3.0)], train_samples=S).to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
x, y = torch.randn(B, 13, device=dev), torch.randn(B, 1, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    loss = model.loss(x, y, n=B)
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
torch.cuda.synchronize()
print("ok", float(loss))
