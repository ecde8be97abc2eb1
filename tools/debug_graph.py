"""Bisect helper: the toy model's CUDA-graph training loop with ragged batches, one toggle set per process.
    python tools/debug_graph.py [nostack] [nocol] [nofuse] [eager]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import whvi_b200 as W  # noqa: E402
from whvi_b200.graphs import GraphedStepCache  # noqa: E402

flags = set(sys.argv[1:])
torch.manual_seed(11)
model = W.WHVIRegression([W.WHVILinear(3, 16, lambda_=2.0), torch.nn.ReLU(), W.WHVILinear(16, 16, lambda_=2.0),
                          torch.nn.ReLU(), W.WHVILinear(16, 1)], train_samples=4).cuda()
for m in model._whvi_layers():
    w = m.weight_submodule
    if isinstance(w, W.WHVIStackedMatrix) and "nostack" in flags:
        w.one_launch = False
    if isinstance(w, W.WHVIColumnMatrix) and "nocol" in flags:
        w.one_launch = False
if "nofuse" in flags:
    model.fuse = False
x = torch.randn(150, 3, device="cuda")
y = x[:, :1] + x[:, 1:2] ** 2
opt = torch.optim.Adam(model.parameters(), lr=torch.tensor(2e-2, device="cuda"), capturable=True)
model.train()
sizes = [int(a) for a in flags if a.isdigit()] or [64, 64, 22]
step = None if "eager" in flags else GraphedStepCache(model, opt, n=150, scheduled=True)
lo = 0
for it in range(12):
    b = sizes[it % len(sizes)]
    xb, yb = x[:b].clone(), y[:b].clone()
    if step is None:
        loss = model.loss(xb, yb, n=150)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
    else:
        loss = step(xb, yb)
    torch.cuda.synchronize()
    print(it, b, float(loss), flush=True)
print("OK", sorted(flags))
