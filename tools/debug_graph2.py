"""Bisect helper 2: tests/test_modules_gpu.py::test_train_model_cuda_graph's body with toggles.
    python tools/debug_graph2.py [nostack] [nocol] [nofuse] [noeval] [nosched]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import whvi_b200 as W  # noqa: E402
import whvi_b200.graphs as G  # noqa: E402

flags = set(sys.argv[1:])
torch.manual_seed(11)
model = W.WHVIRegression([W.WHVILinear(3, 16, lambda_=2.0), torch.nn.ReLU(), W.WHVILinear(16, 16, lambda_=2.0),
                          torch.nn.ReLU(), W.WHVILinear(16, 1)], train_samples=4).cuda()
with torch.no_grad():
    for name, p in model.named_parameters():
        if name.endswith(("s1", "s2", "g_mu")):
            p.normal_()
for m in model._whvi_layers():
    w = m.weight_submodule
    if isinstance(w, W.WHVIStackedMatrix) and "nostack" in flags:
        w.one_launch = False
    if isinstance(w, W.WHVIColumnMatrix) and "nocol" in flags:
        w.one_launch = False
if "nofuse" in flags:
    model.fuse = False
torch.manual_seed(0)
x = torch.randn(150, 3, device="cuda")
y = x[:, :1] + x[:, 1:2] ** 2
ds = torch.utils.data.TensorDataset(x, y)
loader = torch.utils.data.DataLoader(ds, batch_size=64)
opt = torch.optim.Adam(model.parameters(), lr=torch.tensor(2e-2, device="cuda"), capturable=True)
sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda t: 1.0 / (1.0 + 1e-3 * t))
if "noeval" not in flags:
    model.eval()
    print("before", model.eval_model(x, y), flush=True)
# synchronise after every replay to find the first bad one
orig = G.GraphedTrainStep.__call__
count = [0]
def checked(self, xx, yy):
    out = orig(self, xx, yy)
    torch.cuda.synchronize()
    count[0] += 1
    return out
G.GraphedTrainStep.__call__ = checked
try:
    model.train_model(loader, opt, sched, epochs1=20, epochs2=20, cuda_graph=True)
    print("after", model.eval_model(x, y), "steps", count[0], flush=True)
    print("OK", sorted(flags))
except Exception as e:
    print("FAILED at replay", count[0], type(e).__name__, str(e)[:200], flush=True)
