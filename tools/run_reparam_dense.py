"""Time / profile kernel (4), the tcgen05 dense reparameterisation GEMM.
    python tools/run_reparam_dense.py [D] [S]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from whvi_b200 import functional as F  # noqa: E402
D = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda:0")
mu, eps = torch.randn(D, device=dev), torch.randn(S, D, device=dev)
L = torch.tril(torch.randn(D, D, device=dev)) / D ** 0.5
for _ in range(3):
    g = F.reparam_dense(mu, L, eps)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    g = F.reparam_dense(mu, L, eps)
b.record(); b.synchronize()
ms = a.elapsed_time(b) / 10
flops = 3 * 2.0 * S * D * (D + 128) / 2  # 3 tf32 MMAs per product, lower triangle only
ref = mu + eps @ L.t()
err = ((g - ref).abs().max() / ref.abs().max()).item()
torch.backends.cuda.matmul.allow_tf32 = False
c, d = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
c.record()
for _ in range(10):
    ref = torch.addmm(mu, eps, L.t())
d.record(); d.synchronize()
print(f"D={D} S={S}: {ms:.3f} ms, {flops / ms / 1e9:.1f} TFLOP/s tf32 issued (useful {flops / 3 / ms / 1e9:.1f}), "
      f"rel diff vs torch fp32 addmm {err:.2e}; cuBLAS fp32 addmm (full matrix) {c.elapsed_time(d) / 10:.3f} ms")
