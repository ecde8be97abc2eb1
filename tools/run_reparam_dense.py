"""Time / profile kernel (4), the tcgen05 dense reparameterisation GEMMs, and kernel (5) dense (KL with log-det).
    python tools/run_reparam_dense.py [D] [S] [--once]     (--once: one call of each, for ncu)"""
import json
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from whvi_b200 import functional as F  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
D = int(args[0]) if len(args) > 0 else 4096
S = int(args[1]) if len(args) > 1 else 128
once = "--once" in sys.argv
dev = torch.device("cuda:0")
mu, eps = torch.randn(D, device=dev), torch.randn(S, D, device=dev)
L = torch.randn(D, D, device=dev)
L.tril_().div_(D ** 0.5)
L.diagonal().abs_().add_(0.05)
dg = torch.randn(S, D, device=dev)
mu_r, L_r = mu.clone().requires_grad_(), L.clone().requires_grad_()


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / reps


def fwd():
    with torch.no_grad():
        return F.reparam_dense(mu, L, eps)


def fwd_bwd():
    L_r.grad = None
    mu_r.grad = None
    F.reparam_dense(mu_r, L_r, eps).backward(dg)


def kl():
    L_r.grad = None
    mu_r.grad = None
    F.kl_gaussian_dense(mu_r, L_r, 0.1).backward()


if once:
    fwd(); fwd_bwd(); kl()
    torch.cuda.synchronize()
    sys.exit(0)

reps = 20 if D <= 8192 else 5
out = {"D": D, "S": S}
out["fwd_ms"] = timed(fwd, reps)
out["fwd_bwd_ms"] = timed(fwd_bwd, reps)
out["kl_fwd_bwd_ms"] = timed(kl, reps)
flops = 3 * 2.0 * S * D * (D + 128) / 2   # three tf32 MMAs per product, lower triangle only
out["fwd_tflops_tf32_issued"] = flops / out["fwd_ms"] / 1e9
out["fwd_hbm_floor_ms"] = (4.0 * D * (D + 128) / 2 + 8.0 * S * D) / 6552.6e9 * 1e3   # L's triangle once + eps in + g out
g = fwd()
ref = torch.addmm(mu.double(), eps.double(), L.double().t()) if D <= 8192 else None
if ref is not None:
    out["fwd_rel_err_vs_fp64"] = ((g.double() - ref).abs().max() / ref.abs().max()).item()
    fwd_bwd()
    dl_ref = torch.tril(dg.double().t() @ eps.double())
    out["bwd_rel_err_vs_fp64"] = ((L_r.grad.double() - dl_ref).abs().max() / dl_ref.abs().max()).item()
torch.backends.cuda.matmul.allow_tf32 = False
out["cublas_fp32_addmm_full_ms"] = timed(lambda: torch.addmm(mu, eps, L.t()), reps)
out["cublas_fp32_dL_full_ms"] = timed(lambda: torch.tril(dg.t() @ eps), reps)
print(json.dumps(out))
