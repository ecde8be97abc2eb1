"""Generate golden fixtures by running the REFERENCE itself (CPU) in this container.

    python tests/golden/make_golden.py          # writes tests/golden/*.npz

Needs ``/root/reference`` (read-only mount) and ``oracle/_ref/fwht_cpp.so`` (built from
the reference's unmodified ``src/fwht/cpp/fwht.cpp`` by ``oracle/build.py``).  The
reference cannot travel to the GPU box, so the outputs are committed as small ``.npz``
fixtures; tests only ever read the fixtures.

What is recorded (all from reference code paths, nothing from this repo):
  fwht.npz      known-answer vectors of test/walsh.py:12-20; fwht_cpp / python
                FWHTFunction / WHT_matmul outputs on seeded inputs (D = 4, 32, 1024, 4096)
  kl.npz        WHVISquarePow2Matrix.kl values + autograd grads; utils.kl_diag_normal
  mnll.npz      GaussianLikelihood.mnll_batch_estimate incl. the fixed case of
                test/likelihoods.py:8-31
  layers.npz    reference-AS-WRITTEN forward outputs and all parameter/input grads of
                Square (D=16), Stacked (3->16, 13->32 with bias) and Column (16->1, 1->8)
                layers with the eps draws captured (monkeypatched torch.randn)
  init.npz      state_dict of a freshly constructed model under torch.manual_seed(0)
  toy.npz       README toy model (README.md:25-44): loss, KL, MNLL and every grad for one
                batch with captured eps; state_dict key names
  train.npz     three Adam steps of the toy model as written: per-step losses, captured eps,
                parameters before and after
  dims.npz      setup_dimensions on a 199 x 200 grid; WHVILinear's weight class and parameter
                count for every (n_in, n_out) <= 40
  paper.npz     PAPER-semantics layer (docstring src/weights.py:77) from a dense fp64
                H-matrix formula using the reference's own build_H, with autograd grads
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path("/root/reference")

sys.path.insert(0, str(REF))
sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
# src/weights.py:8 imports the CUDA extension unconditionally; stub it.
sys.modules.setdefault("fwht_cuda", types.ModuleType("fwht_cuda"))

import fwht_cpp  # noqa: E402  (reference C++ FWHT, compiled unmodified)
import src.fwht.cpp.fwht as cpp_fwht  # noqa: E402
import src.fwht.python.fwht as python_fwht  # noqa: E402
from src.layers import WHVILinear  # noqa: E402
from src.likelihoods import GaussianLikelihood  # noqa: E402
from src.networks import WHVIRegression  # noqa: E402
from src.utils import build_H, kl_diag_normal  # noqa: E402
from src.weights import WHVISquarePow2Matrix  # noqa: E402


class CaptureRandn:
    """Record every torch.randn draw made inside the reference (weights.py:81,:92)."""

    def __init__(self):
        self.draws = []

    def __enter__(self):
        self._orig = torch.randn

        def randn(*a, **k):
            out = self._orig(*a, **k)
            self.draws.append(out.detach().clone())
            return out

        torch.randn = randn
        return self

    def __exit__(self, *exc):
        torch.randn = self._orig


def npy(t):
    return t.detach().cpu().numpy()


def gen_fwht():
    out = {}
    out["kat_in"] = np.array([[1.0, 2.0, 3.0, 4.0], [0.0, 1.0, 2.0, 3.0]], dtype=np.float32)
    out["kat_out_expected"] = np.array([[10.0, -2.0, -4.0, 0.0], [6.0, -2.0, -4.0, 0.0]], dtype=np.float32)
    out["kat_out_cpp"] = npy(cpp_fwht.FWHTFunction.apply(torch.tensor(out["kat_in"])))
    g = torch.Generator().manual_seed(1234)
    for D, B in ((4, 2), (32, 40), (64, 7), (1024, 19), (4096, 3)):
        a = torch.randn(B, D, generator=g)
        out[f"in_{D}"] = npy(a)
        out[f"cpp_{D}"] = npy(fwht_cpp.forward(a))
        out[f"py_{D}"] = npy(python_fwht.FWHTFunction.apply(a))
        if D <= 1024:
            out[f"matmul_{D}"] = npy(python_fwht.WHT_matmul().apply(a))
            H = build_H(D, torch.device("cpu")).double()
            out[f"dense64_{D}"] = npy((H @ a.double().T).T)
    np.savez_compressed(HERE / "fwht.npz", **out)


def gen_kl():
    out = {}
    torch.manual_seed(7)
    for idx, (D, lam) in enumerate(((16, 2.0), (128, 3.0), (64, 1e-5), (8, 0.37))):
        m = WHVISquarePow2Matrix(D, lambda_=lam)
        with torch.no_grad():
            m.g_mu.copy_(torch.randn(D))
            m.g_rho.copy_(torch.randn(D) * 2.0)
        kl = m.kl
        kl.backward()
        out[f"D_{idx}"], out[f"lam_{idx}"] = np.int64(D), np.float64(lam)
        out[f"mu_{idx}"], out[f"rho_{idx}"] = npy(m.g_mu), npy(m.g_rho)
        out[f"kl_{idx}"] = npy(kl)
        out[f"dmu_{idx}"], out[f"drho_{idx}"] = npy(m.g_mu.grad), npy(m.g_rho.grad)
    # the general 4-argument form, as test/utils.py:22-34 exercises it
    mu1, sd1, mu2, sd2 = torch.randn(10), torch.exp(torch.randn(10)), torch.randn(10), torch.exp(torch.randn(10))
    out["gen_mu1"], out["gen_sd1"], out["gen_mu2"], out["gen_sd2"] = map(npy, (mu1, sd1, mu2, sd2))
    out["gen_kl"] = npy(kl_diag_normal(mu1, sd1, mu2, sd2))
    out["gen_kl_torch"] = npy(torch.distributions.kl.kl_divergence(
        torch.distributions.MultivariateNormal(mu1, torch.diag(sd1)),
        torch.distributions.MultivariateNormal(mu2, torch.diag(sd2))))
    np.savez_compressed(HERE / "kl.npz", **out)


def gen_mnll():
    out = {}
    y = torch.reshape(torch.tensor([0.0, 1.0, 2.0, -1.0]), (-1, 1))
    y_hat = torch.tensor([[0.2, 1.1, 2.2, -1.3], [-0.1, 1.05, 2, -1.1]]).T.unsqueeze(1)
    out["fixed_y"], out["fixed_yhat"] = npy(y), npy(y_hat)
    out["fixed_mnll"] = npy(GaussianLikelihood(sigma=1.0).mnll_batch_estimate(y, y_hat, 12))
    torch.manual_seed(3)
    y, y_hat = torch.randn(24, 1), torch.randn(24, 1, 80)
    out["rand_y"], out["rand_yhat"] = npy(y), npy(y_hat)
    out["rand_mnll"] = npy(GaussianLikelihood(sigma=15.21).mnll_batch_estimate(y, y_hat, 116))
    y, y_hat = torch.randn(9, 3), torch.randn(9, 3, 5)
    out["multi_y"], out["multi_yhat"] = npy(y), npy(y_hat)
    out["multi_mnll"] = npy(GaussianLikelihood(sigma=0.7).mnll_batch_estimate(y, y_hat, 40))
    np.savez_compressed(HERE / "mnll.npz", **out)


def randomise(layer: torch.nn.Module, seed: int):
    """O(1) parameters so that relative errors are meaningful."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in layer.named_parameters():
            if name.endswith("g_rho"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.5 - 1.0)
            else:
                p.copy_(torch.randn(p.shape, generator=g))


def gen_layers():
    out = {}
    cases = {
        "square16": (16, 16, False, 6),
        "square16_bias": (16, 16, True, 6),
        "stacked_3_16": (3, 16, False, 5),
        "stacked_13_32_bias": (13, 32, True, 4),
        "column_16_1": (16, 1, False, 7),
        "column_1_8_bias": (1, 8, True, 7),
    }
    for k, (name, (n_in, n_out, bias, B)) in enumerate(cases.items()):
        layer = WHVILinear(n_in, n_out, lambda_=2.0, bias=bias)
        randomise(layer, 100 + k)
        x = torch.randn(B, n_in, generator=torch.Generator().manual_seed(200 + k), requires_grad=True)
        dy = torch.randn(B, n_out, generator=torch.Generator().manual_seed(300 + k))
        with CaptureRandn() as cap:
            y = layer(x)
        (y * dy).sum().backward()
        out[f"{name}.shape"] = np.array([n_in, n_out, int(bias), B], dtype=np.int64)
        out[f"{name}.x"], out[f"{name}.dy"], out[f"{name}.y"] = npy(x), npy(dy), npy(y)
        out[f"{name}.dx"] = npy(x.grad)
        out[f"{name}.n_eps"] = np.int64(len(cap.draws))
        for i, e in enumerate(cap.draws):
            out[f"{name}.eps{i}"] = npy(e)
        for pname, p in layer.named_parameters():
            out[f"{name}.param.{pname}"] = npy(p)
            out[f"{name}.grad.{pname}"] = npy(p.grad) if p.grad is not None else np.zeros(p.shape, np.float32)
        out[f"{name}.kl"] = npy(layer.kl)
    np.savez_compressed(HERE / "layers.npz", **out)


def gen_toy():
    """README.md:25-44 toy regression, one training batch, reference as written."""
    out = {}
    torch.manual_seed(0)
    x = torch.randn(200, 3)
    y = torch.reshape(x[:, 0] + x[:, 1] ** 2 - 0.3 * x[:, 2] ** 3, (-1, 1))
    model = WHVIRegression([WHVILinear(3, 16, lambda_=2.0), torch.nn.ReLU(), WHVILinear(16, 1)], train_samples=3)
    randomise(model.sequential, 11)
    model.train()
    xb, yb = x[:64], y[:64]
    with CaptureRandn() as cap:
        loss = model.loss(xb, yb, n=150)
    loss.backward()
    out["x"], out["y"] = npy(xb), npy(yb)
    out["loss"], out["kl"], out["mnll"] = npy(loss), npy(model.current_kl), npy(model.current_mnll)
    out["n_eps"] = np.int64(len(cap.draws))
    for i, e in enumerate(cap.draws):
        out[f"eps{i}"] = npy(e)
    keys = []
    for name, p in model.named_parameters():
        keys.append(name)
        out[f"param.{name}"] = npy(p)
        out[f"grad.{name}"] = npy(p.grad)
    out["state_dict_keys"] = np.array(list(model.state_dict().keys()))
    # eval-mode forward shape/values with 4 samples
    model.eval_samples = 4
    model.eval()
    with CaptureRandn() as cap, torch.no_grad():
        pred = model(x[150:160])
    out["eval_x"], out["eval_pred"] = npy(x[150:160]), npy(pred)
    out["eval_n_eps"] = np.int64(len(cap.draws))
    for i, e in enumerate(cap.draws):
        out[f"eval_eps{i}"] = npy(e)
    # WHVIRegression.eval_model (src/networks.py:101-116, :130-132): RMSE of the MC mean and test MNLL
    with CaptureRandn() as cap:
        rmse, mnll = model.eval_model(x[160:200], y[160:200])
    out["evalm_x"], out["evalm_y"] = npy(x[160:200]), npy(y[160:200])
    out["evalm_rmse"], out["evalm_mnll"] = np.float64(rmse), np.float64(mnll)
    out["evalm_n_eps"] = np.int64(len(cap.draws))
    for i, e in enumerate(cap.draws):
        out[f"evalm_eps{i}"] = npy(e)
    np.savez_compressed(HERE / "toy.npz", **out)


def gen_train():
    """Three optimisation steps of the README toy model with the reference as written: losses per
    step and every parameter afterwards (fixed minibatch), eps captured per step.  Plain SGD with a
    step size scaled to the first gradient: Adam would turn the reference's rounding-noise gradients
    (exactly zero in exact arithmetic) into full-size steps and make the comparison
    ill-conditioned."""
    out = {}
    torch.manual_seed(0)
    x = torch.randn(200, 3)
    y = torch.reshape(x[:, 0] + x[:, 1] ** 2 - 0.3 * x[:, 2] ** 3, (-1, 1))
    model = WHVIRegression([WHVILinear(3, 16, lambda_=2.0), torch.nn.ReLU(), WHVILinear(16, 1)], train_samples=2)
    randomise(model.sequential, 23)
    model.train()
    for name, p in model.named_parameters():
        out[f"init.{name}"] = npy(p).copy()       # npy() aliases the parameter, which SGD updates in place
    xb, yb = x[:64], y[:64]
    model.loss(xb, yb, n=150).backward()          # probe gradient scale (its eps draws are not recorded)
    gmax = max(float(p.grad.abs().max()) for p in model.parameters())
    model.zero_grad()
    lr = 0.02 / gmax
    out["lr"] = np.float64(lr)
    opt = torch.optim.SGD(model.parameters(), lr=lr)
    losses = []
    for step in range(3):
        with CaptureRandn() as cap:
            loss = model.loss(xb, yb, n=150)
        loss.backward()
        opt.step()
        opt.zero_grad()
        losses.append(float(loss))
        out[f"n_eps_{step}"] = np.int64(len(cap.draws))
        for i, e in enumerate(cap.draws):
            out[f"eps_{step}_{i}"] = npy(e)
    out["x"], out["y"], out["losses"] = npy(xb), npy(yb), np.array(losses)
    for name, p in model.named_parameters():
        out[f"final.{name}"] = npy(p)
    np.savez_compressed(HERE / "train.npz", **out)


def gen_dims():
    """WHVIStackedMatrix.setup_dimensions (src/weights.py:135-160) on a grid, and which weight class /
    how many parameters WHVILinear builds for every (n_in, n_out) <= 40 (src/layers.py:31-38)."""
    from src.weights import WHVIStackedMatrix
    ins, outs = np.arange(2, 201), np.arange(1, 201)
    table = np.zeros((len(ins), len(outs), 4), dtype=np.int64)
    for i, a in enumerate(ins):
        for j, b in enumerate(outs):
            table[i, j] = WHVIStackedMatrix.setup_dimensions(int(a), int(b))
    kinds = np.zeros((40, 40), dtype=np.int64)      # 0 column, 1 column transposed, 2 square, 3 stacked
    counts = np.zeros((40, 40), dtype=np.int64)
    names = {"WHVIColumnMatrix": 0, "WHVISquarePow2Matrix": 2, "WHVIStackedMatrix": 3}
    for a in range(1, 41):
        for b in range(1, 41):
            layer = WHVILinear(a, b)
            sub = layer.weight_submodule
            k = names[type(sub).__name__]
            if k == 0 and getattr(sub, "transposed", False):
                k = 1
            kinds[a - 1, b - 1] = k
            counts[a - 1, b - 1] = sum(p.numel() for p in layer.parameters())
    np.savez_compressed(HERE / "dims.npz", ins=ins, outs=outs, table=table, kinds=kinds, counts=counts)


def gen_paper():
    """PAPER semantics y = x @ (S1 H diag(g) H S2)^T in dense fp64 with the reference's
    own build_H (src/utils.py:74-101); grads from torch autograd."""
    out = {}
    for idx, (S, B, D) in enumerate(((3, 5, 16), (2, 4, 128), (4, 3, 1024))):
        gen = torch.Generator().manual_seed(900 + idx)
        H = build_H(D, torch.device("cpu")).double()
        x = torch.randn(S, B, D, generator=gen, dtype=torch.float64, requires_grad=True)
        s1 = torch.randn(D, generator=gen, dtype=torch.float64, requires_grad=True)
        s2 = torch.randn(D, generator=gen, dtype=torch.float64, requires_grad=True)
        mu = torch.randn(D, generator=gen, dtype=torch.float64, requires_grad=True)
        rho = (torch.randn(D, generator=gen, dtype=torch.float64) * 0.5 - 1.0).requires_grad_()
        eps = torch.randn(S, D, generator=gen, dtype=torch.float64)
        dy = torch.randn(S, B, D, generator=gen, dtype=torch.float64)
        g = mu + torch.nn.functional.softplus(rho) * eps
        g.retain_grad()
        ys = []
        for s in range(S):
            W = torch.diag(s1) @ H @ torch.diag(g[s]) @ H @ torch.diag(s2)
            ys.append(x[s] @ W.T)
        y = torch.stack(ys)
        (y * dy).sum().backward()
        for k, v in dict(x=x, s1=s1, s2=s2, mu=mu, rho=rho, eps=eps, dy=dy, g=g, y=y, dx=x.grad, ds1=s1.grad,
                         ds2=s2.grad, dmu=mu.grad, drho=rho.grad, dg=g.grad).items():
            out[f"{k}_{idx}"] = npy(v)
    np.savez_compressed(HERE / "paper.npz", **out)


def gen_init():
    """Parameter initialisation under a fixed seed (src/weights.py:28-32 draw order)."""
    out = {}
    torch.manual_seed(0)
    model = WHVIRegression([WHVILinear(3, 16, lambda_=2.0), torch.nn.ReLU(), WHVILinear(16, 16, bias=True),
                            torch.nn.ReLU(), WHVILinear(16, 1)])
    for k, v in model.state_dict().items():
        out[k] = npy(v)
    out["__keys__"] = np.array(list(model.state_dict().keys()))
    np.savez_compressed(HERE / "init.npz", **out)


if __name__ == "__main__":
    if len(sys.argv) > 1:      # e.g. `make_golden.py toy`: regenerate only the named fixtures
        for name in sys.argv[1:]:
            globals()[f"gen_{name}"]()
        sys.exit(0)
    gen_init()
    gen_fwht()
    gen_kl()
    gen_mnll()
    gen_layers()
    gen_toy()
    gen_train()
    gen_dims()
    gen_paper()
    for f in sorted(HERE.glob("*.npz")):
        print(f.name, f.stat().st_size, "bytes")
