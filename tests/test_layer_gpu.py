"""GPU parity of kernels (2) fused forward, (3) fused backward, the reparameterisation and
(5) the KL reduction, through the C ABI, against the fp64 CPU oracle and the goldens.
Tolerances (north_star): layer outputs and gradients <= 1e-4 in max|a-b|/max|b|."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-4


def dev():
    return torch.device("cuda:0")


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev())


def make_case(S, B, D, seed, shared=False):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, D) if shared else (S, B, D))
    g = rng.standard_normal((S, D))
    s1, s2 = rng.standard_normal(D), rng.standard_normal(D)
    dy = rng.standard_normal((S, B, D))
    bias = rng.standard_normal(D)
    return x, g, s1, s2, dy, bias


CASES = [(1, 1, 4), (3, 5, 4), (2, 300, 8), (3, 67, 16), (2, 64, 16), (4, 33, 64), (3, 9, 128), (2, 256, 128),
         (2, 7, 512), (3, 5, 1024), (2, 3, 2048), (2, 5, 4096), (2, 3, 8192), (5, 1, 1024), (2, 3, 16384), (2, 2, 32768)]


@pytest.mark.parametrize("S,B,D", CASES)
@pytest.mark.parametrize("shared", [False, True])
def test_forward_backward_vs_oracle(S, B, D, shared):
    from whvi_b200 import functional as F
    x, g, s1, s2, dy, bias = make_case(S, B, D, 17 * D + S + B, shared)
    y = F.layer_forward_raw(t(x), t(g), t(s1), t(s2), t(bias))
    y_ref = O.layer_fwd(x, g, s1, s2, bias)
    assert rel_err(y.cpu().numpy(), y_ref) < TOL
    dx, dg, ds1, ds2, db = F.layer_backward_raw(t(x), t(dy), t(g), t(s1), t(s2), want_dx=True, want_dbias=True)
    rdx, rdg, rds1, rds2, rdb = O.layer_bwd(x, dy, g, s1, s2, want_dbias=True)
    dx = dx.cpu().numpy()
    if shared:
        dx = dx.astype(np.float64).sum(0)
    assert rel_err(dx, rdx) < TOL
    assert rel_err(dg.cpu().numpy(), rdg) < TOL
    assert rel_err(ds1.cpu().numpy(), rds1) < TOL
    assert rel_err(ds2.cpu().numpy(), rds2) < TOL
    assert rel_err(db.cpu().numpy(), rdb) < TOL


@pytest.mark.parametrize("S,B,D", [(2, 37, 16), (3, 70, 128), (2, 9, 1024), (2, 6, 2048), (2, 5, 4096), (2, 3, 8192)])
def test_fused_flags(S, B, D):
    """N1/N2 fusions: ReLU on the output, ReLU mask on dx, Gaussian-MNLL residual as dy and
    its squared-error reduction."""
    from whvi_b200 import functional as F
    x, g, s1, s2, dy, bias = make_case(S, B, D, 3 * D + B)
    rng = np.random.default_rng(D)
    target = rng.standard_normal((B, D))
    y_ref = O.layer_fwd(x, g, s1, s2, bias)
    y, sq = F.layer_forward_raw(t(x), t(g), t(s1), t(s2), t(bias), relu_out=True, target=t(target))
    yr = np.maximum(y_ref, 0.0)
    assert rel_err(y.cpu().numpy(), yr) < TOL
    sq_ref = float(((yr - target[None]) ** 2).sum())
    assert abs(sq.item() - sq_ref) < 1e-4 * sq_ref
    # backward: x plays the role of a ReLU output (mask = x > 0); dy formed from a residual
    coef = 0.37
    yhat = rng.standard_normal((S, B, D))
    dy_explicit = coef * (yhat - target[None])
    dx, dg, ds1, ds2, db = F.layer_backward_raw(t(x), t(yhat), t(g), t(s1), t(s2), want_dx=True, want_dbias=True,
                                                 relu_in=True, target=t(target), coef=torch.tensor(coef, device=dev()))
    x32 = x.astype(np.float32)
    rdx, rdg, rds1, rds2, rdb = O.layer_bwd(x, dy_explicit, g, s1, s2, want_dbias=True)
    assert rel_err(dx.cpu().numpy(), rdx * (x32 > 0)) < TOL
    assert rel_err(dg.cpu().numpy(), rdg) < TOL
    assert rel_err(ds1.cpu().numpy(), rds1) < TOL
    assert rel_err(ds2.cpu().numpy(), rds2) < TOL
    assert rel_err(db.cpu().numpy(), rdb) < TOL


@pytest.mark.parametrize("S,B,D", [(2, 3, 16384), (2, 2, 32768)])
def test_forward_large_dims(S, B, D):
    """D = 2^14, 2^15 (BASELINE config 5 width): fused forward; the backward there is the multi-pass composition of the
    FWHT kernel (functional._layer_backward_multipass), with the fused ReLU mask and deferred scale of the fused path."""
    from whvi_b200 import functional as F
    x, g, s1, s2, dy, bias = make_case(S, B, D, D + S)
    y = F.layer_forward_raw(t(x), t(g), t(s1), t(s2), t(bias), relu_out=False)
    assert rel_err(y.cpu().numpy(), O.layer_fwd(x, g, s1, s2, bias)) < TOL
    ys = F.layer_forward_raw(t(x[0]), t(g), t(s1), t(s2))
    assert rel_err(ys.cpu().numpy(), O.layer_fwd(x[0], g, s1, s2)) < TOL
    sc = torch.tensor([0.37], device=dev())
    dx, dg, ds1, ds2, _ = F.layer_backward_raw(t(x), t(dy), t(g), t(s1), t(s2), relu_in=True, dy_scale=sc)
    rdx, rdg, rds1, rds2 = O.layer_bwd(x, 0.37 * dy, g, s1, s2)
    assert rel_err(dx.cpu().numpy(), rdx * (x.astype(np.float32) > 0)) < TOL
    for got, ref in ((dg, rdg), (ds1, rds1), (ds2, rds2)):
        assert rel_err(got.cpu().numpy(), ref) < TOL


@pytest.mark.parametrize("S,B,D", [(2, 5, 4), (3, 7, 64), (2, 9, 128), (2, 5, 1024), (2, 3, 2048), (2, 3, 4096), (2, 3, 8192),
                                   (2, 3, 16384), (3, 2, 32768)])
def test_forward_from_t2(S, B, D):
    """FROM_T2: the sample-independent first transform hoisted out -- same result as
    the full chain, checked against the oracle."""
    from whvi_b200 import functional as F
    from whvi_b200.fwht import fwht_
    x, g, s1, s2, dy, bias = make_case(S, B, D, 3 * D + S, shared=True)
    t2 = fwht_(t(x) * t(s2))
    for b in (None, bias):
        y = F.layer_forward_raw(t2, t(g), t(s1), t(s2), None if b is None else t(b), from_t2=True)
        assert rel_err(y.cpu().numpy(), O.layer_fwd(x, g, s1, s2, b)) < TOL
    with pytest.raises(RuntimeError, match="FROM_T2"):
        F.layer_forward_raw(t2, t(g), t(s1), t(s2), target=t(x), from_t2=True)


@pytest.mark.parametrize("S,n", [(1, 4), (5, 1028), (16, 40000), (0, 64), (7, 12)])
def test_mc_moments(S, n):
    from whvi_b200 import functional as F
    rng = np.random.default_rng(S * 1000 + n)
    y = rng.standard_normal((S, n)).astype(np.float32)
    a0, b0 = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    a, b = t(a0), t(b0)
    F.mc_moments_(t(y).reshape(S, n), a, b, accumulate=True)
    y64 = y.astype(np.float64)
    assert rel_err(a.cpu().numpy(), a0 + y64.sum(0)) < 1e-6
    assert rel_err(b.cpu().numpy(), b0 + (y64 * y64).sum(0)) < 1e-6
    F.mc_moments_(t(y).reshape(S, n), a, None, accumulate=False)
    assert rel_err(a.cpu().numpy(), y64.sum(0)) < 1e-6 or S == 0
    with pytest.raises(RuntimeError):
        F.mc_moments_(t(y).reshape(S, n), a[:-1].clone(), None)


def test_mc_moments_strided_row_blocks():
    """whvi_mc_moments_strided_f32: out = in + sum_s y on a block of rows of a larger (S, B, D) tensor
    (the form the sample-sharded evaluation uses to deliver partial sums to the rows' owner)."""
    from whvi_b200 import functional as F
    rng = np.random.default_rng(11)
    S, B, D = 5, 12, 64
    y = rng.standard_normal((S, B, D)).astype(np.float32)
    a0, b0 = rng.standard_normal((B, D)).astype(np.float32), rng.standard_normal((B, D)).astype(np.float32)
    yt, a, b = t(y), t(a0), t(b0)
    oa, ob = torch.full((B, D), 7.0, device=dev()), torch.full((B, D), 7.0, device=dev())
    for r0, r1 in ((0, 4), (4, 8), (8, 12)):
        F.mc_moments_into(yt[:, r0:r1], a[r0:r1], b[r0:r1], oa[r0:r1], ob[r0:r1])
    y64 = y.astype(np.float64)
    assert rel_err(oa.cpu().numpy(), a0 + y64.sum(0)) < 1e-6
    assert rel_err(ob.cpu().numpy(), b0 + (y64 * y64).sum(0)) < 1e-6
    F.mc_moments_into(yt[:, 2:5], None, None, oa[2:5], None)
    assert rel_err(oa[2:5].cpu().numpy(), y64[:, 2:5].sum(0)) < 1e-6
    with pytest.raises(RuntimeError):
        F.mc_moments_into(yt[:, :, :32], None, None, oa[:, :32].contiguous(), None)   # rows not contiguous


@pytest.mark.parametrize("S,B,D,chunk", [(8, 5, 64, 3), (6, 3, 4096, 4), (5, 2, 32768, 2)])
def test_predictive_moments_vs_oracle(S, B, D, chunk):
    """BASELINE config 5 path: predictive mean/variance over MC samples without the (S,B,D) tensor."""
    import whvi_b200 as W
    rng = np.random.default_rng(D + S)
    layer = W.WHVISquarePow2Matrix(D, lambda_=1.0, bias=True).to(dev())
    with torch.no_grad():
        for p in (layer.s1, layer.s2, layer.g_mu):
            p.copy_(t(rng.standard_normal(D)))
        layer.bias.copy_(t(rng.standard_normal((1, D))))
    x, eps = rng.standard_normal((B, D)), rng.standard_normal((S, D))
    layer.inject_eps(t(eps))
    sy, sy2, n = layer.predictive_moments(t(x), chunk_samples=chunk)
    assert n == S
    p64 = {k: v.detach().cpu().numpy().astype(np.float64) for k, v in layer.named_parameters()}
    g = O.reparam(p64["g_mu"], p64["g_rho"], eps)
    y = O.layer_fwd(x, g, p64["s1"], p64["s2"], p64["bias"].reshape(-1))
    assert rel_err(sy.cpu().numpy(), y.sum(0)) < TOL
    assert rel_err(sy2.cpu().numpy(), (y * y).sum(0)) < TOL
    # sample shards add up (what reduce_predictive_moments all-reduces)
    layer.inject_eps(t(eps))
    a1, b1, n1 = layer.predictive_moments(t(x), chunk_samples=chunk, sample_range=(0, S // 2))
    layer.inject_eps(t(eps))
    a2, b2, n2 = layer.predictive_moments(t(x), chunk_samples=chunk, sample_range=(S // 2, S))
    assert n1 + n2 == S
    assert rel_err((a1 + a2).cpu().numpy(), y.sum(0)) < TOL
    # scatter_to: the last sample chunk's totals land in caller-provided row blocks
    layer.inject_eps(t(eps))
    oy, oy2 = torch.zeros(B, D, device=dev()), torch.zeros(B, D, device=dev())
    h = max(B // 2, 1)
    blocks = [(0, h, oy[:h], oy2[:h])] + ([(h, B, oy[h:], oy2[h:])] if B > h else [])
    layer.predictive_moments(t(x), chunk_samples=chunk, scatter_to=blocks)
    assert rel_err(oy.cpu().numpy(), y.sum(0)) < TOL and rel_err(oy2.cpu().numpy(), (y * y).sum(0)) < TOL


@pytest.mark.parametrize("S,B,D", [(5, 3, 8192), (4, 2, 16384), (7, 3, 32768), (1, 1, 8192)])
@pytest.mark.parametrize("mode", ["from_t2", "shared", "per_sample"])
def test_fused_layer_moments_vs_oracle(S, B, D, mode):
    """whvi_layer_moments_f32: sum_s y and sum_s y^2 with the running sums kept in tensor memory (no prediction in
    HBM), against the fp64 oracle's explicit predictions; with/without bias, overwrite and accumulate."""
    from whvi_b200 import functional as F
    from whvi_b200.fwht import fwht_
    rng = np.random.default_rng(D + 31 * S + len(mode))
    x = rng.standard_normal((S, B, D) if mode == "per_sample" else (B, D))
    g, s1, s2, bias = rng.standard_normal((S, D)), rng.standard_normal(D), rng.standard_normal(D), rng.standard_normal(D)
    y = O.layer_fwd(x, g, s1, s2, bias)
    xin = fwht_(t(x) * t(s2)) if mode == "from_t2" else t(x)
    sy, sy2 = torch.full((B, D), 7.0, device=dev()), torch.full((B, D), -3.0, device=dev())
    F.layer_moments_raw(xin, t(g), t(s1), t(s2), t(bias), sy, sy2, from_t2=mode == "from_t2")
    assert rel_err(sy.cpu().numpy(), y.sum(0)) < TOL
    assert rel_err(sy2.cpu().numpy(), (y * y).sum(0)) < TOL
    # accumulate a second sample chunk without bias on top
    y2 = O.layer_fwd(x, g[::-1].copy() if mode != "per_sample" else g, s1, s2)
    g2 = t(g[::-1].copy()) if mode != "per_sample" else t(g)
    F.layer_moments_raw(xin, g2, t(s1), t(s2), None, sy, sy2, from_t2=mode == "from_t2", accumulate=True)
    assert rel_err(sy.cpu().numpy(), y.sum(0) + y2.sum(0)) < TOL
    assert rel_err(sy2.cpu().numpy(), (y * y).sum(0) + (y2 * y2).sum(0)) < TOL
    with pytest.raises(RuntimeError):
        F.layer_moments_raw(t(x[..., :4096]), t(g[:, :4096]), t(s1[:4096]), t(s2[:4096]), None, sy[:, :4096].contiguous())


def test_backward_without_dx_and_bias_and_determinism():
    from whvi_b200 import functional as F
    x, g, s1, s2, dy, _ = make_case(3, 41, 256, 5)
    a = F.layer_backward_raw(t(x), t(dy), t(g), t(s1), t(s2), want_dx=False, want_dbias=False)
    b = F.layer_backward_raw(t(x), t(dy), t(g), t(s1), t(s2), want_dx=True, want_dbias=False)
    assert a[0] is None and a[4] is None
    for u, v in zip(a[1:4], b[1:4]):
        assert torch.equal(u, v)  # fixed reduction order => bit-reproducible


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_paper_golden(golden, idx):
    """Dense fp64 H-matrix formula with the reference's own build_H (tests/golden/paper.npz)."""
    from whvi_b200 import functional as F
    gd = golden("paper")
    x, s1, s2, mu, rho, eps, dy = (gd[f"{k}_{idx}"] for k in ("x", "s1", "s2", "mu", "rho", "eps", "dy"))
    xt = t(x).requires_grad_()
    s1t, s2t, mut, rhot = (t(v).requires_grad_() for v in (s1, s2, mu, rho))
    gg = F.reparam(mut, rhot, t(eps))
    y = F.whvi_layer(xt, gg, s1t, s2t)
    assert rel_err(gg.detach().cpu().numpy(), gd[f"g_{idx}"]) < 1e-6
    assert rel_err(y.detach().cpu().numpy(), gd[f"y_{idx}"]) < TOL
    (y * t(dy)).sum().backward()
    for name, tt in (("dx", xt), ("ds1", s1t), ("ds2", s2t), ("dmu", mut), ("drho", rhot)):
        assert rel_err(tt.grad.cpu().numpy(), gd[f"{name}_{idx}"]) < TOL, name


def test_shared_x_autograd_sums_over_samples():
    from whvi_b200 import functional as F
    x, g, s1, s2, dy, bias = make_case(4, 6, 64, 9, shared=True)
    xt = t(x).requires_grad_()
    bt = t(bias).requires_grad_()
    y = F.whvi_layer(xt, t(g), t(s1), t(s2), bt)
    (y * t(dy)).sum().backward()
    rdx, _, _, _, rdb = O.layer_bwd(x, dy, g, s1, s2, want_dbias=True)
    assert rel_err(xt.grad.cpu().numpy(), rdx) < TOL
    assert rel_err(bt.grad.cpu().numpy(), rdb) < TOL


def test_kl_matches_reference_golden(golden):
    from whvi_b200 import functional as F
    gd = golden("kl")
    for i in range(4):
        lam = float(gd[f"lam_{i}"])
        mu, rho = t(gd[f"mu_{i}"]).requires_grad_(), t(gd[f"rho_{i}"]).requires_grad_()
        kl = F.kl_gaussian(mu, rho, lam, 0)
        (3.0 * kl).backward()
        ref = float(gd[f"kl_{i}"])
        assert abs(kl.item() - ref) <= 1e-5 * max(1.0, abs(ref))
        assert rel_err(mu.grad.cpu().numpy() / 3.0, gd[f"dmu_{i}"]) < 1e-5
        assert rel_err(rho.grad.cpu().numpy() / 3.0, gd[f"drho_{i}"]) < 1e-5


@pytest.mark.parametrize("D", [1, 7, 100, 4096, 32768])
@pytest.mark.parametrize("mode", [0, 1])
def test_kl_vs_oracle(D, mode):
    from whvi_b200 import functional as F
    rng = np.random.default_rng(D + mode)
    mu, rho = rng.standard_normal(D), rng.standard_normal(D) * 2
    mt, rt = t(mu).requires_grad_(), t(rho).requires_grad_()
    kl = F.kl_gaussian(mt, rt, 0.7, mode)
    kl.backward()
    v, dmu, drho = O.kl(mu, rho, 0.7, mode, grads=True)
    assert abs(kl.item() - v) <= 2e-5 * max(1.0, abs(v))
    assert rel_err(mt.grad.cpu().numpy(), dmu) < 1e-5
    assert rel_err(rt.grad.cpu().numpy(), drho) < 1e-4


@pytest.mark.parametrize("S,D", [(1, 4), (3, 100), (64, 128), (7, 8192)])
def test_reparam_vs_oracle(S, D):
    from whvi_b200 import functional as F
    rng = np.random.default_rng(S * D)
    mu, rho, eps, dg = rng.standard_normal(D), rng.standard_normal(D), rng.standard_normal((S, D)), rng.standard_normal((S, D))
    mt, rt = t(mu).requires_grad_(), t(rho).requires_grad_()
    g = F.reparam(mt, rt, t(eps))
    assert rel_err(g.detach().cpu().numpy(), O.reparam(mu, rho, eps)) < 1e-6
    (g * t(dg)).sum().backward()
    dmu, drho = O.reparam_bwd(rho, eps, dg)
    assert rel_err(mt.grad.cpu().numpy(), dmu) < 1e-5
    assert rel_err(rt.grad.cpu().numpy(), drho) < 1e-5


def test_layer_full_size_properties():
    """BASELINE config 4 shard (D = 4096, 2^17 rows): adjointness <W x, dy> = <x, W^T dy>
    ties the forward and the backward kernels together at full size, and a row sample is
    checked against the fp64 oracle."""
    from whvi_b200 import functional as F
    S, B, D = 16, 8192, 4096
    gen = torch.Generator(device=dev()).manual_seed(1)
    x = torch.randn(S, B, D, device=dev(), generator=gen)
    dy = torch.randn(S, B, D, device=dev(), generator=gen)
    g = torch.randn(S, D, device=dev(), generator=gen)
    s1 = torch.randn(D, device=dev(), generator=gen)
    s2 = torch.randn(D, device=dev(), generator=gen)
    y = F.layer_forward_raw(x, g, s1, s2)
    dx, dg, ds1, ds2, _ = F.layer_backward_raw(x, dy, g, s1, s2)
    lhs = (y.double() * dy.double()).sum().item()
    rhs = (x.double() * dx.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * abs(lhs)
    # <y, dy> also equals <s1, ds1> = <s2, ds2> = <g, dg> (each parameter enters linearly)
    for p, dp in ((s1, ds1), (s2, ds2), (g, dg)):
        assert abs((p.double() * dp.double()).sum().item() - lhs) <= 1e-4 * abs(lhs)
    rows = [(0, 0), (3, 4097), (15, 8191)]
    xs = np.stack([x[s, b].cpu().numpy() for s, b in rows]).astype(np.float64)
    for i, (s, b) in enumerate(rows):
        ref = O.layer_fwd(xs[i][None, None], g[s].cpu().numpy().astype(np.float64)[None], s1.cpu().numpy().astype(np.float64),
                          s2.cpu().numpy().astype(np.float64))[0, 0]
        assert rel_err(y[s, b].cpu().numpy(), ref) < TOL


@pytest.mark.parametrize("S,D", [(16, 128), (64, 256), (37, 512), (128, 1024), (300, 256), (1, 128), (128, 2048), (128, 4096), (40, 8192),
                                 (256, 4096), (513, 640)])
def test_reparam_dense_tcgen05_vs_oracle(S, D):
    """Kernel (4): g = mu + eps @ L^T on the tcgen05 tensor cores (3xTF32 split).  Not in the
    reference (parity unpinned): checked against the fp64 oracle formula and torch autograd."""
    from whvi_b200 import functional as F
    rng = np.random.default_rng(S + D)
    mu, eps = rng.standard_normal(D), rng.standard_normal((S, D))
    L = np.tril(rng.standard_normal((D, D))) / np.sqrt(D)
    Lgarbage = L + np.triu(rng.standard_normal((D, D)), 1)  # entries above the diagonal must be ignored
    mt, Lt = t(mu).requires_grad_(), t(Lgarbage).requires_grad_()
    g = F.reparam_dense(mt, Lt, t(eps))
    ref = O.reparam(mu, L, eps, dense=True)
    assert rel_err(g.detach().cpu().numpy(), ref) < 1e-5
    dg = rng.standard_normal((S, D))
    (g * t(dg)).sum().backward()
    rdmu, rdL = O.reparam_dense_bwd(eps, dg)
    assert rel_err(mt.grad.cpu().numpy(), rdmu) < 1e-5
    dL = Lt.grad.cpu().numpy()
    assert rel_err(dL, rdL) < 1e-5
    assert np.all(np.triu(dL, 1) == 0)


@pytest.mark.parametrize("D", [128, 384, 1024, 4096])
def test_kl_dense_vs_oracle(D):
    """Kernel (5), dense form: KL(N(mu, L L^T) || N(0, lambda I)) with the log-determinant and its gradients
    (whvi_kl_dense_f32) against the fp64 oracle formula (tests/test_oracle.py ties that to torch's MVN KL)."""
    from whvi_b200 import functional as F
    rng = np.random.default_rng(D)
    lam = 0.05
    mu = rng.standard_normal(D)
    L = np.tril(rng.standard_normal((D, D))) / np.sqrt(D)
    L[np.arange(D), np.arange(D)] = np.abs(np.diagonal(L)) + 0.05
    Lg = L + np.triu(rng.standard_normal((D, D)), 1)
    mt, Lt = t(mu).requires_grad_(), t(Lg).requires_grad_()
    kl = F.kl_gaussian_dense(mt, Lt, lam)
    v, dmu, dL = O.kl_dense(mu.astype(np.float32).astype(np.float64), L.astype(np.float32).astype(np.float64), lam, grads=True)
    assert abs(kl.item() - v) < 1e-5 * abs(v)
    (3.0 * kl).backward()
    assert rel_err(mt.grad.cpu().numpy(), 3.0 * dmu) < 1e-5
    assert rel_err(Lt.grad.cpu().numpy(), 3.0 * dL) < 1e-5
    with torch.no_grad():
        assert abs(F.kl_gaussian_dense(mt, Lt, lam).item() - v) < 1e-5 * abs(v)   # value-only call (no gradient buffers)


def test_dense_covariance_module():
    """WHVISquarePow2Matrix(covariance="dense"): starts at the reference's diagonal posterior (L = diag(softplus(rho)),
    same init RNG order), so with the same noise its outputs, KL (sigma^2 form) and non-L gradients equal the diagonal
    module's; the L gradient is tril(dg^T eps) of the oracle."""
    from whvi_b200.weights import WHVISquarePow2Matrix
    D, S, B = 256, 5, 9
    torch.manual_seed(3)
    a = WHVISquarePow2Matrix(D, lambda_=0.1, bias=True, kl_mode=1).to(dev())
    torch.manual_seed(3)
    b = WHVISquarePow2Matrix(D, lambda_=0.1, bias=True, covariance="dense").to(dev())
    assert "g_L" in b.state_dict() and "g_rho" not in b.state_dict()
    eps = torch.randn(S, D, device=dev())
    x = torch.randn(B, D, device=dev())
    outs = []
    for m in (a, b):
        m.mc_samples = S
        m._eps_queue.append(eps.clone())
        y = m(x)
        kl = m.kl
        (y.square().sum() + kl).backward()
        outs.append((y.detach().cpu().numpy(), kl.item()))
    assert rel_err(outs[1][0], outs[0][0]) < 1e-5
    assert abs(outs[1][1] - outs[0][1]) < 1e-5 * abs(outs[0][1])
    for n in ("s1", "s2", "g_mu", "bias"):
        assert rel_err(getattr(b, n).grad.cpu().numpy(), getattr(a, n).grad.cpu().numpy()) < 1e-4
    # d/d sigma_i of the diagonal module = the diagonal of dL
    sig_grad = a.g_rho.grad / torch.sigmoid(a.g_rho.detach())
    assert rel_err(torch.diagonal(b.g_L.grad).cpu().numpy(), sig_grad.cpu().numpy()) < 1e-4
    assert np.all(np.triu(b.g_L.grad.cpu().numpy(), 1) == 0)


@pytest.mark.parametrize("S,B,D,bias,shared", [(2, 37, 128, True, False), (3, 9, 512, False, False), (2, 21, 1024, True, False),
                                                (2, 5, 2048, False, False), (2, 5, 4096, True, False), (3, 6, 1024, False, True)])
def test_loss_layer_vs_oracle(S, B, D, bias, shared):
    """whvi_layer_loss_f32: forward + residual + backward for a unit coefficient in one pass."""
    from whvi_b200 import functional as F
    x, g, s1, s2, _, bvec = make_case(S, B, D, 7 * D + B, shared)
    rng = np.random.default_rng(D + 1)
    target = rng.standard_normal((B, D))
    bb = bvec if bias else None
    sq, dx, dg, ds1, ds2, db = F.layer_loss_raw(t(x), t(g), t(s1), t(s2), None if bb is None else t(bb), t(target),
                                                want_dx=True, relu_in=True)
    y = O.layer_fwd(x, g, s1, s2, bb)
    r = y - target[None]
    assert abs(sq.item() - float((r ** 2).sum())) < 1e-4 * float((r ** 2).sum())
    rdx, rdg, rds1, rds2, rdb = O.layer_bwd(x, r, g, s1, s2, want_dbias=True)
    x32 = x.astype(np.float32)
    dxn = dx.cpu().numpy()
    if shared:
        mask = (x32 > 0)[None]
        ref_dx = (O.layer_bwd(np.broadcast_to(x, (S, B, D)).copy(), r, g, s1, s2)[0] * mask)
        assert rel_err(dxn, ref_dx) < TOL
    else:
        assert rel_err(dxn, rdx * (x32 > 0)) < TOL
    assert rel_err(dg.cpu().numpy(), rdg) < TOL
    assert rel_err(ds1.cpu().numpy(), rds1) < TOL
    assert rel_err(ds2.cpu().numpy(), rds2) < TOL
    if bias:
        assert rel_err(db.cpu().numpy(), rdb) < TOL


def test_deferred_scale_rides_on_the_graph():
    """The fused loss layer hands its producer a dx for a UNIT loss coefficient; the coefficient travels in a
    DeferredScale shared by the two autograd nodes, so a hook that copies the gradient tensor in between (the
    failure mode of a data_ptr-keyed side table) changes nothing, and without a holder dx is scaled in place."""
    from whvi_b200 import functional as F
    S, B, D = 2, 4, 128
    x, g, s1, s2, _, _ = make_case(S, B, D, 5)
    x2, g2, s1b, s2b, _, _ = make_case(S, B, D, 6)
    tgt = t(np.zeros((B, D)))

    def run(shared_holder, hook):
        xt, s1t = t(x).requires_grad_(), t(s1).requires_grad_()
        holder = F.DeferredScale() if shared_holder else None
        h = F.whvi_layer(xt, t(g), s1t, t(s2), None, True, False, holder)
        if hook:
            h.register_hook(lambda gr: gr.clone())     # the producer sees a COPY of the loss layer's dx
        sq = F.whvi_layer_loss(h, t(g2), t(s1b), t(s2b), None, tgt, relu_in=True, dx_scale_to=holder)
        (0.37 * sq).backward()
        assert holder is None or holder.value is None  # consumed
        return xt.grad.cpu().numpy(), s1t.grad.cpu().numpy()

    base = run(False, False)
    for shared, hook in ((True, False), (True, True), (False, True)):
        got = run(shared, hook)
        assert rel_err(got[0], base[0]) < 1e-6 and rel_err(got[1], base[1]) < 1e-6


@pytest.mark.parametrize("S,B,D", [(2, 37, 16), (3, 9, 128), (2, 7, 1024), (2, 5, 4096), (2, 3, 8192), (2, 2, 32768)])
@pytest.mark.parametrize("shared", [False, True])
def test_forward_bf16_io(S, B, D, shared):
    """whvi_layer_fwd_bf16: bf16 activations in HBM, fp32 parameters and arithmetic.  Stated tolerance: equal,
    bit for bit, to the fp32 kernel's output on the same inputs rounded once to bf16 (so within 2^-9 relative of it)."""
    from whvi_b200 import functional as F
    x, g, s1, s2, _, bias = make_case(S, B, D, 3 * D + S, shared)
    xb = t(x).to(torch.bfloat16)
    for relu in (False, True):
        y = F.layer_forward_bf16(xb, t(g), t(s1), t(s2), t(bias), relu_out=relu)
        ref = F.layer_forward_raw(xb.float(), t(g), t(s1), t(s2), t(bias), relu_out=relu)
        assert y.dtype == torch.bfloat16 and torch.equal(y, ref.to(torch.bfloat16))
    y_ref = O.layer_fwd(xb.float().cpu().numpy().astype(np.float64), g, s1, s2, bias)
    assert rel_err(F.layer_forward_bf16(xb, t(g), t(s1), t(s2), t(bias)).float().cpu().numpy(), y_ref) < 2.0 ** -8
    if shared:   # hoisted first transform (config 5's evaluation path) with bf16 t2
        t2 = F.layer_forward_raw  # noqa: F841 (documented pairing: t2 = H(s2 * x) from the fp32 FWHT, stored as bf16)
        from whvi_b200 import fwht_
        t2b = fwht_((xb.float() * t(s2))).to(torch.bfloat16)
        y2 = F.layer_forward_bf16(t2b, t(g), t(s1), t(s2), t(bias), from_t2=True)
        ref2 = F.layer_forward_raw(t2b.float(), t(g), t(s1), t(s2), t(bias), from_t2=True)
        assert torch.equal(y2, ref2.to(torch.bfloat16))


def test_shared_input_transform_is_hoisted_in_the_autograd_path():
    """First layer of a network (one (B, D) input block for all samples): above HOIST_MIN_ELEMENTS the autograd function
    computes t2 = H(s2 x) once (whvi_fwht_scaled_f32) and runs the forward from it; same outputs and gradients as the
    plain kernels."""
    from whvi_b200 import functional as F
    S, B, D = 4, 1024, 4096
    assert S * B * D >= F.HOIST_MIN_ELEMENTS
    gen = torch.Generator(device=dev()).manual_seed(3)
    x = torch.randn(B, D, device=dev(), generator=gen, requires_grad=True)
    g = torch.randn(S, D, device=dev(), generator=gen, requires_grad=True)
    s1 = torch.randn(D, device=dev(), generator=gen, requires_grad=True)
    s2 = torch.randn(D, device=dev(), generator=gen, requires_grad=True)
    bias = torch.randn(D, device=dev(), generator=gen)
    dy = torch.randn(S, B, D, device=dev(), generator=gen)
    before = dict(F.LAUNCH_COUNTS)
    y = F.whvi_layer(x, g, s1, s2, bias, relu_out=True)
    assert F.LAUNCH_COUNTS.get("whvi_fwht_scaled_f32", 0) == before.get("whvi_fwht_scaled_f32", 0) + 1
    y_ref = F.layer_forward_raw(x.detach(), g.detach(), s1.detach(), s2.detach(), bias, relu_out=True)
    assert rel_err(y.detach().cpu().numpy(), y_ref.cpu().numpy()) < 1e-5
    (y * dy).sum().backward()
    # a folded ReLU's mask is the CONSUMER's job (relu_in), so the function's backward sees dy as it is
    dx, dg, ds1, ds2, _ = F.layer_backward_raw(x.detach(), dy, g.detach(), s1.detach(), s2.detach(), want_dx=True)
    assert rel_err(x.grad.cpu().numpy(), dx.sum(0).cpu().numpy()) < TOL
    assert rel_err(g.grad.cpu().numpy(), dg.cpu().numpy()) < TOL
    assert rel_err(s1.grad.cpu().numpy(), ds1.cpu().numpy()) < TOL
    assert rel_err(s2.grad.cpu().numpy(), ds2.cpu().numpy()) < TOL
