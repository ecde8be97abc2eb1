"""GPU: the drop-in modules against (a) goldens produced by the reference itself, in
semantics="reference" with the captured noise injected, and (b) the fp64 oracle in the
default PAPER semantics."""
import numpy as np
import pytest
import torch

import whvi_b200 as W
from conftest import rel_err
from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4


def dev():
    return torch.device("cuda:0")


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev())


LAYER_CASES = ["square16", "square16_bias", "stacked_3_16", "stacked_13_32_bias", "column_16_1", "column_1_8_bias"]


def load_layer(g, name, semantics):
    n_in, n_out, bias, B = (int(v) for v in g[f"{name}.shape"])
    layer = W.WHVILinear(n_in, n_out, lambda_=2.0, bias=bool(bias), semantics=semantics)
    sd = {k: torch.from_numpy(g[f"{name}.param.{k}"]) for k in layer.state_dict().keys()}
    layer.load_state_dict(sd)
    return layer.to(dev()), n_in, n_out, B


@pytest.mark.parametrize("name", LAYER_CASES)
def test_reference_semantics_reproduce_the_reference(golden, name):
    """Same parameters + same eps => same outputs, KL and gradients as the reference code."""
    g = golden("layers")
    layer, n_in, n_out, B = load_layer(g, name, "reference")
    blocks = layer.square_blocks()
    assert int(g[f"{name}.n_eps"]) == len(blocks)
    for i, b in enumerate(blocks):
        b.inject_eps(t(g[f"{name}.eps{i}"]))
    x = t(g[f"{name}.x"]).requires_grad_()
    y = layer(x)
    assert y.shape == (B, n_out)
    assert rel_err(y.detach().cpu().numpy(), g[f"{name}.y"]) < TOL
    (y * t(g[f"{name}.dy"])).sum().backward()
    assert rel_err(x.grad.cpu().numpy(), g[f"{name}.dx"]) < TOL
    for pname, p in layer.named_parameters():
        ref = g[f"{name}.grad.{pname}"]
        got = np.zeros_like(ref) if p.grad is None else p.grad.cpu().numpy()
        if np.max(np.abs(ref)) == 0:
            assert np.max(np.abs(got)) < 1e-6, pname
        else:
            assert rel_err(got, ref) < TOL, pname
    assert abs(float(layer.kl.detach()) - float(g[f"{name}.kl"])) < 1e-4 * max(1.0, abs(float(g[f"{name}.kl"])))


def test_reference_semantics_toy_model(golden):
    """README toy model, one training batch: loss, KL, MNLL and every gradient."""
    g = golden("toy")
    model = W.WHVIRegression([W.WHVILinear(3, 16, lambda_=2.0, semantics="reference"), torch.nn.ReLU(),
                              W.WHVILinear(16, 1, semantics="reference")], train_samples=3)
    assert list(model.state_dict().keys()) == [str(k) for k in g["state_dict_keys"]]
    model.load_state_dict({k: torch.from_numpy(g[f"param.{k}"]) for k in model.state_dict().keys()})
    model = model.to(dev()).train()
    blocks = [b for layer in model._whvi_layers() for b in layer.square_blocks()]
    S, n_blocks = 3, len(blocks)
    assert int(g["n_eps"]) == S * n_blocks  # reference order: sample-major, then layer, then block
    for i, b in enumerate(blocks):
        b.inject_eps(torch.stack([t(g[f"eps{s * n_blocks + i}"]) for s in range(S)]))
    loss = model.loss(t(g["x"]), t(g["y"]), n=150)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    assert abs(float(model.current_kl) - float(g["kl"])) < 1e-4 * abs(float(g["kl"]))
    assert abs(float(model.current_mnll) - float(g["mnll"])) < 1e-4 * abs(float(g["mnll"]))
    # gradients that are exactly zero in exact arithmetic show up as ~1e-5 rounding noise in
    # the reference (its dense W carries 1e-8 off-diagonal noise), so the scale of the
    # comparison is the largest gradient of the model, not of each tensor
    scale = max(float(np.max(np.abs(g[f"grad.{name}"]))) for name, _ in model.named_parameters())
    for name, p in model.named_parameters():
        diff = float(np.max(np.abs(p.grad.cpu().numpy().astype(np.float64) - g[f"grad.{name}"])))
        assert diff < TOL * scale, (name, diff, scale)
    # eval-mode predictions, 4 samples
    model.eval_samples = 4
    model.eval()
    S = 4
    for i, b in enumerate(blocks):
        b.inject_eps(torch.stack([t(g[f"eval_eps{s * n_blocks + i}"]) for s in range(S)]))
    with torch.no_grad():
        pred = model(t(g["eval_x"]))
    assert pred.shape == g["eval_pred"].shape
    assert rel_err(pred.cpu().numpy(), g["eval_pred"]) < TOL
    # WHVIRegression.eval_model: RMSE of the MC mean and test MNLL as the reference returns them
    assert int(g["evalm_n_eps"]) == S * n_blocks
    for i, b in enumerate(blocks):
        b.inject_eps(torch.stack([t(g[f"evalm_eps{s * n_blocks + i}"]) for s in range(S)]))
    rmse, mnll = model.eval_model(t(g["evalm_x"]), t(g["evalm_y"]))
    assert isinstance(rmse, float) and isinstance(mnll, float)
    assert abs(rmse - float(g["evalm_rmse"])) < TOL * float(g["evalm_rmse"])
    assert abs(mnll - float(g["evalm_mnll"])) < TOL * abs(float(g["evalm_mnll"]))


def oracle_layer(layer, x, eps_per_block):
    """PAPER-semantics expected output of a WHVILinear from the fp64 oracle pieces."""
    w = layer.weight_submodule
    f64 = lambda p: p.detach().cpu().numpy().astype(np.float64)
    if isinstance(w, W.WHVISquarePow2Matrix):
        g = O.reparam(f64(w.g_mu), f64(w.g_rho), eps_per_block[0])
        y = O.layer_fwd(x, g, f64(w.s1), f64(w.s2), None if w.bias is None else f64(w.bias))
        return y
    if isinstance(w, W.WHVIStackedMatrix):
        pad = [(0, 0)] * (x.ndim - 1) + [(0, w.D_in - w.n_in)]
        xp = np.pad(x, pad)
        ys = []
        for blk, eps in zip(w.weight_matrices, eps_per_block):
            g = O.reparam(f64(blk.g_mu), f64(blk.g_rho), eps)
            ys.append(O.layer_fwd(xp, g, f64(blk.s1), f64(blk.s2)))
        y = np.concatenate(ys, axis=-1)
        if w.bias is not None:
            y = y + f64(w.bias)
        return y[..., :w.n_out]
    sub = w.weight_submodule
    g = O.reparam(f64(sub.g_mu), f64(sub.g_rho), eps_per_block[0])
    S = g.shape[0]
    wv = np.stack([O.column_weight_paper(g[s], f64(sub.s1), f64(sub.s2), w.D) for s in range(S)])  # (S, D)
    xs = np.broadcast_to(x, (S,) + x.shape[-2:]) if x.ndim == 2 else x
    y = np.einsum("sbi,si->sb", xs, wv)[..., None] if w.transposed else xs * wv[:, None, :]
    if w.bias is not None:
        y = y + f64(w.bias)
    return y


@pytest.mark.parametrize("n_in,n_out,bias", [(16, 16, False), (128, 128, True), (3, 16, False), (13, 128, True),
                                             (128, 1, False), (1, 8, True), (20, 50, False), (1024, 1024, False),
                                             (2, 2, True), (2, 5, False), (1, 1, False), (2, 1, False)])
@pytest.mark.parametrize("shared", [True, False])
def test_paper_semantics_vs_oracle(n_in, n_out, bias, shared):
    torch.manual_seed(n_in * 1000 + n_out)
    S, B = 3, 6
    layer = W.WHVILinear(n_in, n_out, lambda_=3.0, bias=bias).to(dev())
    with torch.no_grad():
        for name, p in layer.named_parameters():
            p.copy_(torch.randn_like(p) * (0.5 if name.endswith("g_rho") else 1.0))
    rng = np.random.default_rng(n_in + n_out)
    x = rng.standard_normal((B, n_in) if shared else (S, B, n_in))
    blocks = layer.square_blocks()
    eps = [rng.standard_normal((S, b.D)) for b in blocks]
    for b, e in zip(blocks, eps):
        b.inject_eps(t(e))
    layer.mc_samples = S
    y = layer(t(x))
    layer.mc_samples = None
    assert y.shape == (S, B, n_out)
    ref = oracle_layer(layer, x, eps)
    assert rel_err(y.detach().cpu().numpy(), ref) < TOL


@pytest.mark.parametrize("n_in,n_out,bias", [(3, 16, False), (13, 128, True), (20, 50, True), (16, 40, False), (128, 300, True), (1024, 2048, False)])
@pytest.mark.parametrize("shared", [True, False])
@pytest.mark.parametrize("relu", [False, True])
def test_stacked_one_launch_equals_block_by_block(n_in, n_out, bias, shared, relu):
    """WHVIStackedMatrix as one grouped launch per direction (whvi_stacked_fwd/bwd_f32, src/weights.py:179-208) against the
    same module run block after block like the reference: outputs, KL, input and every parameter gradient."""
    import copy
    torch.manual_seed(n_in + n_out)
    S, B = 3, 7
    a = W.WHVILinear(n_in, n_out, lambda_=3.0, bias=bias).to(dev())
    with torch.no_grad():
        for name, p in a.named_parameters():
            p.copy_(torch.randn_like(p) * (0.5 if name.endswith("g_rho") else 1.0))
    b = copy.deepcopy(a)
    b.weight_submodule.one_launch = False
    assert a.weight_submodule.grouped and not b.weight_submodule.grouped
    x = torch.randn((B, n_in) if shared else (S, B, n_in), device=dev())
    if relu:
        x = x.relu()     # the layer as the consumer of a fused ReLU: dx is masked by (x > 0)
    G, D = a.weight_submodule.stack, a.weight_submodule.D_in
    eps = torch.randn(G, S, D, device=dev())
    dy = torch.randn(S, B, n_out, device=dev())
    res = []
    for layer, fused in ((a, True), (b, False)):
        for k, blk in enumerate(layer.square_blocks()):
            blk.inject_eps(eps[k])
        layer.mc_samples = S
        xi = x.clone().requires_grad_()
        if fused:
            y = layer.weight_submodule.forward(xi, relu_out=relu, relu_in=relu)
        else:
            y = layer.weight_submodule.forward(xi)
            y = y.relu() if relu else y
        layer.mc_samples = None
        kl = layer.kl
        # a fused ReLU's mask is applied by the CONSUMER's backward (relu_in), so the upstream gradient arrives masked
        w = dy * (y.detach() > 0) if relu else dy
        ((y * w).sum() + 0.7 * kl).backward()
        dx = xi.grad * (x > 0) if (relu and not fused) else xi.grad
        res.append((y.detach(), kl.detach(), dx, {n: p.grad for n, p in layer.named_parameters()}))
    (y1, kl1, dx1, g1), (y0, kl0, dx0, g0) = res
    assert y1.shape == (S, B, n_out)
    assert rel_err(y1.cpu().numpy(), y0.cpu().numpy()) < 1e-5
    assert abs(kl1.item() - kl0.item()) < 1e-5 * abs(kl0.item())
    assert rel_err(dx1.cpu().numpy(), dx0.cpu().numpy()) < TOL
    for n in g0:
        assert rel_err(g1[n].cpu().numpy(), g0[n].cpu().numpy()) < TOL, n
    # the parameters were packed evenly spaced on the first call: the kernels read them where they lie
    from whvi_b200 import functional as F
    assert F.uniform_stride([blk.s1 for blk in a.square_blocks()]) is not None


@pytest.mark.parametrize("n,transposed,bias", [(128, True, True), (100, True, False), (1, True, True), (8, False, True), (50, False, False),
                                               (2, True, False), (3000, True, True), (3000, False, True)])
@pytest.mark.parametrize("shared", [True, False])
def test_column_one_call_equals_op_chain(n, transposed, bias, shared):
    """WHVIColumnMatrix in one C-ABI call per direction (whvi_column_fwd/bwd_f32, src/weights.py:231-251) against round 1's op
    chain (reparam -> FWHT of g -> torch products), which test_paper_semantics_vs_oracle pins to the fp64 oracle."""
    import copy
    torch.manual_seed(n)
    S, B = 3, 9
    a = (W.WHVILinear(n, 1, lambda_=3.0, bias=bias) if transposed else W.WHVILinear(1, n, lambda_=3.0, bias=bias)).to(dev())
    if n == 1:
        a = W.WHVILinear(1, 1, lambda_=3.0, bias=bias).to(dev())
    with torch.no_grad():
        for name, p in a.named_parameters():
            p.copy_(torch.randn_like(p) * (0.5 if name.endswith("g_rho") else 1.0))
    b = copy.deepcopy(a)
    b.weight_submodule.one_launch = False
    assert a.weight_submodule.fused and not b.weight_submodule.fused
    width = a.weight_submodule.D if a.weight_submodule.transposed else 1
    x = torch.randn((B, width) if shared else (S, B, width), device=dev())
    relu = a.weight_submodule.transposed   # as the consumer of a folded ReLU
    if relu:
        x = x.relu()
    D = a.weight_submodule.D_adjusted
    eps = torch.randn(S, D, device=dev())
    res = []
    for layer, fused in ((a, True), (b, False)):
        layer.square_blocks()[0].inject_eps(eps)
        layer.mc_samples = S
        xi = x.clone().requires_grad_()
        y = layer.weight_submodule.forward(xi, relu_in=relu) if fused else layer.weight_submodule.forward(xi)
        layer.mc_samples = None
        dy = torch.cos(torch.arange(y.numel(), device=dev(), dtype=torch.float32)).reshape(y.shape)
        (y * dy).sum().backward()
        dx = xi.grad * (x > 0) if (relu and not fused) else xi.grad
        res.append((y.detach(), dx, {k: p.grad for k, p in layer.named_parameters()}))
    (y1, dx1, g1), (y0, dx0, g0) = res
    assert y1.shape == y0.shape == (S, B, 1 if a.weight_submodule.transposed else n)
    assert rel_err(y1.cpu().numpy(), y0.cpu().numpy()) < 1e-5
    assert rel_err(dx1.cpu().numpy(), dx0.cpu().numpy()) < TOL
    for k in g0:
        ref = g0[k].cpu().numpy()
        if np.abs(ref).max() == 0:
            assert np.abs(g1[k].cpu().numpy()).max() == 0, k
        else:
            assert rel_err(g1[k].cpu().numpy(), ref) < TOL, k


@pytest.mark.parametrize("n_in,n_out", [(13, 128), (20, 50), (100, 1), (1, 40)])
def test_non_square_layers_standalone_two_dimensional(n_in, n_out):
    """Outside a WHVINetwork (mc_samples unset) the one-call Stacked / Column paths behave like the reference's modules:
    (batch, n_in) -> (batch, n_out), one noise draw, autograd through the input; also under no_grad."""
    import copy
    torch.manual_seed(n_in * 7 + n_out)
    a = W.WHVILinear(n_in, n_out, lambda_=2.0, bias=True).to(dev())
    b = copy.deepcopy(a)
    b.weight_submodule.one_launch = False
    x = torch.randn(11, n_in, device=dev())
    outs = []
    for layer in (a, b):
        for blk in layer.square_blocks():
            blk.inject_eps(torch.full((1, blk.D), 0.5, device=dev()))
        xi = x.clone().requires_grad_()
        y = layer(xi)
        assert y.shape == (11, n_out)
        y.sin().sum().backward()
        outs.append((y.detach().cpu().numpy(), xi.grad.cpu().numpy()))
        with torch.no_grad():
            for blk in layer.square_blocks():
                blk.inject_eps(torch.full((1, blk.D), 0.5, device=dev()))
            assert rel_err(layer(x).cpu().numpy(), outs[-1][0]) < 1e-6
    assert rel_err(outs[0][0], outs[1][0]) < 1e-5
    assert rel_err(outs[0][1], outs[1][1]) < TOL


def test_stacked_parameters_are_packed_once_and_survive_flat_params():
    """The grouped launch reads the blocks' parameters where they lie: the module packs them evenly spaced on first use
    (same values, same Parameter objects, same state_dict), and FlatParams' own re-homing keeps them evenly spaced."""
    from whvi_b200 import functional as F
    torch.manual_seed(1)
    layer = W.WHVILinear(13, 128, lambda_=2.0).to(dev())
    before = {k: v.clone() for k, v in layer.state_dict().items()}
    params = list(layer.parameters())
    blocks = layer.square_blocks()
    with torch.no_grad():   # scatter them: the first forward has to pack
        for i, b in enumerate(blocks):
            b.s1.data = torch.cat([torch.zeros(4 * i, device=dev()), b.s1.data])[4 * i:]
    assert F.uniform_stride([b.s1 for b in blocks]) is None
    layer.mc_samples = 2
    y0 = layer(torch.ones(4, 13, device=dev()))
    assert F.uniform_stride([b.s1 for b in blocks]) is not None   # packed by the module, or already evenly spaced by the allocator
    assert all(p is q for p, q in zip(params, layer.parameters()))
    assert all(torch.equal(v, layer.state_dict()[k]) for k, v in before.items())
    flat = W.FlatParams(layer.parameters())
    assert F.uniform_stride([b.s1 for b in blocks]) is not None and F.uniform_stride([b.g_rho for b in blocks]) is not None
    for b in blocks:
        b.inject_eps(torch.zeros(2, b.D, device=dev()))
    y1 = layer(torch.ones(4, 13, device=dev()))
    y1.sum().backward()
    assert flat.attached() and float(flat.grad.abs().sum()) > 0
    layer.mc_samples = None
    assert y0.shape == y1.shape == (2, 4, 128)


@pytest.mark.parametrize("n_in,n_out", [(16, 16), (13, 128), (100, 1), (1, 40)])
def test_empty_batch(n_in, n_out):
    """Edge case the reference's ops handle implicitly: a batch of zero rows gives an empty output and zero gradients."""
    layer = W.WHVILinear(n_in, n_out, lambda_=2.0, bias=True).to(dev())
    layer.mc_samples = 3
    x = torch.zeros(0, n_in, device=dev(), requires_grad=True)
    y = layer(x)
    layer.mc_samples = None
    assert y.shape == (3, 0, n_out)
    (y.sum() + layer.kl).backward()
    for name, p in layer.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
        if not name.endswith(("g_mu", "g_rho")):   # only the KL term reaches mu / rho
            assert float(p.grad.abs().sum()) == 0.0, name


def test_column_layers_in_a_fused_network():
    """[Column 1 -> 64 (produces a folded ReLU), Square 64, Column 64 -> 1 (consumes one)]: the fused network against every
    module run on its own."""
    import copy
    torch.manual_seed(9)
    S, B = 3, 21
    fused = W.WHVIRegression([W.WHVILinear(1, 64, lambda_=3.0, bias=True), torch.nn.ReLU(), W.WHVILinear(64, 64, lambda_=3.0),
                              torch.nn.ReLU(), W.WHVILinear(64, 1, lambda_=3.0, bias=True)], train_samples=S).to(dev()).train()
    with torch.no_grad():
        for name, p in fused.named_parameters():
            if name.endswith(("s1", "s2", "g_mu")):
                p.normal_()
    plain = copy.deepcopy(fused)
    plain.fuse = False
    for m in plain._whvi_layers():
        if isinstance(m.weight_submodule, W.WHVIColumnMatrix):
            m.weight_submodule.one_launch = False
    x, y = torch.randn(B, 1, device=dev()), torch.randn(B, 1, device=dev())
    eps = [torch.randn(S, b.D, device=dev()) for layer in fused._whvi_layers() for b in layer.square_blocks()]
    out = []
    for model in (fused, plain):
        for b, e in zip([b for layer in model._whvi_layers() for b in layer.square_blocks()], eps):
            b.inject_eps(e)
        loss = model.loss(x, y, n=500)
        loss.backward()
        out.append((loss.item(), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}))
    assert abs(out[0][0] - out[1][0]) < 1e-5 * abs(out[1][0])
    assert out[0][1].keys() == out[1][1].keys()
    scale = max(float(v.abs().max()) for v in out[1][1].values())
    for n, v in out[1][1].items():
        assert float((out[0][1][n] - v).abs().max()) < TOL * scale, n


def test_stacked_network_fused_equals_unfused():
    """BASELINE config 3's shape (13 -> 128 -> 128 -> 1 with ReLUs): the grouped Stacked launch with the ReLU folded into it
    and into the consumer's backward gives the same loss and gradients as every module run on its own."""
    import copy
    torch.manual_seed(5)
    S, B = 4, 33
    fused = W.WHVIRegression([W.WHVILinear(13, 128, lambda_=3.0, bias=True), torch.nn.ReLU(), W.WHVILinear(128, 128, lambda_=3.0),
                              torch.nn.ReLU(), W.WHVILinear(128, 1, lambda_=3.0)], train_samples=S).to(dev()).train()
    plain = copy.deepcopy(fused)
    plain.fuse = False
    for m in plain._whvi_layers():
        if isinstance(m.weight_submodule, (W.WHVIStackedMatrix, W.WHVIColumnMatrix)):
            m.weight_submodule.one_launch = False
    x, y = torch.randn(B, 13, device=dev()), torch.randn(B, 1, device=dev())
    blocks = [b for layer in fused._whvi_layers() for b in layer.square_blocks()]
    eps = [torch.randn(S, b.D, device=dev()) for b in blocks]
    out = []
    for model in (fused, plain):
        for b, e in zip([b for layer in model._whvi_layers() for b in layer.square_blocks()], eps):
            b.inject_eps(e)
        loss = model.loss(x, y, n=1000)
        loss.backward()
        out.append((loss.item(), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}))
    assert abs(out[0][0] - out[1][0]) < 1e-5 * abs(out[1][0])
    assert out[0][1].keys() == out[1][1].keys()
    scale = max(float(v.abs().max()) for v in out[1][1].values())
    for n, v in out[1][1].items():
        assert float((out[0][1][n] - v).abs().max()) < TOL * scale, n


def test_standalone_call_is_two_dimensional_and_dense_sample_agrees():
    torch.manual_seed(3)
    layer = W.WHVISquarePow2Matrix(64, lambda_=1.0).to(dev())
    with torch.no_grad():
        for p in layer.parameters():
            p.copy_(torch.randn_like(p))
    x = torch.randn(5, 64, device=dev())
    eps = torch.randn(1, 64, device=dev())
    layer.inject_eps(eps)
    y = layer(x)
    assert y.shape == (5, 64)
    layer.inject_eps(eps)
    Wd = layer.sample()  # dense (D, D), same eps
    assert rel_err((x @ Wd.T).cpu().detach().numpy(), y.cpu().detach().numpy()) < 1e-4


def test_network_shapes_like_reference_test():
    # test/networks.py:11-23
    for k in (1, 2, 7, 20):
        net = W.WHVIRegression([torch.nn.Linear(1, 8), torch.nn.ReLU(), W.WHVILinear(8, 8), torch.nn.ReLU(),
                                torch.nn.Linear(8, k)], train_samples=5, eval_samples=6).to(dev())
        net.train()
        assert net(torch.randn(50, 1, device=dev())).size() == (50, k, 5)
        net.eval()
        assert net(torch.randn(50, 1, device=dev())).size() == (50, k, 6)


def test_rng_mode_reference_draw_order():
    model = W.WHVIRegression([W.WHVILinear(3, 16), torch.nn.ReLU(), W.WHVILinear(16, 1)], train_samples=2,
                             rng_mode="reference").to(dev()).train()
    blocks = [b for layer in model._whvi_layers() for b in layer.square_blocks()]
    torch.manual_seed(123)
    expected = [[torch.randn(b.D, device=dev()) for b in blocks] for _ in range(2)]  # sample-major loop
    torch.manual_seed(123)
    model._predraw_reference_order(2)
    for i, b in enumerate(blocks):
        got = b._eps_queue.pop(0)
        assert torch.equal(got, torch.stack([expected[s][i] for s in range(2)]))


def test_train_and_eval_model_run_and_learn():
    torch.manual_seed(0)
    x = torch.randn(200, 3, device=dev())
    y = torch.reshape(x[:, 0] + x[:, 1] ** 2 - 0.3 * x[:, 2] ** 3, (-1, 1))
    ds = torch.utils.data.TensorDataset(x[:150], y[:150])
    loader = torch.utils.data.DataLoader(ds, batch_size=64)
    model = W.WHVIRegression([W.WHVILinear(3, 16, lambda_=2.0), torch.nn.ReLU(), W.WHVILinear(16, 1)]).to(dev())
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: (1 + 0.0005 * s) ** (-0.3))
    model.train()
    first = float(model.loss(x[:64], y[:64], n=150))
    model.train_model(loader, opt, sched, epochs1=30, epochs2=30, pbar_update_period=1000)
    err, mnll = model.eval_model(x[150:], y[150:])
    assert np.isfinite(err) and np.isfinite(mnll)
    model.train()
    last = float(model.loss(x[:64], y[:64], n=150))
    assert last < first


@pytest.mark.parametrize("D,bias", [(64, False), (128, True), (1024, False), (2048, True), (4096, False)])
def test_fused_relu_and_mnll_match_unfused(D, bias):
    """fuse=True (ReLU and Gaussian MNLL folded into the layer kernels) must give the same
    loss and gradients as fuse=False (every module its own op, like the reference)."""
    S, B = 3, 37
    torch.manual_seed(D)

    def make(fuse):
        torch.manual_seed(1)
        m = W.WHVIRegression([W.WHVILinear(D, D, lambda_=2.0, bias=bias), torch.nn.ReLU(),
                              W.WHVILinear(D, D, lambda_=2.0, bias=bias), torch.nn.ReLU(),
                              W.WHVILinear(D, D, lambda_=2.0, bias=bias)], train_samples=S, sigma=0.8, fuse=fuse)
        with torch.no_grad():
            for name, p in m.named_parameters():
                if name.endswith(("s1", "s2")):
                    p.mul_(100.0 / D ** 0.5)   # O(1)-gain layers so that activations stay O(1)
                if name.endswith("g_mu"):
                    p.copy_(torch.randn_like(p))
        return m.to(dev()).train()

    fused, plain = make(True), make(False)
    x = torch.randn(B, D, device=dev())
    y = torch.randn(B, D, device=dev())
    eps = [torch.randn(S, D, device=dev()) for _ in range(3)]
    losses = []
    for model in (fused, plain):
        for layer, e in zip(model._whvi_layers(), eps):
            layer.square_blocks()[0].inject_eps(e)
        loss = model.loss(x, y, n=500)
        loss.backward()
        losses.append(loss.item())
    assert abs(losses[0] - losses[1]) < 1e-4 * abs(losses[1])
    scale = max(float(p.grad.abs().max()) for p in plain.parameters())
    for (name, a), (_, b) in zip(fused.named_parameters(), plain.named_parameters()):
        assert float((a.grad - b.grad).abs().max()) < TOL * scale, name
    # predictions (forward()) are unaffected by the loss-side fusion
    for model in (fused, plain):
        for layer, e in zip(model._whvi_layers(), eps):
            layer.square_blocks()[0].inject_eps(e)
    with torch.no_grad():
        pf, pp = fused(x), plain(x)
    assert pf.shape == (B, D, S)
    assert rel_err(pf.cpu().numpy(), pp.cpu().numpy()) < TOL


def _toy_model(seed, sigma_noise_off=True):
    import whvi_b200 as W
    torch.manual_seed(seed)
    model = W.WHVIRegression([W.WHVILinear(3, 16, lambda_=2.0), torch.nn.ReLU(), W.WHVILinear(16, 16, lambda_=2.0),
                              torch.nn.ReLU(), W.WHVILinear(16, 1)], train_samples=4).cuda()
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith(("s1", "s2")):
                p.normal_()          # O(1) scales so that every layer actually learns
            if name.endswith("g_mu"):
                p.normal_()
            if sigma_noise_off and name.endswith("g_rho"):
                p.fill_(-30.0)       # softplus(-30) ~ 1e-13: the MC noise vanishes, runs are comparable
    return model


def test_cuda_graph_step_matches_eager():
    """a captured step (forward + ELBO + backward + Adam) replays to the same
    parameters as the eager loop; the capture's warm-up steps leave no trace."""
    import copy
    from whvi_b200.graphs import GraphedTrainStep
    g = torch.Generator(device="cuda").manual_seed(5)
    xs = [torch.randn(32, 3, device="cuda", generator=g) for _ in range(6)]
    ys = [x[:, :1] ** 2 - x[:, 1:2] for x in xs]
    eager = _toy_model(3)
    graphed = copy.deepcopy(eager)
    opt_e = torch.optim.Adam(eager.parameters(), lr=torch.tensor(1e-2, device="cuda"), capturable=True)
    opt_g = torch.optim.Adam(graphed.parameters(), lr=torch.tensor(1e-2, device="cuda"), capturable=True)
    eager.train(), graphed.train()
    for m, o in ((eager, opt_e), (graphed, opt_g)):   # an eager step on the default stream first: the
        m.loss(xs[5], ys[5], n=150).backward()        # capture must cope with its leftover autograd state
        o.step()
        o.zero_grad(set_to_none=True)
    step = GraphedTrainStep(graphed, opt_g, xs[0], ys[0], n=150)
    for p, q in zip(eager.parameters(), graphed.parameters()):
        assert torch.equal(p, q), "warm-up steps of the capture must not train the model"
    losses_e, losses_g = [], []
    for x, y in zip(xs, ys):
        loss = eager.loss(x, y, n=150)
        loss.backward()
        opt_e.step()
        opt_e.zero_grad(set_to_none=True)
        losses_e.append(float(loss))
        losses_g.append(float(step(x, y)))
    assert np.allclose(losses_e, losses_g, rtol=1e-4), (losses_e, losses_g)
    for (name, p), q in zip(eager.named_parameters(), graphed.parameters()):
        assert rel_err(q.detach().cpu().numpy(), p.detach().cpu().numpy()) < 1e-4, name
    with pytest.raises(RuntimeError, match="capturable"):
        GraphedTrainStep(graphed, torch.optim.Adam(graphed.parameters(), lr=1e-3), xs[0], ys[0], n=150)


def test_train_model_cuda_graph():
    """train_model(cuda_graph=True) keeps the reference's loop semantics (two phases, scheduler
    stepping a tensor lr, ragged last batch) and actually trains."""
    model = _toy_model(11, sigma_noise_off=False)
    torch.manual_seed(0)
    x = torch.randn(150, 3, device="cuda")
    y = x[:, :1] + x[:, 1:2] ** 2
    ds = torch.utils.data.TensorDataset(x, y)
    loader = torch.utils.data.DataLoader(ds, batch_size=64)   # 64, 64, 22
    opt = torch.optim.Adam(model.parameters(), lr=torch.tensor(2e-2, device="cuda"), capturable=True)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda t: 1.0 / (1.0 + 1e-3 * t))
    model.eval()
    before = model.eval_model(x, y)[1]
    model.train_model(loader, opt, sched, epochs1=20, epochs2=20, cuda_graph=True)
    after = model.eval_model(x, y)[1]
    assert np.isfinite(after) and after < before
    assert float(opt.param_groups[0]["lr"]) < 2e-2   # the scheduler reached the captured step


def test_three_training_steps_match_the_reference(golden):
    """loss -> backward -> optimizer step, three times, against the reference run as written on CPU
    (tests/golden/make_golden.py gen_train): per-step losses and every parameter afterwards."""
    g = golden("train")
    model = W.WHVIRegression([W.WHVILinear(3, 16, lambda_=2.0, semantics="reference"), torch.nn.ReLU(),
                              W.WHVILinear(16, 1, semantics="reference")], train_samples=2)
    model.load_state_dict({k: torch.from_numpy(g[f"init.{k}"]) for k in model.state_dict().keys()})
    model = model.to(dev()).train()
    opt = torch.optim.SGD(model.parameters(), lr=float(g["lr"]))
    blocks = [b for layer in model._whvi_layers() for b in layer.square_blocks()]
    S, n_blocks = 2, len(blocks)
    x, y = t(g["x"]), t(g["y"])
    for step in range(3):
        assert int(g[f"n_eps_{step}"]) == S * n_blocks
        for i, b in enumerate(blocks):
            b.inject_eps(torch.stack([t(g[f"eps_{step}_{s * n_blocks + i}"]) for s in range(S)]))
        loss = model.loss(x, y, n=150)
        loss.backward()
        opt.step()
        opt.zero_grad()
        assert abs(loss.item() - float(g["losses"][step])) < TOL * abs(float(g["losses"][step])), step
    change = max(float(np.abs(g[f"final.{k}"] - g[f"init.{k}"]).max()) for k, _ in model.named_parameters())
    for name, p in model.named_parameters():
        diff = float(np.max(np.abs(p.detach().cpu().numpy().astype(np.float64) - g[f"final.{name}"])))
        assert diff < 1e-3 * change, (name, diff, change)   # 1e-4 per gradient, three accumulated steps


def test_square_blocks_narrower_than_a_float4():
    """D = 1, 2 blocks (WHVILinear(2, 2), stacked layers with 2 inputs) run on the width-4 kernels
    through zero padding; values AND all gradients equal the width-D layer's (fp64 oracle)."""
    from oracle import oracle as O
    from whvi_b200 import functional as F
    for D in (1, 2):
        rng = np.random.default_rng(40 + D)
        S, B = 3, 5
        x, g, dy = rng.standard_normal((S, B, D)), rng.standard_normal((S, D)), rng.standard_normal((S, B, D))
        s1, s2, bias = rng.standard_normal(D), rng.standard_normal(D), rng.standard_normal(D)
        xt, gt, s1t, s2t, bt = (t(v).requires_grad_() for v in (x, g, s1, s2, bias))
        y = F.whvi_layer(xt, gt, s1t, s2t, bt)
        assert y.shape == (S, B, D)
        assert rel_err(y.detach().cpu().numpy(), O.layer_fwd(x, g, s1, s2, bias)) < TOL
        (y * t(dy)).sum().backward()
        rdx, rdg, rds1, rds2, rdb = O.layer_bwd(x, dy, g, s1, s2, want_dbias=True)
        for got, ref in ((xt.grad, rdx), (gt.grad, rdg), (s1t.grad, rds1), (s2t.grad, rds2), (bt.grad, rdb)):
            assert rel_err(got.cpu().numpy(), ref) < TOL
    net = W.WHVIRegression([W.WHVILinear(2, 2), torch.nn.ReLU(), W.WHVILinear(2, 1)], train_samples=3).to(dev()).train()
    xb = torch.randn(7, 2, device=dev())
    net.loss(xb, xb[:, :1], n=7).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())


def test_device_prefetcher_short_last_batch():
    """a DataLoader's shorter last batch must come out with its own length and values (it used to
    raise, or to broadcast a 1-row batch over the whole buffer)."""
    from whvi_b200.utils import DevicePrefetcher
    sizes = [8, 8, 8, 3, 1, 8]
    host = [(torch.randn(n, 5).pin_memory(), torch.randn(n, 1).pin_memory()) for n in sizes]
    seen = 0
    for (hx, hy), (dx, dy) in zip(host, DevicePrefetcher(iter(host), dev())):
        assert dx.shape == hx.shape and dy.shape == hy.shape
        assert torch.equal(dx.cpu(), hx) and torch.equal(dy.cpu(), hy)
        seen += 1
    assert seen == len(sizes)


@pytest.mark.parametrize("D,layers", [(64, 2), (8192, 1), (8192, 2)])
def test_eval_model_from_fused_predictive_sums(D, layers):
    """WHVIRegression.eval_model (src/networks.py:101-115, :130-133) computed from sum_s y_hat and sum_s y_hat^2 reduced
    inside the last layer (no (batch, out, S) tensor) equals the reference formulation on the full prediction tensor."""
    from whvi_b200.networks import WHVINetwork
    torch.manual_seed(D + layers)
    mods = []
    for i in range(layers):
        mods += [W.WHVILinear(D, D, lambda_=2.0, bias=(i == layers - 1))] + ([torch.nn.ReLU()] if i < layers - 1 else [])
    model = W.WHVIRegression(mods, eval_samples=6, sigma=0.7).to(dev())
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith(("s1", "s2")):
                p.mul_(100.0 / D ** 0.5)
            if name.endswith(("g_mu", "bias")):
                p.copy_(torch.randn_like(p))
    B = 5
    x, y = torch.randn(B, D, device=dev()), torch.randn(B, D, device=dev())
    eps = [torch.randn(6, D, device=dev()) for _ in range(layers)]
    out = []
    for fused in (True, False):
        for layer, e in zip(model._whvi_layers(), eps):
            layer.square_blocks()[0].inject_eps(e)
        if fused:
            out.append(model.eval_model(x, y))
        else:
            out.append(WHVINetwork.eval_model(model, x, y, lambda yp, yt: torch.sqrt(((yp.mean(dim=2) - yt) ** 2).mean())))
    assert abs(out[0][0] - out[1][0]) < 1e-4 * abs(out[1][0]), out
    assert abs(out[0][1] - out[1][1]) < 1e-4 * abs(out[1][1]), out
