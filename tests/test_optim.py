"""Flat parameter storage + fused Adam (the optimizer.step() of the reference's step loop,
src/networks.py:80-82; the reference's optimizer is torch.optim.Adam, src/evaluation.py:15-27 -- which is
therefore the checker here)."""
import copy

import numpy as np
import pytest
import torch

import whvi_b200 as W
from conftest import rel_err
from whvi_b200.optim import FlatAdam, FlatParams


def test_flat_params_views_cpu():
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
    before = [p.detach().clone() for p in net.parameters()]
    keys = list(net.state_dict().keys())
    flat = FlatParams(net.parameters())
    assert list(net.state_dict().keys()) == keys
    assert all(torch.equal(a, b) for a, b in zip(before, net.parameters()))
    assert all(o % 4 == 0 for o in flat.offsets) and flat.numel % 4 == 0
    x = torch.randn(4, 5)
    net(x).sum().backward()
    assert flat.attached() and float(flat.grad.abs().sum()) > 0
    g1 = [p.grad.clone() for p in net.parameters()]
    net(x).sum().backward()          # autograd accumulates into the views in place
    assert flat.attached() and all(torch.allclose(2 * a, p.grad) for a, p in zip(g1, net.parameters()))
    with torch.no_grad():
        flat.param.mul_(0.5)         # the flat buffer IS the parameters
    assert all(torch.allclose(0.5 * a, b) for a, b in zip(before, net.parameters()))
    net.zero_grad(set_to_none=True)  # someone detaches the views ...
    assert not flat.attached()
    flat.zero_grad()                 # ... and they come back
    assert flat.attached() and float(flat.grad.abs().sum()) == 0.0


def test_graph_lr_check_cpu():
    """a float lr with a scheduler would be frozen into the captured graph -- must raise."""
    from whvi_b200.graphs import _check_optimizer
    p = [torch.nn.Parameter(torch.zeros(3))]
    _check_optimizer(torch.optim.Adam(p, lr=1e-3, capturable=True), scheduled=False)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        _check_optimizer(torch.optim.Adam(p, lr=1e-3, capturable=True), scheduled=True)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        _check_optimizer(torch.optim.Adam(p, lr=torch.tensor(1e-3), capturable=True), scheduled=True)
    with pytest.raises(RuntimeError, match="capturable"):
        _check_optimizer(torch.optim.Adam(p, lr=1e-3), scheduled=False)


def _toy(seed):
    torch.manual_seed(seed)
    return W.WHVIRegression([W.WHVILinear(3, 16, lambda_=2.0), torch.nn.ReLU(), W.WHVILinear(16, 16, lambda_=2.0),
                             torch.nn.ReLU(), W.WHVILinear(16, 1)], train_samples=4).cuda().train()


@pytest.mark.gpu
@pytest.mark.parametrize("tensor_lr", [False, True])
def test_flat_adam_matches_torch_adam(tensor_lr):
    a = _toy(1)
    b = copy.deepcopy(a)
    lr = lambda: torch.tensor(3e-2, device="cuda") if tensor_lr else 3e-2
    opt_a = torch.optim.Adam(a.parameters(), lr=lr(), capturable=tensor_lr)
    opt_b = FlatAdam(FlatParams(b.parameters()), lr=lr())
    sch_a = torch.optim.lr_scheduler.LambdaLR(opt_a, lambda t: 1.0 / (1.0 + 0.1 * t))
    sch_b = torch.optim.lr_scheduler.LambdaLR(opt_b, lambda t: 1.0 / (1.0 + 0.1 * t))
    g = torch.Generator(device="cuda").manual_seed(3)
    for step in range(6):
        x = torch.randn(32, 3, device="cuda", generator=g)
        y = x[:, :1] ** 2 - x[:, 1:2]
        for m, o, s in ((a, opt_a, sch_a), (b, opt_b, sch_b)):
            torch.manual_seed(100 + step)          # the same eps draws for both models
            m.loss(x, y, n=150).backward()
            o.step()
            s.step()
            o.zero_grad()
    for (name, p), q in zip(a.named_parameters(), b.parameters()):
        assert rel_err(q.detach().cpu().numpy(), p.detach().cpu().numpy()) < 2e-5, name


@pytest.mark.gpu
def test_flat_adam_in_a_captured_step():
    """FlatAdam is graph-safe (device-resident step count and lr, gradient memset inside the capture)."""
    from whvi_b200.graphs import GraphedTrainStep
    g = torch.Generator(device="cuda").manual_seed(5)
    xs = [torch.randn(32, 3, device="cuda", generator=g) for _ in range(5)]
    ys = [x[:, :1] ** 2 - x[:, 1:2] for x in xs]
    eager = _toy(7)
    with torch.no_grad():
        for n, p in eager.named_parameters():
            if n.endswith(("s1", "s2", "g_mu")):
                p.normal_()                        # O(1) weights: no gradient is rounding noise for Adam to amplify
            if n.endswith("g_rho"):
                p.fill_(-30.0)                     # no MC noise: eager and replayed runs are comparable
    graphed = copy.deepcopy(eager)
    opt_e = FlatAdam(FlatParams(eager.parameters()), lr=torch.tensor(1e-2, device="cuda"))
    opt_g = FlatAdam(FlatParams(graphed.parameters()), lr=torch.tensor(1e-2, device="cuda"))
    step = GraphedTrainStep(graphed, opt_g, xs[0], ys[0], n=150, scheduled=True)
    for p, q in zip(eager.parameters(), graphed.parameters()):
        assert torch.equal(p, q), "warm-up steps of the capture must not train the model"
    assert float(opt_g.step_t) == 0.0
    for x, y in zip(xs, ys):
        eager.loss(x, y, n=150).backward()
        opt_e.step()
        opt_e.zero_grad()
        step(x, y)
    assert float(opt_g.step_t) == len(xs)
    for (name, p), q in zip(eager.named_parameters(), graphed.parameters()):
        assert rel_err(q.detach().cpu().numpy(), p.detach().cpu().numpy()) < 1e-4, name
