"""CPU: the drop-in modules keep the reference's construction-time contract (state_dict
keys, parameter init under a seed, shape dispatch, MNLL formula).  Forward passes need the
GPU and live in test_modules_gpu.py."""
import numpy as np
import pytest
import torch

import whvi_b200 as W
from conftest import rel_err


def test_state_dict_keys_and_init_match_reference(golden):
    g = golden("init")
    torch.manual_seed(0)
    model = W.WHVIRegression([W.WHVILinear(3, 16, lambda_=2.0), torch.nn.ReLU(), W.WHVILinear(16, 16, bias=True),
                              torch.nn.ReLU(), W.WHVILinear(16, 1)])
    keys = list(model.state_dict().keys())
    assert keys == [str(k) for k in g["__keys__"]]
    for k, v in model.state_dict().items():
        assert np.array_equal(v.numpy(), g[k]), k  # same RNG draw order => bit-identical init


def test_shape_dispatch():
    # SURVEY Appendix C probe of the reference's dispatch
    assert isinstance(W.WHVILinear(3, 16).weight_submodule, W.WHVIStackedMatrix)
    assert isinstance(W.WHVILinear(16, 1).weight_submodule, W.WHVIColumnMatrix)
    assert W.WHVILinear(16, 1).weight_submodule.transposed
    assert isinstance(W.WHVILinear(1, 128).weight_submodule, W.WHVIColumnMatrix)
    assert isinstance(W.WHVILinear(128, 128).weight_submodule, W.WHVISquarePow2Matrix)
    st = W.WHVILinear(13, 128).weight_submodule
    assert (st.D_in, st.D_out, st.padding, st.stack) == (16, 128, 3, 8)
    assert W.WHVIStackedMatrix.setup_dimensions(3, 16) == (4, 16, 1, 4)
    for p in (29, 31):  # exact powers where the reference's float log needed its fix-up branch
        assert W.WHVIStackedMatrix.setup_dimensions(2 ** p, 2 ** p) == (2 ** p, 2 ** p, 0, 1)
    assert sum(p.numel() for p in W.WHVILinear(13, 128).parameters()) == 512


def test_mnll_matches_reference(golden):
    g = golden("mnll")
    for name, sigma, n in (("fixed", 1.0, 12), ("rand", 15.21, 116), ("multi", 0.7, 40)):
        lik = W.GaussianLikelihood(sigma)
        v = lik.mnll_batch_estimate(torch.from_numpy(g[f"{name}_y"]), torch.from_numpy(g[f"{name}_yhat"]), n)
        assert abs(float(v) - float(g[f"{name}_mnll"])) < 1e-4 * max(1.0, abs(float(g[f"{name}_mnll"])))


def test_no_cpu_fallback():
    layer = W.WHVILinear(8, 8)
    with pytest.raises(RuntimeError):
        layer(torch.randn(4, 8))
    with pytest.raises(RuntimeError):
        _ = layer.kl


def test_kl_diag_normal_helper(golden):
    from whvi_b200.utils import kl_diag_normal
    g = golden("kl")
    v = kl_diag_normal(*(torch.from_numpy(g[f"gen_{k}"]) for k in ("mu1", "sd1", "mu2", "sd2")))
    assert abs(float(v) - float(g["gen_kl"])) < 1e-4


def test_dimensions_and_dispatch_match_reference_grid(golden):
    """setup_dimensions (integer arithmetic here, float log + fix-up branch in the reference,
    src/weights.py:150-152) and WHVILinear's shape dispatch / parameter counts on a grid generated
    by the reference itself."""
    g = golden("dims")
    ins, outs, table, kinds, counts = (np.asarray(g[k]) for k in ("ins", "outs", "table", "kinds", "counts"))  # decompress once
    for i, a in enumerate(ins):
        for j, b in enumerate(outs):
            assert W.WHVIStackedMatrix.setup_dimensions(int(a), int(b)) == tuple(int(v) for v in table[i, j]), (a, b)
    for a in range(1, 41):
        for b in range(1, 41):
            layer = W.WHVILinear(a, b)
            sub = layer.weight_submodule
            kind = {W.WHVIColumnMatrix: 0, W.WHVISquarePow2Matrix: 2, W.WHVIStackedMatrix: 3}[type(sub)]
            if kind == 0 and sub.transposed:
                kind = 1
            assert kind == int(kinds[a - 1, b - 1]), (a, b)
            assert sum(p.numel() for p in layer.parameters()) == int(counts[a - 1, b - 1]), (a, b)


def test_uniform_stride_detection():
    """Host logic of the grouped Stacked launch: per-block parameter vectors are read where they lie only when they are evenly
    spaced, contiguous and 16-byte aligned relative to each other."""
    import torch
    from whvi_b200 import functional as F
    store = torch.zeros(5, 4, 16)
    assert F.uniform_stride([store[k, 0] for k in range(5)]) == 64
    assert F.uniform_stride([store[k, 2] for k in range(5)]) == 64
    assert F.uniform_stride([store[0, 1]]) == 16
    assert F.uniform_stride([store[0, 0], store[2, 0], store[3, 0]]) is None            # uneven
    assert F.uniform_stride([torch.zeros(16) for _ in range(3)]) in (None, 16, 32, 64, 128)  # separate allocations: whatever they are, not a crash
    flat = torch.zeros(100)
    assert F.uniform_stride([flat[0:16], flat[18:34]]) is None                           # 72 bytes apart: not a multiple of 16
    assert F.uniform_stride([flat[0:16], flat[8:24]]) is None                            # overlapping
