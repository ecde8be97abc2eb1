"""TEST INFRASTRUCTURE / CPU BASELINE ONLY: torch-CPU restatement of the reference's layer
path AS WRITTEN, for timing on the GPU box's host cores (the reference's Python package
cannot travel there).  Used by bench.py's ``cpu_baseline`` leg and ``--impl reference``.

Restates, op for op:
  * ``fwht_cat``     -- the vectorised butterfly, reference src/fwht/python/fwht.py:52-55
  * ``wht_matmul``   -- dense ``(H @ x.T).T``, src/fwht/python/fwht.py:17-32
  * ``fwht_cpu``     -- the CPU dispatch of WHVISquarePow2Matrix.fwht, src/weights.py:37-41
                        (matmul below D = 2^12, vectorised butterfly from there on)
  * ``w_bar``        -- src/weights.py:66-73 (diag(s2) -> fwht -> row-scale -> fwht -> row-scale)
  * ``sample_lrt``   -- src/weights.py:87-93: h @ (w_bar(mu) + w_bar(sigma*eps)).T
Autograd provides the backward exactly as it does in the reference.  When
``oracle/_ref/fwht_cpp.so`` (the reference's own C++ FWHT, compiled unmodified) is present,
``fwht_cpp_forward`` exposes it for the FWHT CPU baseline (kind = "reference").
"""
from __future__ import annotations

import math
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

_H_CACHE: dict[int, torch.Tensor] = {}


def build_H(D: int) -> torch.Tensor:  # src/utils.py:74-101
    if D not in _H_CACHE:
        H = torch.ones(1, 1)
        while H.size(0) < D:
            H = torch.cat([torch.cat([H, H], dim=1), torch.cat([H, -H], dim=1)], dim=0)
        _H_CACHE[D] = H
    return _H_CACHE[D]


def wht_matmul(x: torch.Tensor) -> torch.Tensor:
    return (build_H(x.size(1)) @ x.T).T


def fwht_cat(x: torch.Tensor) -> torch.Tensor:
    out = x.unsqueeze(2)
    for _ in range(int(math.log2(x.shape[1]))):
        out = torch.cat((out[:, ::2] + out[:, 1::2], out[:, ::2] - out[:, 1::2]), dim=2)
    return out.squeeze(1)


def fwht_cpu(x: torch.Tensor) -> torch.Tensor:
    return wht_matmul(x) if x.size(1) < 2 ** 12 else fwht_cat(x)


def w_bar(u, s1, s2):
    rs = lambda d, A: (d * A.T).T  # matmul_diag_left, src/utils.py:4-12
    return rs(s1, fwht_cpu(rs(u, fwht_cpu(torch.diag(s2)))))


def sample_lrt(h, s1, s2, g_mu, g_rho, eps):
    return h @ (w_bar(g_mu, s1, s2) + w_bar(F.softplus(g_rho) * eps, s1, s2)).T


def layer_fwd_bwd_seconds(D: int, B: int, n_samples: int = 1, threads: int | None = None) -> float:
    """Wall time of n_samples forward+backward passes of one as-written WHVILinear(D, D)
    on B rows, all host threads."""
    import time
    if threads:
        torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    s1 = (torch.randn(D, generator=g) * 0.01).requires_grad_()
    s2 = (torch.randn(D, generator=g) * 0.01).requires_grad_()
    mu = torch.zeros(D, requires_grad=True)
    rho = (torch.rand(D, generator=g) - 3).requires_grad_()
    h = torch.randn(B, D, generator=g, requires_grad=True)
    dy = torch.randn(B, D, generator=g)
    t0 = time.perf_counter()
    for _ in range(n_samples):
        eps = torch.randn(D, generator=g)
        y = sample_lrt(h, s1, s2, mu, rho, eps)
        y.backward(dy)
    return time.perf_counter() - t0


def fwht_cpp_module():
    """The reference's own compiled C++ FWHT (oracle/_ref/fwht_cpp.so) or None."""
    ref_dir = Path(__file__).resolve().parent / "_ref"
    if not (ref_dir / "fwht_cpp.so").exists():
        return None
    if str(ref_dir) not in sys.path:
        sys.path.insert(0, str(ref_dir))
    try:
        import fwht_cpp
        return fwht_cpp
    except Exception:
        return None


def fwht_cuda_module():
    """The reference's CUDA FWHT extension recompiled for sm_100a (oracle/_ref/fwht_cuda.so, built by
    oracle/build.py --ref-cuda with the torch-API renames of SURVEY F3 only) or None.  A GPU
    baseline and parity cross-check for D <= 2^12 (its launch shape is invalid beyond)."""
    ref_dir = Path(__file__).resolve().parent / "_ref"
    if not (ref_dir / "fwht_cuda.so").exists():
        return None
    if str(ref_dir) not in sys.path:
        sys.path.insert(0, str(ref_dir))
    try:
        import fwht_cuda
        return fwht_cuda
    except Exception:
        return None


class _RefBytecodeFinder:
    """Import ``src`` / ``src.*`` from the bytecode files oracle/build.py:build_ref_py wrote (``*.pycode`` = a standard
    .pyc image under another suffix).  Packages are the directories; nothing but the compiled reference code is run."""

    def __init__(self, root: Path):
        self.root = root

    def find_spec(self, name, path=None, target=None):
        import importlib.machinery
        import importlib.util
        if name != "src" and not name.startswith("src."):
            return None
        rel = self.root.joinpath(*name.split("."))
        if rel.is_dir():   # package: optional __init__.pycode
            init = rel / "__init__.pycode"
            spec = importlib.machinery.ModuleSpec(name, self, origin=str(init if init.exists() else rel), is_package=True)
            spec.submodule_search_locations = [str(rel)]
            return spec
        code = rel.with_suffix(".pycode")
        if code.exists():
            return importlib.machinery.ModuleSpec(name, self, origin=str(code))
        return None

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        import marshal
        origin = Path(module.__spec__.origin)
        if origin.is_dir() or not origin.exists():
            return
        data = origin.read_bytes()
        exec(marshal.loads(data[16:]), module.__dict__)   # 16-byte pyc header (magic, flags, hash), then the code object


def reference_package():
    """The reference's OWN layer code (oracle/_ref/refpy: its modules byte-compiled unmodified by
    oracle/build.py:build_ref_py) or None when it did not travel.  ``fwht_cuda`` -- imported unconditionally by
    src/weights.py:8 even on CPU -- is stubbed exactly as tests/golden/make_golden.py does.
    Returns the imported ``src.layers`` module."""
    import importlib
    import importlib.util
    import types
    ref_dir = Path(__file__).resolve().parent / "_ref" / "refpy"
    if not (ref_dir / "src" / "weights.pycode").exists():
        return None
    if importlib.util.MAGIC_NUMBER != (ref_dir / "src" / "weights.pycode").read_bytes()[:4]:
        return None   # compiled by another interpreter version
    if not any(isinstance(f, _RefBytecodeFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _RefBytecodeFinder(ref_dir))
    sys.modules.setdefault("fwht_cuda", types.ModuleType("fwht_cuda"))
    try:
        return importlib.import_module("src.layers")
    except Exception:
        return None


def reference_layer_fwd_bwd_seconds(D: int, B: int, n_samples: int = 1, seed: int = 0):
    """Wall time of n_samples forward+backward passes of the reference's own ``WHVILinear(D, D)`` on B rows
    (its CPU path as written: src/weights.py:34-41, :66-93), all host threads; None if the package is absent."""
    import time
    layers = reference_package()
    if layers is None:
        return None
    torch.manual_seed(seed)
    layer = layers.WHVILinear(D, D)
    h = torch.randn(B, D, requires_grad=True)
    dy = torch.randn(B, D)
    t0 = time.perf_counter()
    for _ in range(n_samples):
        y = layer(h)          # one eps draw per call = one MC sample (src/weights.py:92)
        y.backward(dy)
    return time.perf_counter() - t0
