"""TEST INFRASTRUCTURE ONLY: Python face of the CPU oracle.

Two layers:

* ctypes wrappers over ``oracle/_build/libwhvi_oracle.so`` (the plain-C restatement in
  ``whvi_oracle.c`` / ``whvi_oracle_impl.h``), numpy in / numpy out, fp32 and fp64;
* tiny numpy restatements used to pin the C code: ``build_H`` (reference
  ``src/utils.py:74-101``), ``fwht_dense`` (``(H @ a.T).T`` as in ``test/walsh.py:26``)
  and ``fwht_cat`` (the vectorised butterfly of ``src/fwht/python/fwht.py:52-55``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs import
this module.  Nothing under ``whvi_b200/`` does.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, c_double, c_float, c_int, c_int64
from pathlib import Path

import numpy as np

from . import build as _build

_LIB = None


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        so = _build.ORACLE_SO
        if not so.exists():
            _build.build_oracle()
        _LIB = ctypes.CDLL(str(so))
        _declare(_LIB)
    return _LIB


def _declare(L: ctypes.CDLL) -> None:
    for suf, ct in (("f32", c_float), ("f64", c_double)):
        P = POINTER(ct)
        getattr(L, f"oracle_fwht_{suf}").argtypes = [P, P, c_int64, c_int64]
        getattr(L, f"oracle_fwht_{suf}").restype = None
        getattr(L, f"oracle_reparam_{suf}").argtypes = [P, P, P, P, c_int64, c_int64, c_int]
        getattr(L, f"oracle_reparam_{suf}").restype = None
        getattr(L, f"oracle_reparam_bwd_{suf}").argtypes = [P, P, P, P, P, c_int64, c_int64]
        getattr(L, f"oracle_reparam_bwd_{suf}").restype = None
        getattr(L, f"oracle_layer_fwd_{suf}").argtypes = [P, c_int64, P, P, P, P, P, c_int64, c_int64, c_int64]
        getattr(L, f"oracle_layer_fwd_{suf}").restype = None
        getattr(L, f"oracle_layer_bwd_{suf}").argtypes = [P, c_int64, P, P, P, P, P, P, P, P, P,
                                                         c_int64, c_int64, c_int64]
        getattr(L, f"oracle_layer_bwd_{suf}").restype = None
        getattr(L, f"oracle_ref_w_bar_{suf}").argtypes = [P, P, P, P, c_int64]
        getattr(L, f"oracle_ref_w_bar_{suf}").restype = None
        getattr(L, f"oracle_ref_sample_lrt_{suf}").argtypes = [P, P, P, P, P, P, c_int64, c_int64]
        getattr(L, f"oracle_ref_sample_lrt_{suf}").restype = None
        getattr(L, f"oracle_paper_weight_{suf}").argtypes = [P, P, P, P, c_int64]
        getattr(L, f"oracle_paper_weight_{suf}").restype = None
        getattr(L, f"oracle_kl_{suf}").argtypes = [P, P, c_double, c_int64, c_int, P, P]
        getattr(L, f"oracle_kl_{suf}").restype = c_double
        getattr(L, f"oracle_mnll_{suf}").argtypes = [P, P, c_double, c_int64, c_int64, c_int64, c_int64]
        getattr(L, f"oracle_mnll_{suf}").restype = c_double


def _suf(dtype) -> tuple[str, type]:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32", c_float
    if dtype == np.float64:
        return "f64", c_double
    raise TypeError(f"oracle supports float32/float64, got {dtype}")


def _p(a: np.ndarray | None, ct):
    if a is None:
        return ctypes.cast(None, POINTER(ct))
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(POINTER(ct))


def _c(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a), dtype=dtype)


# --------------------------------------------------------------------------- numpy pins
def build_H(D: int, dtype=np.float64) -> np.ndarray:
    """Sylvester Walsh-Hadamard matrix, as reference src/utils.py:74-101 builds it."""
    assert D >= 1 and D & (D - 1) == 0
    H = np.ones((1, 1), dtype=dtype)
    while H.shape[0] < D:
        H = np.block([[H, H], [H, -H]])
    return H


def fwht_dense(a: np.ndarray) -> np.ndarray:
    """(H @ a.T).T, the reference tests' dense check (test/walsh.py:26)."""
    a = np.asarray(a)
    return (build_H(a.shape[1], a.dtype) @ a.T).T


def fwht_cat(a: np.ndarray) -> np.ndarray:
    """numpy restatement of the vectorised butterfly, src/fwht/python/fwht.py:52-55."""
    a = np.asarray(a)
    x = a[:, :, None]
    for _ in range(int(np.log2(a.shape[1]))):
        x = np.concatenate((x[:, ::2] + x[:, 1::2], x[:, ::2] - x[:, 1::2]), axis=2)
    return x[:, 0, :]


# --------------------------------------------------------------------------- C wrappers
def fwht(a: np.ndarray) -> np.ndarray:
    """Batched FWHT of the rows of ``a`` (rows, D); follows src/fwht/cpp/fwht.cpp:3-21."""
    a = np.asarray(a)
    suf, ct = _suf(a.dtype)
    assert a.ndim == 2
    a = _c(a, a.dtype)
    out = np.empty_like(a)
    getattr(lib(), f"oracle_fwht_{suf}")(_p(a, ct), _p(out, ct), a.shape[0], a.shape[1])
    return out


def softplus(r: np.ndarray) -> np.ndarray:
    r = np.asarray(r)
    return np.where(r > 20, r, np.log1p(np.exp(np.minimum(r, 20))))


def reparam(mu, rho_or_L, eps, dense: bool = False) -> np.ndarray:
    """g[s] = mu + softplus(rho)*eps[s] (src/weights.py:82-83) or mu + L eps[s]."""
    eps = np.asarray(eps)
    suf, ct = _suf(eps.dtype)
    S, D = eps.shape
    mu, p, eps = _c(mu, eps.dtype), _c(rho_or_L, eps.dtype), _c(eps, eps.dtype)
    g = np.empty((S, D), dtype=eps.dtype)
    getattr(lib(), f"oracle_reparam_{suf}")(_p(mu, ct), _p(p, ct), _p(eps, ct), _p(g, ct), S, D, int(dense))
    return g


def reparam_bwd(rho, eps, dg):
    eps = np.asarray(eps)
    suf, ct = _suf(eps.dtype)
    S, D = eps.shape
    rho, eps, dg = _c(rho, eps.dtype), _c(eps, eps.dtype), _c(dg, eps.dtype)
    dmu = np.empty(D, dtype=eps.dtype)
    drho = np.empty(D, dtype=eps.dtype)
    getattr(lib(), f"oracle_reparam_bwd_{suf}")(_p(rho, ct), _p(eps, ct), _p(dg, ct), _p(dmu, ct), _p(drho, ct), S, D)
    return dmu, drho


def layer_fwd(x, g, s1, s2, bias=None) -> np.ndarray:
    """PAPER forward y[s,b] = s1*H(g[s]*H(s2*x[s,b])) (+bias); x is (B,D) (shared by all
    samples) or (S,B,D); g is (S,D).  Docstring formula at src/weights.py:77."""
    g = np.asarray(g)
    dt = g.dtype
    suf, ct = _suf(dt)
    S, D = g.shape
    x = _c(x, dt)
    if x.ndim == 2:
        B, xs = x.shape[0], 0
    else:
        assert x.shape[0] == S
        B, xs = x.shape[1], x.shape[1] * D
    s1, s2, g = _c(s1, dt), _c(s2, dt), _c(g, dt)
    bias = None if bias is None else _c(bias, dt).reshape(-1)
    y = np.empty((S, B, D), dtype=dt)
    getattr(lib(), f"oracle_layer_fwd_{suf}")(_p(x, ct), xs, _p(g, ct), _p(s1, ct), _p(s2, ct), _p(bias, ct),
                                              _p(y, ct), S, B, D)
    return y


def layer_bwd(x, dy, g, s1, s2, want_dbias: bool = False):
    """PAPER backward.  Returns dx (S,B,D) [summed over s if x was
    shared], dg (S,D), ds1, ds2 (D,), and dbias when asked."""
    g = np.asarray(g)
    dt = g.dtype
    suf, ct = _suf(dt)
    S, D = g.shape
    x = _c(x, dt)
    shared = x.ndim == 2
    B, xs = (x.shape[0], 0) if shared else (x.shape[1], x.shape[1] * D)
    dy = _c(dy, dt).reshape(S, B, D)
    s1, s2, g = _c(s1, dt), _c(s2, dt), _c(g, dt)
    dx = np.empty((S, B, D), dtype=dt)
    dg = np.empty((S, D), dtype=dt)
    ds1 = np.empty(D, dtype=dt)
    ds2 = np.empty(D, dtype=dt)
    dbias = np.empty(D, dtype=dt) if want_dbias else None
    getattr(lib(), f"oracle_layer_bwd_{suf}")(_p(x, ct), xs, _p(dy, ct), _p(g, ct), _p(s1, ct), _p(s2, ct),
                                              _p(dx, ct), _p(dg, ct), _p(ds1, ct), _p(ds2, ct), _p(dbias, ct),
                                              S, B, D)
    if shared:
        dx = dx.sum(axis=0)
    out = (dx, dg, ds1, ds2)
    return out + (dbias,) if want_dbias else out


def ref_w_bar(u, s1, s2) -> np.ndarray:
    """Reference-as-written w_bar(u), src/weights.py:66-73 (D x D)."""
    u = np.asarray(u)
    dt = u.dtype
    suf, ct = _suf(dt)
    D = u.shape[0]
    u, s1, s2 = _c(u, dt), _c(s1, dt), _c(s2, dt)
    W = np.empty((D, D), dtype=dt)
    getattr(lib(), f"oracle_ref_w_bar_{suf}")(_p(u, ct), _p(s1, ct), _p(s2, ct), _p(W, ct), D)
    return W


def ref_sample_lrt(h, mu, sig_eps, s1, s2) -> np.ndarray:
    """Reference-as-written sample_lrt for one MC sample, src/weights.py:87-93."""
    h = np.asarray(h)
    dt = h.dtype
    suf, ct = _suf(dt)
    B, D = h.shape
    h, mu, se, s1, s2 = (_c(v, dt) for v in (h, mu, sig_eps, s1, s2))
    y = np.empty((B, D), dtype=dt)
    getattr(lib(), f"oracle_ref_sample_lrt_{suf}")(_p(h, ct), _p(mu, ct), _p(se, ct), _p(s1, ct), _p(s2, ct),
                                                   _p(y, ct), B, D)
    return y


def paper_weight(g, s1, s2) -> np.ndarray:
    """Dense W = diag(s1) H diag(g) H diag(s2) (D x D)."""
    g = np.asarray(g)
    dt = g.dtype
    suf, ct = _suf(dt)
    D = g.shape[0]
    g, s1, s2 = _c(g, dt), _c(s1, dt), _c(s2, dt)
    W = np.empty((D, D), dtype=dt)
    getattr(lib(), f"oracle_paper_weight_{suf}")(_p(g, ct), _p(s1, ct), _p(s2, ct), _p(W, ct), D)
    return W


def kl(mu, rho, lambda_: float, mode: int = 0, grads: bool = False):
    """KL as src/utils.py:49-71 called from src/weights.py:52-64 (mode 0 = reference,
    variance interpretation; mode 1 = sigma squared)."""
    mu = np.asarray(mu)
    dt = mu.dtype
    suf, ct = _suf(dt)
    D = mu.shape[0]
    mu, rho = _c(mu, dt), _c(rho, dt)
    dmu = np.empty(D, dtype=dt) if grads else None
    drho = np.empty(D, dtype=dt) if grads else None
    v = getattr(lib(), f"oracle_kl_{suf}")(_p(mu, ct), _p(rho, ct), float(lambda_), D, int(mode), _p(dmu, ct),
                                           _p(drho, ct))
    return (v, dmu, drho) if grads else v


def reparam_dense_bwd(eps, dg):
    """Gradients of g[s] = mu + L eps[s] (dense lower-triangular L; superset of src/weights.py:82-83, parity unpinned --
    not in the reference): dmu = sum_s dg[s], dL = tril(dg^T eps)."""
    eps, dg = np.asarray(eps, dtype=np.float64), np.asarray(dg, dtype=np.float64)
    return dg.sum(0), np.tril(dg.T @ eps)


def kl_dense(mu, L, lambda_: float, grads: bool = False):
    """KL( N(mu, L L^T) || N(0, lambda I) ), L lower triangular with a positive diagonal (entries above the diagonal are
    ignored): 0.5 (D ln lambda - 2 sum ln L_ii - D + |L|_F^2 / lambda + |mu|^2 / lambda).  The dense superset of
    src/utils.py:49-71 (parity unpinned -- the reference's posterior is diagonal); with L = diag(sigma) it equals
    kl(mode=1), which tests/test_oracle.py checks."""
    mu, L = np.asarray(mu, dtype=np.float64), np.tril(np.asarray(L, dtype=np.float64))
    D = mu.shape[0]
    d = np.diagonal(L)
    v = 0.5 * (D * np.log(lambda_) - 2.0 * np.log(d).sum() - D + (L * L).sum() / lambda_ + (mu * mu).sum() / lambda_)
    if not grads:
        return float(v)
    return float(v), mu / lambda_, L / lambda_ - np.diag(1.0 / d)


def mnll(y, y_hat, sigma: float, n: int) -> float:
    """MNLL estimator, src/likelihoods.py:18-29; y (m,n_out), y_hat (m,n_out,n_mc)."""
    y_hat = np.asarray(y_hat)
    dt = y_hat.dtype
    suf, ct = _suf(dt)
    m, n_out, n_mc = y_hat.shape
    y, y_hat = _c(y, dt), _c(y_hat, dt)
    return getattr(lib(), f"oracle_mnll_{suf}")(_p(y, ct), _p(y_hat, ct), float(sigma), int(n), m, n_out, n_mc)


# --------------------------------------------------------------------------- composed layers
def next_pow2(n: int) -> int:
    return 1 << max(0, (int(n) - 1).bit_length())


def stacked_dims(n_in: int, n_out: int):
    """WHVIStackedMatrix.setup_dimensions, src/weights.py:135-160, with integer bit ops
    instead of math.log (the `next_power == 2*D_in` branch only compensates
    float round-off at exact powers of two)."""
    D_in = next_pow2(n_in)
    padding = D_in - n_in
    stack = -(-n_out // D_in)
    return D_in, D_in * stack, padding, stack


def column_weight_paper(g, s1, s2, n: int) -> np.ndarray:
    """PAPER Column weights (src/weights.py:239-245 flattens the sampled D x D matrix and
    keeps the first n entries = row 0): w = s1[0] * (s2 * H g)[:n]."""
    g = np.asarray(g)
    Hg = fwht(g.reshape(1, -1))[0]
    return (np.asarray(s1)[0] * np.asarray(s2) * Hg)[:n]
