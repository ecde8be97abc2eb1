"""CPU: the C-ABI library loads, exports every symbol include/whvi_b200.h declares, and
the Python frontends reject bad input the way the reference's TORCH_CHECKs do.  No
compute call is made (there is no GPU here)."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "whvi_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"WHVI_API\s+[\w\s\*]+?\b(whvi_\w+)\s*\(", text)))


def test_header_declares_something():
    syms = declared_symbols()
    assert "whvi_fwht_f32" in syms and "whvi_last_error" in syms


def test_library_exports_every_declared_symbol():
    from whvi_b200 import _lib
    L = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared_symbols():
        assert hasattr(L, name), f"{name} declared in include/whvi_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in whvi_b200/_lib.py"
    assert _lib.lib().whvi_abi_version() == 1
    assert _lib.lib().whvi_max_dim() >= 1 << 20


def test_fwht_frontend_rejects_like_the_reference():
    from whvi_b200 import FWHTFunction
    with pytest.raises(RuntimeError, match="X must be a CUDA tensor"):
        FWHTFunction.apply(torch.randn(2, 4))


def test_no_oracle_import_in_product():
    """The product path must never route through oracle/ (or /root/reference)."""
    for py in (ROOT / "whvi_b200").rglob("*.py"):
        src = py.read_text()
        assert "oracle" not in src.replace("# oracle", ""), f"{py} mentions the oracle"
        assert "/root/reference" not in src
