"""Host-side proofs about the kernels' compile-time layouts and work split (no GPU): builds and
runs tools/sim_layout.cpp, which includes the very headers the kernels use (layout.cuh, plan.cuh)
and checks that every view is a permutation, the shared-memory maps are bijections, transposition
writes/reads are bank-conflict-free, the round structure equals a plain FWHT for every k <= n, and
the grid plans cover every tile without idle CTAs."""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


import pytest


@pytest.mark.parametrize("flags", [[], ["-DWHVI_PADDED=1"]], ids=["xor-swizzle (product)", "padded (variant builds)"])
def test_layout_and_plan_simulator(tmp_path, flags):
    exe = tmp_path / "sim_layout"
    subprocess.run(["g++", "-O2", "-std=c++17", *flags, "-I", str(ROOT / "whvi_b200" / "csrc"), "-o", str(exe),
                    str(ROOT / "tools" / "sim_layout.cpp")], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    sys.stdout.write(out.stdout[-2000:])
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "all layout checks passed" in out.stdout
