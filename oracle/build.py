"""TEST INFRASTRUCTURE ONLY: build recipe for the CPU oracle and the compiled reference.

* ``build_oracle()`` compiles ``oracle/whvi_oracle.c`` (the C restatement) with gcc
  into ``oracle/_build/libwhvi_oracle.so``.
* ``build_ref()`` compiles the reference's own C++ CPU FWHT **from the sources where
  they lie** (``/root/reference/src/fwht/cpp/fwht.cpp``, unmodified, never copied into
  this repo) with a direct ``g++`` command line against the installed torch headers,
  into ``oracle/_ref/fwht_cpp.so``.  That directory is git-ignored but travels to the
  GPU box with ``gpurun``.  It is only attempted when ``/root/reference`` exists (this
  container); the GPU box uses the prebuilt file.

* ``build_ref_cuda()`` compiles the reference's CUDA FWHT extension
  (``/root/reference/src/fwht/cuda/fwht_cuda.cpp`` + ``fwht_cuda_kernel.cu``) for sm_100a
  into ``oracle/_ref/fwht_cuda.so`` -- the GPU baseline "recompiled reference kernel"
  and a parity cross-check for D <= 2^12.  The sources do not compile against torch
  2.11 as they are (``X.type()`` / ``X.data<T>()`` were removed), so the
  recipe streams them through the 2-expression ``sed`` SURVEY App. C documents into a
  temporary directory OUTSIDE the repo (deleted afterwards) and runs ``nvcc`` on that;
  the kernels themselves are untouched.  Slow (torch headers through nvcc, ~3 min):
  only built on request (``python oracle/build.py --ref-cuda``) or by
  ``__graft_entry__.build()`` when the file is missing.

* ``build_ref_py()`` byte-compiles the reference's own Python modules on the hot path
  (``/root/reference/src/{utils,weights,layers,likelihoods,networks}.py`` and
  ``src/fwht/{python,cuda}/fwht.py``) UNMODIFIED, from where they lie, into bytecode files
  ``oracle/_ref/refpy/src/**/*.pycode`` (compiled output only -- no reference source text enters the repo
  or ``oracle/_ref``; the suffix is not ``.pyc`` because the GPU-box snapshot drops ``*.pyc``).
  ``oracle/ref_torch.reference_package()`` imports them through a small finder (with the
  ``fwht_cuda`` extension stubbed), which is what ``bench.py --impl reference`` and the
  ``cpu_baseline`` leg time on the GPU box: the reference's own layer code, not a restatement.

Run ``python oracle/build.py`` to build the first three.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from pathlib import Path

HERE = Path(__file__).resolve().parent
BUILD_DIR = HERE / "_build"
REF_DIR = HERE / "_ref"
ORACLE_SO = BUILD_DIR / "libwhvi_oracle.so"
REF_SO = REF_DIR / "fwht_cpp.so"
REF_SRC = Path("/root/reference/src/fwht/cpp/fwht.cpp")
REF_CUDA_SO = REF_DIR / "fwht_cuda.so"
REF_CUDA_DIR = Path("/root/reference/src/fwht/cuda")


def _newer(target: Path, *sources: Path) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(s.exists() and s.stat().st_mtime <= t for s in sources)


def build_oracle(force: bool = False) -> Path:
    src = HERE / "whvi_oracle.c"
    hdr = HERE / "whvi_oracle_impl.h"
    if not force and _newer(ORACLE_SO, src, hdr):
        return ORACLE_SO
    BUILD_DIR.mkdir(exist_ok=True)
    # no -march=native: the .so travels to the GPU box, whose host CPU may differ
    cmd = ["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", "-o", str(ORACLE_SO), str(src), "-lm"]
    subprocess.run(cmd, check=True, cwd=str(HERE))
    return ORACLE_SO


def build_ref(force: bool = False) -> Path | None:
    """Compile the unmodified reference fwht.cpp into oracle/_ref/fwht_cpp.so."""
    if not REF_SRC.exists():
        return REF_SO if REF_SO.exists() else None
    if not force and _newer(REF_SO, REF_SRC):
        return REF_SO
    import torch
    from torch.utils import cpp_extension

    REF_DIR.mkdir(exist_ok=True)
    torch_lib = Path(torch.__file__).resolve().parent / "lib"
    inc = [f"-I{p}" for p in cpp_extension.include_paths()]
    inc.append(f"-I{sysconfig.get_paths()['include']}")
    cmd = [
        "g++", "-O2", "-std=c++17", "-shared", "-fPIC",
        "-DTORCH_EXTENSION_NAME=fwht_cpp", "-DTORCH_API_INCLUDE_EXTENSION_H",
        f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
        *inc, str(REF_SRC), "-o", str(REF_SO),
        f"-L{torch_lib}", "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python",
        f"-Wl,-rpath,{torch_lib}",
    ]
    subprocess.run(cmd, check=True)
    return REF_SO


def build_ref_cuda(force: bool = False) -> Path | None:
    """Compile the reference CUDA FWHT (API-renames only) into oracle/_ref/fwht_cuda.so."""
    srcs = [REF_CUDA_DIR / "fwht_cuda.cpp", REF_CUDA_DIR / "fwht_cuda_kernel.cu"]
    if not all(p.exists() for p in srcs):
        return REF_CUDA_SO if REF_CUDA_SO.exists() else None
    if not force and _newer(REF_CUDA_SO, *srcs):
        return REF_CUDA_SO
    import re
    import tempfile

    import torch
    from torch.utils import cpp_extension

    REF_DIR.mkdir(exist_ok=True)
    torch_lib = Path(torch.__file__).resolve().parent / "lib"
    inc = [f"-I{p}" for p in cpp_extension.include_paths(device_type="cuda")]
    inc.append(f"-I{sysconfig.get_paths()['include']}")
    with tempfile.TemporaryDirectory(prefix="whvi_refcuda_") as tmp:
        patched = []
        for p in srcs:
            text = p.read_text()
            # torch >= 1.? API renames only; kernels are untouched
            text = re.sub(r"\bX\.type\(\)", "X.scalar_type()", text)
            text = re.sub(r"\bX\.data<scalar_t>\(\)", "X.data_ptr<scalar_t>()", text)
            q = Path(tmp) / p.name
            q.write_text(text)
            patched.append(str(q))
        cmd = [
            "nvcc", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
            "-gencode", "arch=compute_100a,code=sm_100a",
            "-DTORCH_EXTENSION_NAME=fwht_cuda", "-DTORCH_API_INCLUDE_EXTENSION_H",
            f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
            "--expt-relaxed-constexpr", "-diag-suppress=20012,20013,20014,20015",
            *inc, *patched, "-o", str(REF_CUDA_SO),
            f"-L{torch_lib}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
            "-Xlinker", f"-rpath={torch_lib}",
        ]
        subprocess.run(cmd, check=True)
    return REF_CUDA_SO


REF_PY_DIR = REF_DIR / "refpy"
REF_PY_SRC = Path("/root/reference/src")
REF_PY_MODULES = ["utils", "weights", "layers", "likelihoods", "networks", "fwht/python/fwht", "fwht/cuda/fwht",
                  "fwht/python/__init__", "fwht/cuda/__init__"]


def build_ref_py(force: bool = False) -> Path | None:
    """Byte-compile the unmodified reference modules into oracle/_ref/refpy/src/**.pycode (imported by ref_torch's finder)."""
    import py_compile
    marker = REF_PY_DIR / "src" / "weights.pycode"
    if not REF_PY_SRC.exists():
        return REF_PY_DIR if marker.exists() else None
    if not force and _newer(marker, *(REF_PY_SRC / f"{m}.py" for m in REF_PY_MODULES)):
        return REF_PY_DIR
    for m in REF_PY_MODULES:
        src = REF_PY_SRC / f"{m}.py"
        if not src.exists():
            continue
        out = REF_PY_DIR / "src" / f"{m}.pycode"
        out.parent.mkdir(parents=True, exist_ok=True)
        # dfile: the path recorded in the code object (tracebacks) -- the reference's own path
        # unchecked-hash pyc: validity does not depend on a source file (there is none on the GPU box)
        py_compile.compile(str(src), cfile=str(out), dfile=f"reference/src/{m}.py", doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    return REF_PY_DIR


def main() -> int:
    print("oracle:", build_oracle(force="--force" in sys.argv))
    print("reference fwht_cpp:", build_ref(force="--force" in sys.argv))
    print("reference python modules (bytecode):", build_ref_py(force="--force" in sys.argv))
    if "--ref-cuda" in sys.argv:
        print("reference fwht_cuda:", build_ref_cuda(force="--force" in sys.argv))
    return 0


if __name__ == "__main__":
    sys.exit(main())
