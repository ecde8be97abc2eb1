"""Build libwhvi_b200.so (the C-ABI CUDA library) in-tree with plain nvcc for sm_100a.

    python -m whvi_b200.build [--force] [--verbose]
    python -m whvi_b200.build --variant NAME -DFOO=1 ...   # A/B builds: tools/lab/libwhvi_b200_NAME.so, own object
                                                           # directory, extra -D flags; time it with tools/lab_bwd.py

No torch headers are involved, so the whole library builds in seconds.  The .so lands
next to this file (git-ignored, but shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "_obj"
LIB = PKG / "libwhvi_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-std=c++17", "-O3", "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
          "-Xcompiler", "-fvisibility=hidden", "-DWHVI_BUILDING"]


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _deps() -> list[Path]:
    return sorted(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "whvi_b200.h"]


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False, variant: str | None = None, defines: tuple[str, ...] = ()) -> Path:
    obj_dir, lib_path = OBJ, LIB
    if variant:  # experiment build: never touches the product library
        lab = PKG.parent / "tools" / "lab"
        lab.mkdir(parents=True, exist_ok=True)
        obj_dir, lib_path = lab / f"_obj_{variant}", lab / f"libwhvi_b200_{variant}.so"
    obj_dir.mkdir(exist_ok=True)
    hdrs = _deps()
    jobs = []
    for src in sources():
        obj = obj_dir / (src.stem + ".o")
        if force or _stale(obj, [src] + hdrs):
            cmd = [NVCC, *ARCH, *CFLAGS, *defines, "-c", str(src), "-o", str(obj)]
            if verbose:
                cmd[1:1] = ["-Xptxas", "-v"]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError(f"nvcc failed: {' '.join(cmd)}\n{r.stderr}")

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    objs = [obj_dir / (s.stem + ".o") for s in sources()]
    if force or jobs or _stale(lib_path, objs):
        run([NVCC, *ARCH, "-shared", "-o", str(lib_path), *map(str, objs), "-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    if not variant:
        build_fastcall(force)
    return lib_path


def build_fastcall(force: bool = False) -> Path | None:
    """The CPython host shim of the FWHT call (csrc_host/fastcall.c: no CUDA, no torch headers; gcc, one second).  Optional:
    without it whvi_b200.fwht_ reaches the same C-ABI entry points through ctypes, ~1 us per call slower."""
    import sysconfig
    src = PKG / "csrc_host" / "fastcall.c"
    out = PKG / ("_fastcall" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))
    inc = sysconfig.get_paths()["include"]
    if not (Path(inc) / "Python.h").exists():
        return None
    if force or _stale(out, [src]):
        r = subprocess.run([os.environ.get("CC", "gcc"), "-O2", "-fPIC", "-shared", "-I", inc, str(src), "-o", str(out)],
                           capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            return None
    return out


if __name__ == "__main__":
    _variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    _defines = tuple(a for a in sys.argv[1:] if a.startswith("-D"))
    if _defines and not _variant:
        sys.exit("extra -D flags need --variant NAME (the product library is always built from the plain sources)")
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, variant=_variant, defines=_defines))
