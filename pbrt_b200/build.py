"""Build libpbrt_b200.so (CUDA, sm_100a only) in-tree with nvcc.

The library is the product; there is no other backend and no CPU fallback.  nvcc cross-compiles
without a GPU, so this also runs on the CPU-only build container.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "lib"
LIB = LIB_DIR / "libpbrt_b200.so"
SOURCES = ["film.cu", "splat.cu", "splat_class.cu"]
HEADERS = [CSRC / "common.cuh", CSRC / "to_byte_table.inc", ROOT / "include" / "pbrt_b200.h"]

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    # The reference is Rust, which never contracts a*b+c; kernels that want FMA ask for it.
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O2",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libpbrt_b200 cannot be built")


def _stale(out: Path, deps: list[Path]) -> bool:
    if not out.exists():
        return True
    t = out.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False, variant: str = "", defines: list[str] | None = None) -> Path:
    """Compile every .cu for sm_100a and link the shared library. Returns its path.

    `variant` / `defines` build an experiment copy (lib/libpbrt_b200_<variant>.so with extra -D macros) beside the
    product library; PBRT_B200_LIB selects which one pbrt_b200/_lib.py loads."""
    LIB_DIR.mkdir(exist_ok=True)
    objdir = PKG / "build" / variant if variant else PKG / "build"
    objdir.mkdir(parents=True, exist_ok=True)
    nvcc = _nvcc()
    lib = LIB_DIR / f"libpbrt_b200_{variant}.so" if variant else LIB
    flags = NVCC_FLAGS + [f"-D{d}" for d in (defines or [])]
    objs = []
    for src in SOURCES:
        s = CSRC / src
        o = objdir / (s.stem + ".o")
        if force or _stale(o, [s, *HEADERS, Path(__file__)]):
            cmd = [nvcc, *flags, "-c", str(s), "-o", str(o)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError(f"nvcc failed on {src}")
        objs.append(o)
    if force or _stale(lib, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(lib), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return lib


if __name__ == "__main__":
    # python pbrt_b200/build.py [--force] [-v] [--variant NAME -DMACRO=V ...]
    _variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else ""
    _defs = [a[2:] for a in sys.argv if a.startswith("-D")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=_variant, defines=_defs))
