"""The reference's film / filter / texture tests restated in C++ (tests/cpp/test_film.cpp) over
include/pbrt_b200.hpp -> C ABI -> CUDA."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "tests" / "cpp" / "build" / "test_film"


def build_cpp_test() -> Path:
    import sys

    sys.path.insert(0, str(ROOT))
    import build_native

    lib = build_native.build()
    src = ROOT / "tests" / "cpp" / "test_film.cpp"
    deps = [src, ROOT / "include" / "pbrt_b200.hpp", ROOT / "include" / "pbrt_b200.h", lib]
    if not BIN.exists() or any(d.stat().st_mtime > BIN.stat().st_mtime for d in deps):
        BIN.parent.mkdir(exist_ok=True)
        subprocess.run(
            ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-o", str(BIN), str(src), f"-L{lib.parent}", "-lpbrt_b200",
             f"-Wl,-rpath,{lib.parent}"],
            check=True, capture_output=True, text=True,
        )
    return BIN


def test_cpp_mirror_compiles_and_links():
    """CPU box: the C++ mirror of the reference interfaces compiles against the C ABI and links."""
    assert build_cpp_test().exists()


@pytest.mark.gpu
def test_reference_tests_in_cpp():
    r = subprocess.run([str(build_cpp_test())], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all reference tests passed" in r.stdout
