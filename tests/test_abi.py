"""The C-ABI library loads on a CPU-only box and exports exactly what include/pbrt_b200.h declares."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "pbrt_b200.h"


def header_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(pbrt_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound(pb):
    from pbrt_b200 import _lib

    syms = header_symbols()
    assert len(syms) >= 40
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [s for s in syms if s not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    unbound = [s for s in syms if s not in _lib.PROTOTYPES]
    assert not unbound, f"declared in the header but not bound in _lib.py: {unbound}"
    extra = [s for s in _lib.PROTOTYPES if s not in syms]
    assert not extra, f"bound in _lib.py but not declared in the header: {extra}"


def test_library_is_sm100a_only(pb):
    from pbrt_b200 import _lib

    out = subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_exact_splat_kernel_has_no_fused_accumulate(pb):
    """ptxas fuses f32x2 mul+add into one FFMA2 even with --fmad=false (one rounding instead of two).
    The exact kernel therefore forms its products with scalar multiplies and only the sums are packed:
    it must contain packed adds, and any FFMA2 left in it must take a scalar ".F32" register addend
    (the -0 trick), never an accumulator."""
    from pbrt_b200 import _lib

    sass = subprocess.run(["cuobjdump", "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", sass)
    checked = 0
    for b in blocks:
        name = b.split("\n", 1)[0]
        if "splat_window_kernel" not in name:
            continue
        ffma2 = re.findall(r"FFMA2 [^;]*;", b)
        fadd2 = re.findall(r"FADD2 [^;]*;", b)
        if "Lb0E" in name:  # exact
            assert fadd2 and len(fadd2) >= len(ffma2), name
            assert len(re.findall(r"FMUL [^;]*;", b)) >= len(fadd2), name
            for ins in ffma2:
                assert re.search(r", U?R\d+\.F32 ;$", ins), f"{name}: fused accumulate {ins}"
            checked += 1
        else:  # fma mode accumulates with (scalar) fused multiply-adds: many more FFMA than the exact twin's divisions
            assert len(re.findall(r"FFMA R", b)) > 100 and not fadd2, name
    assert checked >= 4


def test_no_device_is_a_loud_error(pb):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pb.PbrtError) as e:
        pb.init(0)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)
    with pytest.raises(pb.PbrtError):
        pb.ConstantTexture(10.0).evaluate_batch(4)


def test_version_and_last_error(pb):
    from pbrt_b200 import _lib

    assert _lib.lib.pbrt_b200_version() >= 100
    assert isinstance(_lib.last_error(), str)


# ---- host-side logic of the library that needs no device: the Filter objects

def test_box_filter_matches_reference_doctest(pb, kats):
    k = kats["box_filter_xwidth_1"]
    f = pb.BoxFilter.create_box_filter(k["params"])
    assert list(f.radius()) == k["radius"] and list(f.inv_radius()) == k["inv_radius"]
    assert f.evaluate((0.25, 0.1)) == 1.0
    f2 = pb.make_filter("box", k["params"])
    assert list(f2.radius()) == k["radius"]
    d = pb.BoxFilter.create_box_filter()
    assert d.radius() == (0.5, 0.5)
    with pytest.raises(ValueError):
        pb.make_filter("nope")


@pytest.mark.parametrize("name", list(oracle.FILTERS))
def test_filter_tables_match_oracle_bit_exact(pb, orc, name):
    kind, radius, p0, p1 = oracle.FILTERS[name]
    want = oracle.filter_table(orc, kind, radius, p0, p1)
    cls = {"box": pb.BoxFilter, "triangle": pb.TriangleFilter, "gaussian": pb.GaussianFilter,
           "mitchell": pb.MitchellFilter, "lanczos": pb.LanczosSincFilter}[name]
    f = cls(radius) if name in ("box", "triangle") else (cls(radius, p0) if name != "mitchell" else cls(radius, p0, p1))
    got = pb.filter_table(f)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))

    # the generic path of Film::new (256 evaluate calls through the trait) gives the same table
    class Wrapped(pb.Filter):
        def evaluate(self, p):
            return f.evaluate(p)

        def radius(self):
            return f.radius()

        def inv_radius(self):
            return f.inv_radius()

    assert np.array_equal(pb.filter_table(Wrapped()).view(np.uint32), want.view(np.uint32))


def test_user_defined_filter_table(pb):
    class Tent(pb.Filter):
        def evaluate(self, p):
            return max(0.0, 1.0 - abs(p[0])) * max(0.0, 1.0 - abs(p[1]))

        def radius(self):
            return (1.0, 1.0)

        def inv_radius(self):
            return (1.0, 1.0)

    t = pb.filter_table(Tent()).reshape(16, 16)
    assert t[0, 0] == np.float32((1 - 0.5 / 16) ** 2) or abs(t[0, 0] - (1 - 0.5 / 16) ** 2) < 1e-7
    assert np.allclose(t, t.T)


def test_argument_errors_need_no_device(pb):
    """Bad arguments are rejected before any CUDA call, with a message."""
    import ctypes as C

    from pbrt_b200 import _lib

    L = _lib.lib
    assert L.pbrt_film_merge_tile(None, _lib.i32x4((0, 0, 1, 1)), None, 0) == _lib.E_INVALID
    assert "null" in _lib.last_error()
    assert L.pbrt_film_check(None) == _lib.E_INVALID
    assert L.pbrt_film_get_sample_bounds(None, _lib.i32x4((0, 0, 0, 0))) == _lib.E_INVALID
    assert L.pbrt_film_add_samples_tile(None, _lib.i32x4((0, 0, 1, 1)), 1, None, None, 0, 0) == _lib.E_INVALID
    assert L.pbrt_film_add_samples_tile_rgb(None, _lib.i32x4((0, 0, 1, 1)), 1, None, None, None, 0, 0) == _lib.E_INVALID
    assert "null" in _lib.last_error()
    assert L.pbrt_film_destroy(None) == _lib.OK  # Drop of nothing
    h = C.c_void_p()
    assert L.pbrt_filter_create(99, 1.0, 1.0, 0.0, 0.0, C.byref(h)) == _lib.E_INVALID
    assert "unknown filter kind" in _lib.last_error()
    assert L.pbrt_filter_table(None, None) == _lib.E_INVALID


def test_header_is_valid_c_and_cpp(tmp_path):
    """include/pbrt_b200.h is a C header (the FFI boundary): it must compile as C11 and as C++17."""
    c = tmp_path / "t.c"
    c.write_text('#include "pbrt_b200.h"\nint main(void) { int (*f)(void) = pbrt_b200_version; return f != 0 ? 0 : 1; }\n')
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-fsyntax-only", f"-I{ROOT / 'include'}", str(c)], check=True)
    cpp = tmp_path / "t.cpp"
    cpp.write_text('#include "pbrt_b200.hpp"\nint main() { pbrt::Bounds2i b = pbrt::Bounds2i::from({0, 0}, {2, 2}); return b.area() == 4 ? 0 : 1; }\n')
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", f"-I{ROOT / 'include'}", str(cpp)], check=True)


def test_rust_ffi_declares_every_header_symbol():
    """rust/src/ffi.rs is uncompiled source (no rustc here); at least keep it complete: one `fn` per header symbol."""
    header = (ROOT / "include" / "pbrt_b200.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = set(re.findall(r"\b(pbrt_[a-z0-9_]+)\s*\(", header))
    ffi = (ROOT / "rust" / "src" / "ffi.rs").read_text()
    missing = sorted(n for n in names if not re.search(r"\bfn " + n + r"\(", ffi))
    assert not missing, missing


def test_splat_window_kernels_do_not_spill_and_keep_their_occupancy(pb):
    """Every splat_window_kernel variant must stay spill-free; the 128-column variants of h <= 3 must stay within
    the 128 registers that four resident CTAs per SM allow (shared memory admits four: DESIGN.md section 5)."""
    from pbrt_b200 import _lib

    out = subprocess.run(["cuobjdump", "--dump-resource-usage", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    found = 0
    for name, regs, local in re.findall(r"Function (\S*splat_window_kernel\S*):\s*\n\s*REG:(\d+) .*?LOCAL:(\d+)", out):
        found += 1
        assert int(local) == 0, (name, "spills to local memory")
        m = re.search(r"ILi(\d)ELi(\d+)ELb", name)
        if m and int(m.group(2)) == 128 and int(m.group(1)) <= 3:
            assert int(regs) <= 128, (name, regs)
    assert found >= 16


def test_splat_class_kernels_keep_four_ctas_per_sm(pb):
    """The hot kernel: every 128-column splat_class_kernel variant must stay within the 128 registers that four resident
    CTAs per SM allow, and the radius-2 variants (BASELINE configs[1] and [2]) must not spill at all."""
    from pbrt_b200 import _lib

    out = subprocess.run(["cuobjdump", "--dump-resource-usage", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    found = 0
    for name, regs, local in re.findall(r"Function (\S*splat_class_kernel\S*):\s*\n\s*REG:(\d+) .*?LOCAL:(\d+)", out):
        m = re.search(r"ILi(\d)ELi(\d+)ELb", name)
        if not m or int(m.group(2)) != 128:
            continue
        found += 1
        assert int(regs) <= 128, (name, regs)
        if int(m.group(1)) == 2:
            assert int(local) == 0, (name, "spills to local memory")
    assert found >= 4
