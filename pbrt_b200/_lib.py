"""ctypes binding of libpbrt_b200.so — one prototype per symbol declared in include/pbrt_b200.h.

Importing this module only loads the library (no CUDA call is made), so it works on a CPU-only
box; any compute entry point fails with PBRT_E_CUDA there.  There is no fallback of any kind:
if the shared library is missing the import raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("PBRT_B200_LIB", PKG / "lib" / "libpbrt_b200.so"))

OK, E_INVALID, E_CUDA, E_RANGE, E_NOT_PIXEL_MAJOR, E_UNSUPPORTED, E_NOMEM, E_NONFINITE = range(8)
FILTER_BOX, FILTER_TRIANGLE, FILTER_GAUSSIAN, FILTER_MITCHELL, FILTER_LANCZOS = range(5)
SPLAT_EXACT, SPLAT_FMA, SPLAT_ATOMIC = range(3)


class PbrtError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pbrt_b200 error {code}: {msg}")
        self.code = code


if not LIB_PATH.exists():
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python build_native.py` "
        "(libpbrt_b200 is CUDA-only; there is no CPU fallback)"
    )

lib = C.CDLL(str(LIB_PATH))

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f32p = C.POINTER(C.c_float)
_vp = C.c_void_p
_vpp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); the list is checked against the header in tests/test_abi.py
PROTOTYPES = {
    "pbrt_b200_version": (C.c_int, []),
    "pbrt_b200_last_error": (C.c_char_p, []),
    "pbrt_b200_init": (C.c_int, [C.c_int]),
    "pbrt_b200_set_stream": (C.c_int, [_vp]),
    "pbrt_b200_synchronize": (C.c_int, []),
    "pbrt_b200_device_info": (C.c_int, [_i32p, _i32p, _i32p, _i32p, C.POINTER(C.c_uint64)]),
    "pbrt_b200_launch_count": (C.c_uint64, []),
    "pbrt_b200_overlap_passes": (C.c_int, [C.c_int]),
    "pbrt_b200_malloc": (C.c_int, [C.c_uint64, _vpp]),
    "pbrt_b200_free": (C.c_int, [_vp]),
    "pbrt_b200_host_alloc": (C.c_int, [C.c_uint64, _vpp]),
    "pbrt_b200_host_free": (C.c_int, [_vp]),
    "pbrt_b200_memcpy_h2d": (C.c_int, [_vp, _vp, C.c_uint64]),
    "pbrt_b200_memcpy_d2h": (C.c_int, [_vp, _vp, C.c_uint64]),
    "pbrt_b200_memset": (C.c_int, [_vp, C.c_int, C.c_uint64]),
    "pbrt_b200_ipc_export": (C.c_int, [_vp, C.POINTER(C.c_uint8)]),
    "pbrt_b200_ipc_import": (C.c_int, [C.POINTER(C.c_uint8), _vpp]),
    "pbrt_b200_ipc_close": (C.c_int, [_vp]),
    "pbrt_filter_create": (C.c_int, [C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, _vpp]),
    "pbrt_box_filter_create_from_params": (C.c_int, [C.c_int, C.c_float, C.c_int, C.c_float, _vpp]),
    "pbrt_filter_destroy": (None, [_vp]),
    "pbrt_filter_evaluate": (C.c_float, [_vp, C.c_float, C.c_float]),
    "pbrt_filter_radius": (None, [_vp, _f32p]),
    "pbrt_filter_inv_radius": (None, [_vp, _f32p]),
    "pbrt_filter_table": (C.c_int, [_vp, _f32p]),
    "pbrt_film_create": (C.c_int, [C.c_int32, C.c_int32, _f32p, _f32p, _f32p, C.c_float, C.c_float, C.c_float, _vpp]),
    "pbrt_film_create_sharded": (
        C.c_int,
        [C.c_int32, C.c_int32, _f32p, _f32p, _f32p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, _vpp],
    ),
    "pbrt_film_destroy": (C.c_int, [_vp]),
    "pbrt_film_cropped_pixel_bounds": (C.c_int, [_vp, _i32p]),
    "pbrt_film_owned_pixel_bounds": (C.c_int, [_vp, _i32p]),
    "pbrt_film_get_sample_bounds": (C.c_int, [_vp, _i32p]),
    "pbrt_film_get_physical_extent": (C.c_int, [_vp, _f32p]),
    "pbrt_film_tile_bounds": (C.c_int, [_vp, _i32p, _i32p, _i64p]),
    "pbrt_film_geometry": (C.c_int, [C.c_int32, C.c_int32, _f32p, _f32p, C.c_float, C.c_int, C.c_int, _i32p, _i32p, _i32p, _f32p]),
    "pbrt_film_geometry_tile_bounds": (C.c_int, [_i32p, _f32p, _i32p, _i32p, _i64p]),
    "pbrt_film_route_plan": (C.c_int, [_i32p, _i32p, _f32p, C.c_int32, _i32p, _i32p]),
    "pbrt_film_merge_tile": (C.c_int, [_vp, _i32p, _vp, C.c_int]),
    "pbrt_film_merge_tiles": (C.c_int, [_vp, C.c_int32, _i32p, _i64p, _vp, C.c_int64, C.c_int]),
    "pbrt_film_add_samples_tile": (C.c_int, [_vp, _i32p, C.c_int32, _vp, _vp, C.c_int, C.c_int]),
    "pbrt_film_add_samples_tile_rgb": (C.c_int, [_vp, _i32p, C.c_int32, _vp, _vp, _vp, C.c_int, C.c_int]),
    "pbrt_film_add_samples_tiles": (C.c_int, [_vp, C.c_int32, _i32p, _i64p, C.c_int32, _vp, _vp, C.c_int64, C.c_int, C.c_int]),
    "pbrt_film_add_samples": (C.c_int, [_vp, _i32p, C.c_uint64, _vp, _vp, C.c_int]),
    "pbrt_film_add_splats": (C.c_int, [_vp, C.c_uint64, _vp, _vp, C.c_int]),
    "pbrt_film_set_image": (C.c_int, [_vp, _vp, C.c_int]),
    "pbrt_film_clear": (C.c_int, [_vp]),
    "pbrt_film_resolve_rgb": (C.c_int, [_vp, C.c_float, _vp, C.c_int]),
    "pbrt_film_resolve_rgb8": (C.c_int, [_vp, C.c_float, _vp, C.c_int]),
    "pbrt_film_get_pixel_xyz": (C.c_int, [_vp, C.c_int32, C.c_int32, _f32p]),
    "pbrt_film_read_pixels": (C.c_int, [_vp, _vp, C.c_int]),
    "pbrt_film_device_buffers": (C.c_int, [_vp, _vpp, _vpp, _i64p]),
    "pbrt_film_resolve_rgb_to_frames": (C.c_int, [_vp, C.c_float, C.c_int32, _vpp]),
    "pbrt_film_check": (C.c_int, [_vp]),
    "pbrt_texture_constant_eval_f32": (C.c_int, [C.c_float, C.c_uint64, _vp, C.c_int]),
    "pbrt_texture_constant_eval_rgb": (C.c_int, [_f32p, C.c_uint64, _vp, C.c_int]),
    "pbrt_mipmap_weight_lut": (C.c_int, [_f32p]),
    "pbrt_synth_samples": (C.c_int, [_i32p, _i32p, C.c_int32, C.c_uint64, _vp, _vp]),
    "pbrt_synth_tiles": (C.c_int, [C.c_int32, _i64p, _i64p, C.c_uint64, _vp, C.c_int64]),
}

for _name, (_res, _args) in PROTOTYPES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header and library disagree
    _fn.restype = _res
    _fn.argtypes = _args

# test hook, not part of the public header
lib.pbrt_b200_debug_force_generic_splat.restype = C.c_int
lib.pbrt_b200_debug_force_generic_splat.argtypes = [C.c_int]
lib.pbrt_b200_debug_class_tables.restype = C.c_int
lib.pbrt_b200_debug_class_tables.argtypes = [_f32p, C.c_float, C.c_float, C.POINTER(C.c_uint8), C.c_int, _i32p]
lib.pbrt_b200_debug_class_segments.restype = C.c_int
lib.pbrt_b200_debug_class_segments.argtypes = [C.c_int] * 7 + [_i32p]


def last_error() -> str:
    return (lib.pbrt_b200_last_error() or b"").decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != OK:
        raise PbrtError(rc, last_error())


def i32x4(v) -> C.Array:
    return (C.c_int32 * 4)(*[int(x) for x in v])


def f32arr(v) -> C.Array:
    v = list(v)
    return (C.c_float * len(v))(*[float(x) for x in v])
