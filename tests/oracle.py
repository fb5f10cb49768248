"""ctypes wrapper of the CPU oracle (oracle/pbrt_oracle.c). Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
ORACLE_LIB = ORACLE_DIR / "_build" / "libpbrt_oracle.so"


class B2i(C.Structure):
    _fields_ = [("x0", C.c_int64), ("y0", C.c_int64), ("x1", C.c_int64), ("y1", C.c_int64)]

    def t(self):
        return (self.x0, self.y0, self.x1, self.y1)


class B2f(C.Structure):
    _fields_ = [("x0", C.c_float), ("y0", C.c_float), ("x1", C.c_float), ("y1", C.c_float)]

    def t(self):
        return (self.x0, self.y0, self.x1, self.y1)


class OFilter(C.Structure):
    _fields_ = [("kind", C.c_int), ("radius", C.c_float * 2), ("inv_radius", C.c_float * 2), ("p0", C.c_float),
                ("p1", C.c_float), ("exp_x", C.c_float), ("exp_y", C.c_float)]


class ORng(C.Structure):
    _fields_ = [("state", C.c_uint64), ("inc", C.c_uint64)]


def build() -> Path:
    src = [ORACLE_DIR / "pbrt_oracle.c", ORACLE_DIR / "pbrt_oracle.h", ORACLE_DIR / "Makefile"]
    if not ORACLE_LIB.exists() or any(s.stat().st_mtime > ORACLE_LIB.stat().st_mtime for s in src):
        subprocess.run(["make", "-C", str(ORACLE_DIR)], check=True, capture_output=True)
    return ORACLE_LIB


_f32p = C.POINTER(C.c_float)
_vp = C.c_void_p


def load() -> C.CDLL:
    o = C.CDLL(str(build()))
    sig = {
        "orc_gamma_correct": (C.c_float, [C.c_float]),
        "orc_clamp_f": (C.c_float, [C.c_float] * 3),
        "orc_clamp_i": (C.c_int64, [C.c_int64] * 3),
        "orc_to_byte": (C.c_uint8, [C.c_float]),
        "orc_f2i": (C.c_int64, [C.c_float]),
        "orc_bounds2i_from_points": (B2i, [C.c_int64] * 4),
        "orc_bounds2i_intersect": (B2i, [B2i, B2i]),
        "orc_bounds2i_area": (C.c_int64, [B2i]),
        "orc_bounds2i_inside_exclusive": (C.c_int, [B2i, C.c_int64, C.c_int64]),
        "orc_bounds2i_iter": (C.c_int64, [B2i, C.POINTER(C.c_int64), C.c_int64]),
        "orc_point2f_floor": (None, [_f32p, _f32p]),
        "orc_point2f_ceil": (None, [_f32p, _f32p]),
        "orc_rgb_to_xyz": (None, [_f32p, _f32p]),
        "orc_xyz_to_rgb": (None, [_f32p, _f32p]),
        "orc_filter_init": (None, [C.POINTER(OFilter), C.c_int, C.c_float, C.c_float, C.c_float, C.c_float]),
        "orc_box_filter_create": (None, [C.POINTER(OFilter), C.c_int, C.c_float, C.c_int, C.c_float]),
        "orc_filter_evaluate": (C.c_float, [C.POINTER(OFilter), C.c_float, C.c_float]),
        "orc_filter_table": (None, [C.POINTER(OFilter), _f32p]),
        "orc_film_new": (_vp, [C.c_int64, C.c_int64, _f32p, _f32p, _f32p, C.c_float, C.c_float, C.c_float]),
        "orc_film_free": (None, [_vp]),
        "orc_film_cropped_pixel_bounds": (B2i, [_vp]),
        "orc_film_get_sample_bounds": (B2i, [_vp]),
        "orc_film_get_physical_extent": (B2f, [_vp]),
        "orc_film_tile_bounds": (B2i, [_vp, B2i]),
        "orc_film_get_film_tile": (_vp, [_vp, B2i]),
        "orc_film_merge_film_tile": (None, [_vp, _vp]),
        "orc_film_write_image_rgb": (None, [_vp, C.c_float, _f32p]),
        "orc_film_get_pixel_xyz": (None, [_vp, C.c_int64, C.c_int64, _f32p]),
        "orc_film_pixels": (_vp, [_vp]),
        "orc_film_pixel_count": (C.c_int64, [_vp]),
        "orc_film_table": (_f32p, [_vp]),
        "orc_tile_free": (None, [_vp]),
        "orc_tile_get_pixel_bounds": (B2i, [_vp]),
        "orc_tile_pixel_count": (C.c_int64, [_vp]),
        "orc_tile_pixels": (_vp, [_vp]),
        "orc_tile_get_pixel": (_vp, [_vp, C.c_int64, C.c_int64]),
        "orc_constant_texture_eval_f32": (None, [C.c_int, C.c_float, C.c_uint64, _f32p]),
        "orc_constant_texture_eval_rgb": (None, [C.c_int, _f32p, C.c_uint64, _f32p]),
        "orc_weight_lut": (None, [_f32p]),
        "orc_rng_default": (None, [C.POINTER(ORng)]),
        "orc_rng_set_sequence": (None, [C.POINTER(ORng), C.c_uint64]),
        "orc_rng_uniform_u32": (C.c_uint32, [C.POINTER(ORng)]),
        "orc_rng_uniform_u32_threshold": (C.c_uint32, [C.POINTER(ORng), C.c_uint32]),
        "orc_rng_uniform_float": (C.c_float, [C.POINTER(ORng)]),
        "orc_pfm_encode": (C.c_size_t, [_f32p, C.c_int64, C.c_int64, C.POINTER(C.c_uint8), C.c_size_t]),
        "orc_ext_tile_add_sample": (None, [_vp, C.c_float, C.c_float, _f32p, C.c_float]),
        "orc_ext_tile_add_samples": (None, [_vp, C.c_uint64, _f32p, _f32p]),
        "orc_ext_film_add_splat": (None, [_vp, C.c_float, C.c_float, _f32p]),
        "orc_ext_synth_samples": (None, [B2i, C.c_int, C.c_uint64, _f32p, _f32p]),
        "orc_ext_synth_tile_fill": (None, [_vp, C.c_uint64, C.c_uint64]),
        "orc_ext_film_add_samples_pass": (None, [_vp, B2i, C.c_int, _f32p, _f32p, C.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(o, name)
        fn.restype = res
        fn.argtypes = args
    return o


def fp(a: np.ndarray):
    return a.ctypes.data_as(_f32p)


def farr(v):
    v = list(v)
    return (C.c_float * len(v))(*[float(x) for x in v])


class OracleFilm:
    """Convenience wrapper mirroring the reference's Film over the oracle."""

    def __init__(self, o, res, crop, radius, table, diagonal_mm=35.0, scale=1.0, max_lum=float("inf")):
        self.o = o
        self.table = np.ascontiguousarray(table, dtype=np.float32)
        self.h = o.orc_film_new(int(res[0]), int(res[1]), farr(crop), farr(radius), fp(self.table), diagonal_mm, scale, max_lum)

    def __del__(self):
        if getattr(self, "h", None):
            self.o.orc_film_free(self.h)
            self.h = None

    def cropped(self):
        return self.o.orc_film_cropped_pixel_bounds(self.h).t()

    def sample_bounds(self):
        return self.o.orc_film_get_sample_bounds(self.h).t()

    def physical_extent(self):
        return self.o.orc_film_get_physical_extent(self.h).t()

    def tile_bounds(self, sb):
        return self.o.orc_film_tile_bounds(self.h, B2i(*sb)).t()

    def get_film_tile(self, sb):
        return self.o.orc_film_get_film_tile(self.h, B2i(*sb))

    def tile_pixels(self, tile) -> np.ndarray:
        n = self.o.orc_tile_pixel_count(tile)
        p = self.o.orc_tile_pixels(tile)
        return np.ctypeslib.as_array(C.cast(p, _f32p), shape=(max(n, 1) * 4,))[: n * 4].reshape(n, 4)

    def merge(self, tile):
        self.o.orc_film_merge_film_tile(self.h, tile)

    def pixels(self) -> np.ndarray:
        n = self.o.orc_film_pixel_count(self.h)
        p = self.o.orc_film_pixels(self.h)
        return np.ctypeslib.as_array(C.cast(p, _f32p), shape=(max(n, 1) * 7,))[: n * 7].reshape(n, 7).copy()

    def write_image_rgb(self, splat_scale=1.0) -> np.ndarray:
        n = self.o.orc_film_pixel_count(self.h)
        out = np.zeros((n, 3), dtype=np.float32)
        if n:
            self.o.orc_film_write_image_rgb(self.h, splat_scale, fp(out))
        return out

    def get_pixel_xyz(self, x, y):
        out = (C.c_float * 3)()
        self.o.orc_film_get_pixel_xyz(self.h, x, y, out)
        return (out[0], out[1], out[2])

    def add_samples_pass(self, sb, spp, xy, rgbw, threads=1):
        xy = np.ascontiguousarray(xy, dtype=np.float32)
        rgbw = np.ascontiguousarray(rgbw, dtype=np.float32)
        self.o.orc_ext_film_add_samples_pass(self.h, B2i(*sb), spp, fp(xy), fp(rgbw), threads)

    def add_splat(self, x, y, v):
        self.o.orc_ext_film_add_splat(self.h, x, y, farr(v))


def filter_table(o, kind, radius, p0=0.0, p1=0.0) -> np.ndarray:
    f = OFilter()
    o.orc_filter_init(C.byref(f), kind, radius[0], radius[1], p0, p1)
    t = np.zeros(256, dtype=np.float32)
    o.orc_filter_table(C.byref(f), fp(t))
    return t


def synth_samples(o, bounds, spp, seed=1):
    b = B2i(*bounds)
    n = max((b.x1 - b.x0) * (b.y1 - b.y0), 0) * spp
    xy = np.zeros((n, 2), dtype=np.float32)
    rgbw = np.zeros((n, 4), dtype=np.float32)
    if n:
        o.orc_ext_synth_samples(b, spp, seed, fp(xy), fp(rgbw))
    return xy, rgbw


# filter kinds / pbrt-v3 defaults shared by tests and bench
FILTERS = {
    "box": (0, (0.5, 0.5), 0.0, 0.0),
    "triangle": (1, (2.0, 2.0), 0.0, 0.0),
    "gaussian": (2, (2.0, 2.0), 2.0, 0.0),
    "mitchell": (3, (2.0, 2.0), 1.0 / 3.0, 1.0 / 3.0),
    "lanczos": (4, (4.0, 4.0), 3.0, 0.0),
}
