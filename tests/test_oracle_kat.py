"""Pin the CPU oracle against every known answer the reference's own tests hold for this path.

Each test cites the reference assertion it reproduces (tests/golden/reference_kats.json carries
the same citations).  No GPU needed.
"""
import ctypes as C
import struct

import numpy as np
import pytest

import oracle
from oracle import B2i, OFilter, ORng, OracleFilm, farr


def bits(x: float) -> int:
    return struct.unpack("<I", struct.pack("<f", x))[0]


def box_table():
    return np.ones(256, dtype=np.float32)


# ---------------------------------------------------------------- geometry / prelude

def test_bounds2i_inside_intersect_iter(orc, kats):
    k = kats["bounds2i"]
    b = B2i(*k["inside"]["bounds"])
    assert orc.orc_bounds2i_inside_exclusive(b, *k["inside"]["in"]) == 1
    assert orc.orc_bounds2i_inside_exclusive(b, *k["inside"]["out"]) == 0
    r = orc.orc_bounds2i_intersect(B2i(*k["intersect"]["a"]), B2i(*k["intersect"]["b"]))
    assert list(r.t()) == k["intersect"]["result"]
    r = orc.orc_bounds2i_intersect(B2i(*k["intersect_disjoint"]["a"]), B2i(*k["intersect_disjoint"]["b"]))
    assert list(r.t()) == k["intersect_disjoint"]["result"]  # inverted, not re-sorted (bounds.rs:236-241)
    out = (C.c_int64 * 16)()
    n = orc.orc_bounds2i_iter(B2i(*k["iter"]["bounds"]), out, 8)
    assert n == 4 and [[out[2 * i], out[2 * i + 1]] for i in range(4)] == k["iter"]["points"]
    r = orc.orc_bounds2i_from_points(*k["from_unsorted"]["points"])
    assert list(r.t()) == k["from_unsorted"]["result"]
    # area of a doubly inverted box is positive (bounds.rs:195-198) and iter yields nothing
    inv = B2i(3, 3, 2, 2)
    assert orc.orc_bounds2i_area(inv) == 1
    assert orc.orc_bounds2i_iter(inv, out, 8) == 0


def test_point2f_floor_ceil(orc, kats):
    k = kats["point2f"]
    out = (C.c_float * 2)()
    orc.orc_point2f_floor(farr(k["p"]), out)
    assert list(out) == k["floor"]
    orc.orc_point2f_ceil(farr(k["p"]), out)
    assert list(out) == k["ceil"]


def test_clamp(orc, kats):
    for v, lo, hi, want in kats["clamp"]["float"]:
        assert orc.orc_clamp_f(v, lo, hi) == want
    for v, lo, hi, want in kats["clamp"]["int"]:
        assert orc.orc_clamp_i(v, lo, hi) == want


def test_f2i_saturates_like_rust(orc):
    assert orc.orc_f2i(float("nan")) == 0
    assert orc.orc_f2i(1e30) == 2**63 - 1
    assert orc.orc_f2i(-1e30) == -(2**63)
    assert orc.orc_f2i(-1.9) == -1 and orc.orc_f2i(1.9) == 1


# ---------------------------------------------------------------- spectrum

def test_rgb_to_xyz_bit_patterns(orc):
    # derived from spectrum.rs:141-143 for the colours of film.rs:533-534
    out = (C.c_float * 3)()
    orc.orc_rgb_to_xyz(farr([0, 1, 0]), out)
    assert [bits(v) for v in out] == [0x3EB714BA, 0x3F3714BA, 0x3DF41B76]
    orc.orc_rgb_to_xyz(farr([1, 0, 0]), out)
    assert [bits(v) for v in out] == [0x3ED32D0A, 0x3E59C66D, 0x3C9E6256]


def test_xyz_rgb_roundtrip_close(orc):
    rgb = farr([0.2, 0.5, 0.8])
    xyz, back = (C.c_float * 3)(), (C.c_float * 3)()
    orc.orc_rgb_to_xyz(rgb, xyz)
    orc.orc_xyz_to_rgb(xyz, back)
    assert np.allclose(list(back), [0.2, 0.5, 0.8], atol=1e-5)


# ---------------------------------------------------------------- filters

def test_box_filter_from_params(orc, kats):
    k = kats["box_filter_xwidth_1"]
    f = OFilter()
    orc.orc_box_filter_create(C.byref(f), 1, k["params"]["xwidth"], 0, 0.0)
    assert list(f.radius) == k["radius"] and list(f.inv_radius) == k["inv_radius"]
    assert orc.orc_filter_evaluate(C.byref(f), 0.3, -0.2) == 1.0
    t = np.zeros(256, dtype=np.float32)
    orc.orc_filter_table(C.byref(f), oracle.fp(t))
    assert (t == 1.0).all()


def test_ext_filter_tables_are_sane(orc):
    # EXTENSION filters: no reference values exist; check the shape pbrt-v3 documents.
    for name, (kind, radius, p0, p1) in oracle.FILTERS.items():
        t = oracle.filter_table(orc, kind, radius, p0, p1).reshape(16, 16)
        assert np.isfinite(t).all()
        assert np.array_equal(t, t.T), name  # separable with equal radii
        assert t[0, 0] == t.max(), name
    g = oracle.filter_table(orc, 2, (2.0, 2.0), 2.0).reshape(16, 16)
    assert (np.diff(g[0]) < 0).all()
    m = oracle.filter_table(orc, 3, (2.0, 2.0), 1 / 3, 1 / 3).reshape(16, 16)
    assert m.min() < 0  # Mitchell's negative lobe


# ---------------------------------------------------------------- film

def test_film_sample_and_tile_bounds(orc, kats):
    k = kats["film_1920x1080_crop_quarter_box8"]
    f = OracleFilm(orc, k["resolution"], k["crop"], k["radius"], box_table())
    assert list(f.cropped()) == [480, 270, 1440, 810]
    assert list(f.sample_bounds()) == k["sample_bounds"]
    assert list(f.tile_bounds(k["tile_of_full_frame"]["sample_bounds"])) == k["tile_of_full_frame"]["pixel_bounds"]
    assert list(f.tile_bounds(k["tile_inside"]["sample_bounds"])) == k["tile_inside"]["pixel_bounds"]


def test_film_physical_extent(orc, kats):
    k = kats["film_800x600_physical_extent"]
    want = [float(np.float32(v)) for v in k["extent"]]
    for crop in k["crops"]:
        f = OracleFilm(orc, k["resolution"], crop, k["radius"], box_table(), diagonal_mm=k["diagonal_mm"])
        assert list(f.physical_extent()) == want  # the doctest asserts exact equality


def test_film_degenerate_tile_merges(orc, kats):
    k = kats["film_20x10_degenerate_tile"]
    f = OracleFilm(orc, k["resolution"], k["crop"], k["radius"], box_table())
    left, right = f.get_film_tile(k["left"]), f.get_film_tile(k["right"])
    assert list(orc.orc_tile_get_pixel_bounds(right).t()) == [2, 0, 18, 10]
    f.merge(left)
    f.merge(right)
    assert (f.pixels() == 0).all()


def _fill(of, tile, rgb):
    px = of.tile_pixels(tile)
    px[:, :3] = rgb
    px[:, 3] = 1.0


def test_merge_film_tile_reference_test(orc, kats):
    k = kats["merge_film_tile_test"]
    f = OracleFilm(orc, k["resolution"], k["crop"], k["radius"], box_table(), scale=1.0, max_lum=1.0)
    left, right = f.get_film_tile(k["left"]), f.get_film_tile(k["right"])
    assert list(orc.orc_tile_get_pixel_bounds(left).t()) == [0, 0, 108, 10]
    assert list(orc.orc_tile_get_pixel_bounds(right).t()) == [92, 0, 200, 10]
    _fill(f, left, k["green"])
    _fill(f, right, k["red"])
    f.merge(left)
    f.merge(right)
    xyz_g, xyz_r = (C.c_float * 3)(), (C.c_float * 3)()
    orc.orc_rgb_to_xyz(farr(k["green"]), xyz_g)
    orc.orc_rgb_to_xyz(farr(k["red"]), xyz_r)
    assert f.get_pixel_xyz(4, 4) == tuple(xyz_g)      # film.rs:533
    assert f.get_pixel_xyz(196, 4) == tuple(xyz_r)    # film.rs:534
    # overlap pixel: SURVEY.md App. B
    assert [bits(v) for v in f.get_pixel_xyz(100, 4)] == [0x3F4520E2, 0x3F6D8655, 0x3E0DDA06]
    rgb = f.write_image_rgb(1.0).reshape(10, 200, 3)
    assert [orc.orc_to_byte(float(v)) for v in rgb[4, 4]] == [0, 255, 0]
    assert [orc.orc_to_byte(float(v)) for v in rgb[4, 196]] == [255, 0, 0]
    assert [orc.orc_to_byte(float(v)) for v in rgb[4, 100]] == [188, 188, 0]


def test_write_image_weight_zero_and_scale(orc):
    f = OracleFilm(orc, [4, 2], [0, 0, 1, 1], [0.5, 0.5], box_table(), scale=2.0)
    assert (f.write_image_rgb(1.0) == 0).all()  # w == 0 skips the normalisation (film.rs:355)
    t = f.get_film_tile([0, 0, 4, 2])
    px = f.tile_pixels(t)
    px[:, :3] = [0.25, 0.5, 0.75]
    px[:, 3] = 2.0
    f.merge(t)
    rgb = f.write_image_rgb(1.0)
    assert np.allclose(rgb, np.array([0.25, 0.5, 0.75]) / 2.0 * 2.0, atol=1e-6)


# ---------------------------------------------------------------- textures / LUT

def test_constant_texture(orc, kats):
    k = kats["constant_texture"]
    out = np.zeros(5, dtype=np.float32)
    orc.orc_constant_texture_eval_f32(1, k["float_value"], 5, oracle.fp(out))
    assert (out == 10.0).all()
    orc.orc_constant_texture_eval_f32(0, 0.0, 5, oracle.fp(out))
    assert (out == k["float_default"]).all()
    out3 = np.zeros((4, 3), dtype=np.float32)
    orc.orc_constant_texture_eval_rgb(1, farr(k["spectrum_value"]), 4, oracle.fp(out3))
    assert (out3 == np.array(k["spectrum_value"], dtype=np.float32)).all()
    orc.orc_constant_texture_eval_rgb(0, farr([0, 0, 0]), 4, oracle.fp(out3))
    assert (out3 == 1.0).all()


def test_weight_lut(orc):
    lut = np.zeros(128, dtype=np.float32)
    orc.orc_weight_lut(oracle.fp(lut))
    # SURVEY.md App. B (derived from mipmap.rs:45-51), +-1 ulp of libm
    want = {0: 0.86466473, 1: 0.84904003, 64: 0.22965886, 126: 0.0021481365, 127: 0.0}
    for i, v in want.items():
        assert abs(lut[i] - v) <= 2e-7 * max(1.0, abs(v))


# ---------------------------------------------------------------- rng

def test_pcg32_kats(orc, kats):
    k = kats["rng"]
    r = ORng()
    orc.orc_rng_default(C.byref(r))
    assert r.state == int(k["default_state"], 16) and r.inc == int(k["default_inc"], 16)
    assert [orc.orc_rng_uniform_u32(C.byref(r)) for _ in range(10)] == k["default_u32"]
    orc.orc_rng_default(C.byref(r))
    assert [orc.orc_rng_uniform_u32_threshold(C.byref(r), 4095) for _ in range(10)] == k["default_threshold_4095"]
    orc.orc_rng_default(C.byref(r))
    orc.orc_rng_set_sequence(C.byref(r), 0)
    assert orc.orc_rng_uniform_u32(C.byref(r)) == k["new0_first_u32"]
    orc.orc_rng_default(C.byref(r))
    got = [orc.orc_rng_uniform_float(C.byref(r)) for _ in range(10)]
    assert np.allclose(got, k["default_float"], atol=k["float_tolerance"], rtol=0)
    orc.orc_rng_default(C.byref(r))
    orc.orc_rng_uniform_u32_threshold(C.byref(r), 0xFFFFFFFF // 2)  # rng.rs:158-163: must terminate


# ---------------------------------------------------------------- imageio

def test_pfm_layout(orc):
    w, h = 3, 2
    rgb = np.arange(w * h * 3, dtype=np.float32)
    need = orc.orc_pfm_encode(oracle.fp(rgb), w, h, None, 0)
    buf = (C.c_uint8 * need)()
    orc.orc_pfm_encode(oracle.fp(rgb), w, h, buf, need)
    raw = bytes(buf)
    assert raw.startswith(b"PF\n3 2\n-1\n")
    body = np.frombuffer(raw[len(b"PF\n3 2\n-1\n"):], dtype="<f4").reshape(h, w * 3)
    assert np.array_equal(body[0], rgb.reshape(h, w * 3)[1])  # bottom row first (imageio.rs:198-209)


def test_to_byte(orc):
    assert orc.orc_to_byte(0.0) == 0 and orc.orc_to_byte(1.0) == 255 and orc.orc_to_byte(2.0) == 255
    assert orc.orc_to_byte(-1.0) == 0 and orc.orc_to_byte(float("nan")) == 0
    assert orc.orc_to_byte(0.5) == 188  # SURVEY.md App. B (merge test overlap pixel)


# ---------------------------------------------------------------- extension self-consistency

@pytest.mark.parametrize("name", list(oracle.FILTERS))
def test_ext_threaded_pass_equals_single_thread(orc, name):
    kind, radius, p0, p1 = oracle.FILTERS[name]
    table = oracle.filter_table(orc, kind, radius, p0, p1)
    res, spp = (40, 24), 4
    xy, rgbw = oracle.synth_samples(orc, (0, 0, *res), spp)
    a = OracleFilm(orc, res, [0, 0, 1, 1], radius, table)
    b = OracleFilm(orc, res, [0, 0, 1, 1], radius, table)
    a.add_samples_pass((0, 0, *res), spp, xy, rgbw, threads=1)
    b.add_samples_pass((0, 0, *res), spp, xy, rgbw, threads=5)
    assert np.array_equal(a.pixels().view(np.uint32), b.pixels().view(np.uint32))
    assert a.pixels()[:, 3].min() > 0 or name == "mitchell" or name == "lanczos"


def test_ext_add_sample_box_weight_counts(orc):
    # box filter r=.5: every sample lands on exactly its own pixel with weight 1
    res, spp = (8, 6), 4
    xy, rgbw = oracle.synth_samples(orc, (0, 0, *res), spp)
    f = OracleFilm(orc, res, [0, 0, 1, 1], (0.5, 0.5), box_table())
    f.add_samples_pass((0, 0, *res), spp, xy, rgbw)
    assert (f.pixels()[:, 3] == spp).all()


def test_ext_synth_samples_layout(orc):
    xy, rgbw = oracle.synth_samples(orc, (3, 5, 7, 8), 4)
    px = np.repeat(np.tile(np.arange(3, 7), 3), 4)
    py = np.repeat(np.repeat(np.arange(5, 8), 4), 4)
    assert (np.floor(xy[:, 0]) == px).all() and (np.floor(xy[:, 1]) == py).all()
    assert (rgbw[:, 3] == 1).all() and (rgbw[:, :3] >= 0).all() and (rgbw[:, :3] < 1).all()
    # stratified: sample s of a pixel lies in cell (s % 2, s // 2)
    cell_x = np.floor((xy[:, 0] - px) * 2).astype(int)
    assert (cell_x == np.tile([0, 1, 0, 1], 12)).all()


def test_ext_c1_matches_committed_golden(orc):
    """The oracle reproduces tests/golden/ext_c1_golden.json (made by tests/golden/make_ext_golden.py)."""
    import hashlib
    import json
    from pathlib import Path

    g = json.loads((Path(__file__).resolve().parent / "golden" / "ext_c1_golden.json").read_text())
    for name, want in g["filters"].items():
        kind, rad, p0, p1 = oracle.FILTERS[name]
        table = oracle.filter_table(orc, kind, rad, p0, p1)
        assert hashlib.sha256(table.tobytes()).hexdigest() == want["table_sha256"], name
        film = OracleFilm(orc, (64, 64), [0, 0, 1, 1], rad, table)
        xy, rgbw = oracle.synth_samples(orc, (0, 0, 64, 64), 4, 1)
        film.add_samples_pass((0, 0, 64, 64), 4, xy, rgbw, threads=3)
        px = film.pixels()
        assert hashlib.sha256(np.ascontiguousarray(px[:, :4]).tobytes()).hexdigest() == want["pixels_sha256"], name
        assert hashlib.sha256(film.write_image_rgb(1.0).tobytes()).hexdigest() == want["rgb_sha256"], name
