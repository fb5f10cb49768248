"""Host-side mirror: geometry semantics and the image containers (no GPU)."""
import numpy as np
import pytest


def test_bounds2i_semantics(pb, kats):
    k = kats["bounds2i"]
    B = pb.Bounds2i
    b = B.of([[2, 2], [4, 4]])
    assert b.inside_exclusive((2, 2)) and not b.inside_exclusive((4, 4))
    assert B.intersect(B.of([[1, 1], [3, 3]]), B.of([[2, 2], [4, 4]])) == B.of([[2, 2], [3, 3]])
    inv = B.intersect(B.of([[1, 1], [2, 2]]), B.of([[3, 3], [4, 4]]))
    assert inv.as4() == tuple(k["intersect_disjoint"]["result"])
    assert inv.area() == 1 and list(inv.iter()) == []
    assert [tuple(p) for p in b.iter()] == [tuple(p) for p in k["iter"]["points"]]
    assert B.of([[5, 4], [3, 2]]).as4() == (3, 2, 5, 4)


def _gradient():
    res = 64
    ys, xs = np.mgrid[0:res, 0:res]
    px = np.stack([xs / np.float32(res), ys / np.float32(res), np.ones_like(xs, dtype=np.float32)], axis=-1)
    return px.astype(np.float32).reshape(-1, 3), res


def test_roundtrip_pfm(pb, tmp_path):
    """src/core/imageio.rs:362-390 — bit-exact."""
    from pbrt_b200 import imageio

    px, res = _gradient()
    name = str(tmp_path / "g.pfm")
    imageio.write_image(name, px.reshape(-1), [[0, 0], [res, res]], (res, res))
    got, r = imageio.read_image(name)
    assert (r.x, r.y) == (res, res)
    assert np.array_equal(got.view(np.uint32), px.view(np.uint32))


def test_roundtrip_png(pb, orc, tmp_path):
    """src/core/imageio.rs:325-360 — compare after to_byte / 255."""
    from pbrt_b200 import imageio

    px, res = _gradient()
    name = str(tmp_path / "g.png")
    imageio.write_image(name, px.reshape(-1), [[0, 0], [res, res]], (res, res))
    got, r = imageio.read_image(name)
    want = np.array([orc.orc_to_byte(float(v)) for v in px.reshape(-1)], dtype=np.float32) / np.float32(255)
    assert (r.x, r.y) == (res, res)
    assert np.array_equal(got.reshape(-1), want)


def test_pfm_bytes_match_oracle(pb, orc, tmp_path):
    import ctypes as C

    import oracle
    from pbrt_b200 import imageio

    px, res = _gradient()
    name = str(tmp_path / "o.pfm")
    imageio.write_pfm(name, px, (res, res))
    need = orc.orc_pfm_encode(oracle.fp(px), res, res, None, 0)
    buf = (C.c_uint8 * need)()
    orc.orc_pfm_encode(oracle.fp(px), res, res, buf, need)
    assert open(name, "rb").read() == bytes(buf)


def test_host_to_byte_matches_oracle(pb, orc):
    from pbrt_b200 import imageio

    v = np.linspace(-0.1, 1.2, 4001, dtype=np.float32)
    got = imageio.to_byte(v)
    want = np.array([orc.orc_to_byte(float(x)) for x in v], dtype=np.uint8)
    # numpy's powf and glibc's differ in the last ulp at most: at most a stray LSB on a tie
    assert (np.abs(got.astype(int) - want.astype(int)) <= 1).all()
    assert (got != want).mean() < 1e-3


def test_unknown_extension(pb, tmp_path):
    from pbrt_b200 import imageio

    with pytest.raises(ValueError):
        imageio.write_image(str(tmp_path / "a.xyz"), np.zeros(3, np.float32), [[0, 0], [1, 1]])
    with pytest.raises(NotImplementedError):
        imageio.read_image(str(tmp_path / "a.exr"))


def test_constant_texture_scalar_evaluate_is_host_side(pb, kats):
    k = kats["constant_texture"]
    si = pb.SurfaceInteraction()
    assert pb.ConstantTexture.new(k["float_value"]).evaluate(si) == 10.0
    assert list(pb.ConstantTexture.new(k["spectrum_value"]).evaluate(si)) == k["spectrum_value"]
    assert pb.create_constant_float_texture().evaluate(si) == k["float_default"]
    assert list(pb.create_constant_spectrum_texture().evaluate(si)) == k["spectrum_default"]
    assert pb.create_constant_float_texture(None, {"value": 10.0}).evaluate(si) == 10.0
    assert "ConstantTexture{" in repr(pb.ConstantTexture(10.0))


def test_to_byte_threshold_table_agrees_with_glibc(orc):
    """The device's to_byte is driven by a threshold table generated without libm (correctly rounded powf).
    Check it against the oracle (glibc powf, what Rust's f32::powf calls on Linux) at every threshold and
    at the float just below it."""
    import re
    import struct
    from pathlib import Path

    txt = (Path(__file__).resolve().parent.parent / "pbrt_b200" / "csrc" / "to_byte_table.inc").read_text()
    thr = [int(h, 16) for h in re.findall(r"0x([0-9a-f]{8})u", txt)]
    assert len(thr) == 256 and thr[0] == 0 and thr == sorted(thr)
    for k in range(1, 256):
        v = struct.unpack("<f", struct.pack("<I", thr[k]))[0]
        below = struct.unpack("<f", struct.pack("<I", thr[k] - 1))[0]
        assert orc.orc_to_byte(v) == k, (k, v)
        assert orc.orc_to_byte(below) == k - 1, (k, below)


def test_to_byte_slices_hold_at_most_one_threshold():
    """The resolve kernel looks to_byte up by the top bits of the float (exponent + 6 mantissa bits, a "slice")
    and settles the result with ONE comparison against the next threshold.  That is exact only if no slice
    contains two thresholds and the range below the first slice / above the last maps to 0 / 255."""
    import re
    from pathlib import Path

    txt = (Path(__file__).resolve().parent.parent / "pbrt_b200" / "csrc" / "to_byte_table.inc").read_text()
    thr = [int(h, 16) for h in re.findall(r"0x([0-9a-f]{8})u", txt)][1:]
    cu = (Path(__file__).resolve().parent.parent / "pbrt_b200" / "csrc" / "film.cu").read_text()
    shift = int(re.search(r"SLICE_SHIFT = (\d+);", cu).group(1))
    first = int(re.search(r"SLICE_FIRST_BITS = 0x([0-9a-f]+)u;", cu).group(1), 16)
    one = 0x3F800000
    assert first < thr[0], "inputs below the first slice must all map to byte 0"
    assert thr[-1] <= one, "1.0 must already map to 255"
    for start in range(first, one, 1 << shift):
        inside = [t for t in thr if start < t < start + (1 << shift)]
        assert len(inside) <= 1, (hex(start), inside)


def _geometry(pb, res, crop, radius, diag=35.0, rank=0, nranks=1):
    import ctypes as C

    from pbrt_b200 import _lib

    cropped, owned, sb = (C.c_int32 * 4)(), (C.c_int32 * 4)(), (C.c_int32 * 4)()
    ext = (C.c_float * 4)()
    _lib.check(_lib.lib.pbrt_film_geometry(res[0], res[1], _lib.f32arr(crop), _lib.f32arr(radius), diag, rank, nranks,
                                           cropped, owned, sb, ext))
    return tuple(cropped), tuple(owned), tuple(sb), tuple(ext)


def test_film_geometry_reference_doctests_without_a_device(pb, kats):
    """film.rs:161-164, :197-216, :252-262 through the library's host logic (no GPU needed)."""
    import ctypes as C

    from pbrt_b200 import _lib

    k = kats["film_1920x1080_crop_quarter_box8"]
    cropped, owned, sb, _ = _geometry(pb, k["resolution"], k["crop"], [8.0, 8.0])
    assert sb == (472, 262, 1448, 818)
    assert cropped == owned == (480, 270, 1440, 810)
    out, n = (C.c_int32 * 4)(), C.c_int64()
    for sample_bounds, want in (((0, 0, 1920, 1080), (480, 270, 1440, 810)), ((500, 500, 600, 600), (492, 492, 608, 608))):
        _lib.check(_lib.lib.pbrt_film_geometry_tile_bounds(_lib.i32x4(cropped), _lib.f32arr([8.0, 8.0]),
                                                           _lib.i32x4(sample_bounds), out, C.byref(n)))
        assert tuple(out) == want and n.value == (want[2] - want[0]) * (want[3] - want[1])
    for c in kats["film_800x600_physical_extent"]["crops"]:
        ext = _geometry(pb, [800, 600], c, [8.0, 8.0], diag=100.0)[3]
        assert ext == tuple(np.float32(v) for v in (-0.04, -0.03, 0.04, 0.03))


def test_film_geometry_matches_oracle_on_random_films(pb, orc):
    """Crop, sample, tile bounds (ints, bit-exact) and physical extent (f32, exact) against the oracle for random
    resolutions, crop windows (including inverted and empty ones), radii and sample bounds (including ones that
    miss the film)."""
    import ctypes as C

    import oracle
    from pbrt_b200 import _lib

    rng = np.random.default_rng(2024)
    table = np.ones(256, dtype=np.float32)
    out, n = (C.c_int32 * 4)(), C.c_int64()
    for _ in range(300):
        res = [int(rng.integers(1, 9000)), int(rng.integers(1, 5000))]
        crop = [float(np.float32(v)) for v in rng.random(4)]
        if rng.random() < 0.5:
            crop = [min(crop[0], crop[2]), min(crop[1], crop[3]), max(crop[0], crop[2]), max(crop[1], crop[3])]
        if rng.random() < 0.2:
            crop = [0.0, 0.0, 1.0, 1.0]
        radius = [float(np.float32(rng.choice([0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 8.0, rng.random() * 6 + 0.1]))) for _ in range(2)]
        diag = float(np.float32(rng.random() * 100 + 1))
        of = oracle.OracleFilm(orc, res, crop, radius, table, diagonal_mm=diag)
        cropped, owned, sb, ext = _geometry(pb, res, crop, radius, diag)
        assert cropped == tuple(of.cropped()) and owned == cropped
        assert sb == tuple(of.sample_bounds())
        assert np.array_equal(np.asarray(ext, np.float32).view(np.uint32),
                              np.asarray(of.physical_extent(), np.float32).view(np.uint32))
        for _ in range(4):
            x0, y0 = int(rng.integers(-50, res[0] + 50)), int(rng.integers(-50, res[1] + 50))
            q = (x0, y0, x0 + int(rng.integers(0, 200)), y0 + int(rng.integers(0, 200)))
            _lib.check(_lib.lib.pbrt_film_geometry_tile_bounds(_lib.i32x4(cropped), _lib.f32arr(radius), _lib.i32x4(q),
                                                               out, C.byref(n)))
            want = tuple(of.tile_bounds(q))
            assert tuple(out) == want, (res, crop, radius, q)
            # FilmTile::new allocates max(0, area) pixels and Bounds2i::area is a plain product (film.rs:446,
            # bounds.rs:195-198): a tile inverted on both axes has a positive count, in the reference too
            t = of.get_film_tile(q)
            assert n.value == orc.orc_tile_pixel_count(t), (res, crop, radius, q, want)
            orc.orc_tile_free(t)


def test_sharded_geometry_partitions_the_rows(pb):
    """pbrt_film_create_sharded's row blocks: disjoint, in order, covering the cropped bounds."""
    for res, crop, world in (([1920, 1080], [0, 0, 1, 1], 8), ([640, 483], [0.1, 0.2, 0.9, 0.77], 3), ([7680, 4320], [0, 0, 1, 1], 8)):
        rows = []
        for r in range(world):
            cropped, owned, _, _ = _geometry(pb, res, crop, [2.0, 2.0], rank=r, nranks=world)
            assert owned[0] == cropped[0] and owned[2] == cropped[2]
            rows.append((owned[1], owned[3]))
        assert rows[0][0] == cropped[1] and rows[-1][1] == cropped[3]
        assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
