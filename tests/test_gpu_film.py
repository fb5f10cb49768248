"""Tier 1 parity on the GPU: the reference's own film tests, run through the C ABI, and the CUDA
merge / resolve kernels against the CPU oracle — bit-exact."""
import ctypes as C

import numpy as np
import pytest

import oracle
from oracle import OracleFilm

pytestmark = pytest.mark.gpu


def u32(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def box8(pb):
    return pb.BoxFilter.new([8.0, 8.0])


def test_get_sample_bounds_and_tiles_doctests(gpu, kats):
    k = kats["film_1920x1080_crop_quarter_box8"]
    crop = [[k["crop"][0], k["crop"][1]], [k["crop"][2], k["crop"][3]]]
    film = gpu.Film.new(k["resolution"], crop, box8(gpu), 35.0, "output.png", 1.0, 1.0)
    assert film.get_sample_bounds() == gpu.Bounds2i.of([[472, 262], [1448, 818]])          # film.rs:161-164
    assert film.get_film_tile([[0, 0], [1920, 1080]]).get_pixel_bounds() == gpu.Bounds2i.of(
        [[1920 // 4, 1080 // 4], [3 * 1920 // 4, 3 * 1080 // 4]])                           # film.rs:252-256
    assert film.get_film_tile([[500, 500], [600, 600]]).get_pixel_bounds() == gpu.Bounds2i.of(
        [[492, 492], [608, 608]])                                                           # film.rs:258-262


def test_get_physical_extent_doctest(gpu, kats):
    k = kats["film_800x600_physical_extent"]
    want = gpu.Bounds2f.of([[-0.04, -0.03], [0.04, 0.03]])
    for c in k["crops"]:
        film = gpu.Film.new([800, 600], [[c[0], c[1]], [c[2], c[3]]], box8(gpu), 100.0, "output.png", 1.0, 1.0)
        assert film.get_physical_extent() == want  # exact equality, film.rs:197-200, :213-216


def test_merge_degenerate_tile_doctest(gpu):
    film = gpu.Film.new([20, 10], [[0, 0], [1, 1]], box8(gpu), 35.0, "output.png", 1.0, 1.0)
    left = film.get_film_tile([[0, 0], [10, 10]])
    right = film.get_film_tile([[10, 0], [10, 10]])
    film.merge_film_tile(left)
    film.merge_film_tile(right)  # film.rs:307-311: must not fail
    assert (film.read_pixels() == 0).all()


def _fill(tile, c):
    for pt in tile.get_pixel_bounds().iter():
        px = tile.get_pixel_mut(pt)
        px.contrib_sum = c
        px.filter_weight_sum = 1.0


def test_merge_film_tile_reference_test(gpu, orc, tmp_path):
    """src/core/film.rs:503-535, line for line."""
    name = str(tmp_path / "merge_film_tile.png")
    film = gpu.Film.new([200, 10], [[0, 0], [1, 1]], box8(gpu), 35.0, name, 1.0, 1.0)
    left = film.get_film_tile([[0, 0], [100, 10]])
    right = film.get_film_tile([[100, 0], [200, 10]])
    green, red = [0.0, 1.0, 0.0], [1.0, 0.0, 0.0]
    _fill(left, green)
    _fill(right, red)
    film.merge_film_tile(left)
    film.merge_film_tile(right)
    film.write_image(1.0)
    xg, xr = (C.c_float * 3)(), (C.c_float * 3)()
    orc.orc_rgb_to_xyz(oracle.farr(green), xg)
    orc.orc_rgb_to_xyz(oracle.farr(red), xr)
    assert film.get_pixel_xyz([4, 4]) == tuple(xg)
    assert film.get_pixel_xyz([196, 4]) == tuple(xr)
    assert [hex(v) for v in u32(film.get_pixel_xyz([100, 4]))] == ["0x3f4520e2", "0x3f6d8655", "0x3e0dda06"]
    from pbrt_b200 import imageio

    img, res = imageio.read_png8(name)
    img = img.reshape(10, 200, 3)
    assert (res.x, res.y) == (200, 10)
    assert list(img[4, 4]) == [0, 255, 0] and list(img[4, 196]) == [255, 0, 0] and list(img[4, 100]) == [188, 188, 0]


def test_merge_film_tile_rainbow_vs_oracle(gpu, orc, tmp_path):
    """src/core/film.rs:537-571 (a smoke test there) — here every pixel is compared with the oracle."""
    W, H = 200, 100
    film = gpu.Film.new([W, H], [[0, 0], [1, 1]], box8(gpu), 35.0, str(tmp_path / "rainbow.pfm"), 1.0, 1.0)
    of = OracleFilm(orc, [W, H], [0, 0, 1, 1], [8, 8], np.ones(256, np.float32), max_lum=1.0)
    f32 = np.float32
    for sb in ([[0, 0], [W // 2, H]], [[W // 2, 0], [W, H]]):
        t = film.get_film_tile(sb)
        ot = of.get_film_tile((sb[0][0], sb[0][1], sb[1][0], sb[1][1]))
        b = t.get_pixel_bounds()
        ys, xs = np.mgrid[b.p_min.y:b.p_max.y, b.p_min.x:b.p_max.x]
        px = np.stack([xs.astype(f32) / f32(W), ys.astype(f32) / f32(H), (W - xs).astype(f32) / f32(W),
                       np.ones_like(xs, dtype=f32)], axis=-1).reshape(-1, 4)
        t.pixels[:] = px
        of.tile_pixels(ot)[:] = px
        film.merge_film_tile(t)
        of.merge(ot)
    assert np.array_equal(u32(film.read_pixels()), u32(of.pixels()))
    assert np.array_equal(u32(film.resolve_rgb(1.0)), u32(of.write_image_rgb(1.0)))
    film.write_image(1.0)
    from pbrt_b200 import imageio

    back, _ = imageio.read_image(str(tmp_path / "rainbow.pfm"))
    assert np.array_equal(u32(back), u32(of.write_image_rgb(1.0)))


def test_merge_outside_film_is_range_error(gpu):
    film = gpu.Film.new([20, 10], [[0, 0], [1, 1]], box8(gpu), 35.0, "x.png", 1.0, 1.0)
    t = gpu.FilmTile(film, gpu.Bounds2i.raw(-1, 0, 5, 5), 30)
    with pytest.raises(gpu.PbrtError) as e:
        film.merge_film_tile(t)
    assert e.value.code == 3
    with pytest.raises(gpu.PbrtError):
        film.get_pixel_xyz([20, 0])  # film.rs:391-396
    tile = film.get_film_tile([[0, 0], [4, 4]])
    with pytest.raises(gpu.PbrtError):
        tile.get_pixel([100, 100])  # film.rs:466-471


def _random_tiles(film, of, orc, rng, sbs):
    tiles = []
    for i, sb in enumerate(sbs):
        t = film.get_film_tile([[sb[0], sb[1]], [sb[2], sb[3]]])
        ot = of.get_film_tile(sb)
        assert t.get_pixel_bounds().as4() == tuple(orc.orc_tile_get_pixel_bounds(ot).t())
        assert len(t.pixels) == orc.orc_tile_pixel_count(ot)
        vals = rng.random((len(t.pixels), 4), dtype=np.float32)
        t.pixels[:] = vals
        of.tile_pixels(ot)[:] = vals
        tiles.append((t, ot))
    return tiles


@pytest.mark.parametrize("crop", [[0, 0, 1, 1], [0.1, 0.2, 0.85, 0.9]])
def test_batched_merge_equals_sequential_oracle(gpu, orc, crop):
    """16x16 sample tiles with overlapping halos, merged in ONE launch == oracle merging them in order."""
    res, r = (150, 90), 2.0
    filt = gpu.GaussianFilter((r, r), 2.0)
    table = oracle.filter_table(orc, 2, (r, r), 2.0)
    film = gpu.Film.new(res, [[crop[0], crop[1]], [crop[2], crop[3]]], filt, 35.0, "x.pfm", 1.0, 1.0)
    of = OracleFilm(orc, res, crop, (r, r), table)
    sbx = film.get_sample_bounds().as4()
    assert sbx == of.sample_bounds()
    sbs = [(x, y, min(x + 16, sbx[2]), min(y + 16, sbx[3])) for y in range(sbx[1], sbx[3], 16) for x in range(sbx[0], sbx[2], 16)]
    rng = np.random.default_rng(7)
    tiles = _random_tiles(film, of, orc, rng, sbs)
    # prior content so that order matters everywhere, then two rounds
    film.merge_film_tiles([t for t, _ in tiles])
    for _, ot in tiles:
        of.merge(ot)
    for _ in range(2):  # the second time round the tiling repeats: the library reuses its cell index
        tiles = _random_tiles(film, of, orc, rng, sbs[::-1])
        film.merge_film_tiles([t for t, _ in tiles])
        for _, ot in tiles:
            of.merge(ot)
    assert np.array_equal(u32(film.read_pixels()), u32(of.pixels()))


def test_batched_merge_of_crowded_cells(gpu, orc):
    """Many large random tiles over a small film: every 16x16 cell is covered by far more than the four tiles its
    inline head holds (and by more than one 32-tile chunk of the overflow list), so the CSR walk is exercised;
    pixels must still receive the tiles in ascending order, bit for bit."""
    res, r = (72, 56), 1.5
    filt = gpu.GaussianFilter((r, r), 2.0)
    table = oracle.filter_table(orc, 2, (r, r), 2.0)
    film = gpu.Film.new(res, [[0, 0], [1, 1]], filt, 35.0, "x.pfm", 1.0, 1.0)
    of = OracleFilm(orc, res, [0, 0, 1, 1], (r, r), table)
    rng = np.random.default_rng(11)
    sbs = []
    for _ in range(160):
        x0, y0 = int(rng.integers(0, 50)), int(rng.integers(0, 36))
        sbs.append((x0, y0, x0 + int(rng.integers(6, 40)), y0 + int(rng.integers(6, 30))))
    sbs += [(10, 10, 10, 30), (0, 0, 72, 56)]  # an empty tile and one covering everything
    for _ in range(2):
        tiles = _random_tiles(film, of, orc, rng, sbs)
        film.merge_film_tiles([t for t, _ in tiles])
        for _, ot in tiles:
            of.merge(ot)
    assert np.array_equal(u32(film.read_pixels()), u32(of.pixels()))


def test_resolve_vs_oracle_with_splats_scale_and_zero_weight(gpu, orc):
    res = (97, 61)  # not a multiple of the resolve block: exercises the ragged tail
    film = gpu.Film.new(res, [[0, 0], [1, 1]], gpu.BoxFilter.new([0.5, 0.5]), 35.0, "x.pfm", 0.75, float("inf"))
    of = OracleFilm(orc, res, [0, 0, 1, 1], (0.5, 0.5), np.ones(256, np.float32), scale=0.75)
    rng = np.random.default_rng(3)
    t, ot = film.get_film_tile([[10, 5], [80, 50]]), of.get_film_tile((10, 5, 80, 50))
    vals = rng.random((len(t.pixels), 4), dtype=np.float32)
    vals[::7, 3] = 0.0      # weight 0: normalisation skipped (film.rs:355)
    vals[::11, :3] -= 0.6   # negative colours: clamped by max(0) (film.rs:359-361)
    t.pixels[:] = vals
    of.tile_pixels(ot)[:] = vals
    film.merge_film_tile(t)
    of.merge(ot)
    pts = rng.random((40, 2), dtype=np.float32) * np.array(res, dtype=np.float32)
    cols = rng.random((40, 3), dtype=np.float32)
    for p, c in zip(pts, cols):   # one at a time: the order of float atomics is then fixed
        film.add_splat(p, c)
        of.add_splat(float(p[0]), float(p[1]), c)
    assert np.array_equal(u32(film.read_pixels()), u32(of.pixels()))
    for ss in (1.0, 0.25):
        assert np.array_equal(u32(film.resolve_rgb(ss)), u32(of.write_image_rgb(ss)))
    want8 = np.array([orc.orc_to_byte(float(v)) for v in of.write_image_rgb(1.0).reshape(-1)], dtype=np.uint8)
    got8 = film.resolve_rgb8(1.0).reshape(-1)
    assert np.array_equal(got8, want8)


def test_resolve_rgb8_full_range_vs_oracle(gpu, orc):
    """to_byte on the device agrees with the CPU (glibc powf) over a dense sweep of values."""
    W, H = 512, 128
    film = gpu.Film.new([W, H], [[0, 0], [1, 1]], gpu.BoxFilter.new([0.5, 0.5]), 35.0, "x.png", 1.0, float("inf"))
    img = np.linspace(-0.05, 1.1, W * H * 3, dtype=np.float32).reshape(-1, 3)
    film.set_image(img)
    rgb = film.resolve_rgb(1.0)
    got = film.resolve_rgb8(1.0).reshape(-1)
    want = np.array([orc.orc_to_byte(float(v)) for v in rgb.reshape(-1)], dtype=np.uint8)
    assert np.array_equal(got, want)


def test_set_image_and_clear(gpu, orc):
    film = gpu.Film.new([16, 8], [[0, 0], [1, 1]], gpu.BoxFilter.new([0.5, 0.5]), 35.0, "x.pfm", 1.0, float("inf"))
    rng = np.random.default_rng(5)
    img = rng.random((128, 3), dtype=np.float32)
    film.add_splat((3.5, 2.5), (1, 1, 1))
    film.set_image(img)
    px = film.read_pixels()
    want = np.zeros((128, 3), dtype=np.float32)
    for i in range(128):
        o = (C.c_float * 3)()
        orc.orc_rgb_to_xyz(oracle.farr(img[i]), o)
        want[i] = list(o)
    assert np.array_equal(u32(px[:, :3]), u32(want)) and (px[:, 3] == 1).all() and (px[:, 4:] == 0).all()
    film.clear()
    assert (film.read_pixels() == 0).all()


def test_sharded_film_rows_and_tiles(gpu, orc):
    res, r = (64, 48), 2.0
    filt = gpu.GaussianFilter((r, r), 2.0)
    whole = gpu.Film.new(res, [[0, 0], [1, 1]], filt, 35.0, "x.pfm", 1.0, float("inf"))
    rng = np.random.default_rng(11)
    sb = whole.get_sample_bounds().as4()
    full_tile_px = rng.random(((res[0]) * (res[1]), 4), dtype=np.float32)
    t = whole.get_film_tile([[sb[0], sb[1]], [sb[2], sb[3]]])
    t.pixels[:] = full_tile_px
    whole.merge_film_tile(t)
    ref = whole.read_pixels().reshape(res[1], res[0], 7)
    for n in (2, 4):
        parts = []
        for rank in range(n):
            f = gpu.Film.new(res, [[0, 0], [1, 1]], filt, 35.0, "x.pfm", 1.0, float("inf"), rank=rank, nranks=n)
            ob = f.owned_pixel_bounds
            assert ob.p_min.y == res[1] * rank // n and ob.p_max.y == res[1] * (rank + 1) // n
            assert f.cropped_pixel_bounds.as4() == (0, 0, res[0], res[1])
            tt = f.get_film_tile([[sb[0], sb[1]], [sb[2], sb[3]]])
            assert tt.get_pixel_bounds() == ob  # clipped to the owned rows
            tt.pixels[:] = full_tile_px.reshape(res[1], res[0], 4)[ob.p_min.y:ob.p_max.y].reshape(-1, 4)
            f.merge_film_tile(tt)
            parts.append(f.read_pixels().reshape(-1, res[0], 7))
        assert np.array_equal(u32(np.concatenate(parts, axis=0)), u32(ref))


def test_resolve_rgb8_at_every_byte_threshold(gpu, orc):
    """Device to_byte at each of the 255 thresholds, one ulp either side, and at special values."""
    import re
    import struct
    from pathlib import Path

    txt = (Path(__file__).resolve().parent.parent / "pbrt_b200" / "csrc" / "to_byte_table.inc").read_text()
    thr = np.array([int(h, 16) for h in re.findall(r"0x([0-9a-f]{8})u", txt)][1:], dtype=np.uint32)
    vals = np.concatenate([thr - 1, thr, thr + 1]).view(np.float32)
    special = np.array([0.0, -0.0, -1.0, -1e-5, 1e-30, 0.0031308, 0.00313081, 1.0, 1.5, 1e30, np.inf, -np.inf, np.nan],
                       dtype=np.float32)
    vals = np.concatenate([vals, special, np.random.default_rng(0).random(4096 - len(vals) - len(special), dtype=np.float32)])
    n = len(vals) // 3
    img = vals[: 3 * n].reshape(n, 3)
    film = gpu.Film.new([n, 1], [[0, 0], [1, 1]], gpu.BoxFilter.new([0.5, 0.5]), 35.0, "x.png", 1.0, float("inf"))
    # put the values into the film as XYZ such that resolve returns them unchanged is not possible (3x3 matrices),
    # so drive to_byte through set_image -> resolve_rgb (float) -> compare bytes of the SAME floats
    film.set_image(img)
    rgb = film.resolve_rgb(1.0).reshape(-1)
    got = film.resolve_rgb8(1.0).reshape(-1)
    want = np.array([orc.orc_to_byte(float(v)) for v in rgb], dtype=np.uint8)
    assert np.array_equal(got, want)
    # and the thresholds themselves, fed as grey pixels with weight 1 straight into the film (xyz_to_rgb(to_xyz(v,v,v)) ~ v)
    assert len(np.unique(got)) > 200


def test_merge_and_resolve_special_values_bitwise(gpu, orc):
    """NaN, +-Inf, denormals, +-0 and huge values through merge_film_tile and the write_image loop: same bits as the CPU."""
    res = (64, 8)
    film = gpu.Film.new(res, [[0, 0], [1, 1]], gpu.BoxFilter.new([0.5, 0.5]), 35.0, "x.pfm", 1.5, float("inf"))
    of = OracleFilm(orc, res, [0, 0, 1, 1], (0.5, 0.5), np.ones(256, np.float32), scale=1.5)
    specials = np.array([0.0, -0.0, 1e-45, -1e-45, 1e-38, 3.4e38, -3.4e38, np.inf, -np.inf, np.nan, 1.0, -1.0, 0.5, 1e-20, 123456.78, 7e-8],
                        dtype=np.float32)
    rng = np.random.default_rng(9)
    for _ in range(3):
        t, ot = film.get_film_tile([[0, 0], list(res)]), of.get_film_tile((0, 0, *res))
        vals = rng.choice(specials, size=(len(t.pixels), 4)).astype(np.float32)
        t.pixels[:] = vals
        of.tile_pixels(ot)[:] = vals
        film.merge_film_tile(t)
        of.merge(ot)
    a, b = film.read_pixels(), of.pixels()
    # NaN payloads may differ between x86 and the GPU; compare NaN-ness, and bits everywhere else
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = ~np.isnan(b)
    assert np.array_equal(a.view(np.uint32)[m], b.view(np.uint32)[m])
    ra, rb = film.resolve_rgb(0.75), of.write_image_rgb(0.75)
    assert np.array_equal(np.isnan(ra), np.isnan(rb))
    m = ~np.isnan(rb)
    assert np.array_equal(ra.view(np.uint32)[m], rb.view(np.uint32)[m])
    got8 = film.resolve_rgb8(0.75).reshape(-1)
    want8 = np.array([orc.orc_to_byte(float(v)) for v in rb.reshape(-1)], dtype=np.uint8)
    assert np.array_equal(got8, want8)


def test_merge_film_tile_from_worker_threads(gpu):
    """FilmTile is Send and the reference serialises pixel access through one Mutex (film.rs:73, :316): worker
    threads may merge their tiles into one Film concurrently.  The ABI takes a lock per call; host tile buffers
    go through one per-film staging buffer, so an unlocked race would corrupt tiles.  Integer-valued tiles make
    the sums independent of the merge order."""
    import threading

    film = gpu.Film.new([256, 256], [[0, 0], [1, 1]], box8(gpu), 35.0, "t.png", 1.0, 1.0)
    rng = np.random.default_rng(5)
    jobs = []
    want = np.zeros((256, 256, 4), dtype=np.float64)
    for _ in range(64):
        x0, y0 = rng.integers(0, 200, size=2)
        w, h = rng.integers(8, 56, size=2)
        px = rng.integers(0, 8, size=(h, w, 4)).astype(np.float32)
        jobs.append(((int(x0), int(y0), int(x0 + w), int(y0 + h)), px))
        want[y0:y0 + h, x0:x0 + w] += px
    errors = []

    def worker(chunk):
        try:
            for tb, px in chunk:
                film.merge_tile_raw(tb, px.reshape(-1, 4))
        except Exception as e:  # noqa: BLE001 - surfaced below
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(jobs[i::8],)) for i in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    got = film.read_pixels().reshape(256, 256, 7)
    # weights: plain sum; xyz: rgb_to_xyz of small integers summed — compare the weight channel exactly and the
    # colour channels against the oracle matrix applied per tile in float64 (exact for these magnitudes to 1e-4)
    assert np.array_equal(got[..., 3], want[..., 3].astype(np.float32))
    m = np.array([[0.412453, 0.357580, 0.180423], [0.212671, 0.715160, 0.072169], [0.019334, 0.119193, 0.950227]])
    assert np.allclose(got[..., :3], want[..., :3] @ m.T, rtol=1e-5, atol=1e-4)
