"""EXTENSION parity (no reference parity exists: the reference has no add_sample): the CUDA
splat kernels against the CPU restatement of pbrt-v3 FilmTile::AddSample, on identical samples.

PBRT_SPLAT_EXACT must be bit-identical to the oracle (same order, mul then add); PBRT_SPLAT_FMA
and PBRT_SPLAT_ATOMIC must agree within 1e-5 relative error (BASELINE.json north_star)."""
import numpy as np
import pytest

import oracle
from oracle import OracleFilm

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5  # north_star: per-pixel RGB and weight sums within 1e-5 relative error


def u32(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def make_filter(pb, name, radius=None):
    kind, rad, p0, p1 = oracle.FILTERS[name]
    rad = radius or rad
    cls = {"box": pb.BoxFilter, "triangle": pb.TriangleFilter, "gaussian": pb.GaussianFilter,
           "mitchell": pb.MitchellFilter, "lanczos": pb.LanczosSincFilter}[name]
    if name in ("box", "triangle"):
        return cls(rad), kind, rad, p0, p1
    if name == "mitchell":
        return cls(rad, p0, p1), kind, rad, p0, p1
    return cls(rad, p0), kind, rad, p0, p1


def run_pair(pb, orc, name, res, crop, sb, spp, mode, radius=None, max_lum=float("inf"), jitter=None, seed=1):
    filt, kind, rad, p0, p1 = make_filter(pb, name, radius)
    table = oracle.filter_table(orc, kind, rad, p0, p1)
    film = pb.Film.new(res, [[crop[0], crop[1]], [crop[2], crop[3]]], filt, 35.0, "x.pfm", 1.0, max_lum)
    of = OracleFilm(orc, res, crop, rad, table, max_lum=max_lum)
    xy, rgbw = oracle.synth_samples(orc, sb, spp, seed)
    if jitter is not None:
        xy, rgbw = jitter(xy, rgbw)
    film.add_samples_tile([[sb[0], sb[1]], [sb[2], sb[3]]], spp, xy, rgbw, mode)
    film.check()
    of.add_samples_pass(sb, spp, xy, rgbw, threads=4)
    return film, of


def rel_err(got, want):
    """max over pixels of |got - want| / max(|want|, small) per channel group."""
    denom = np.maximum(np.abs(want), 1e-3)
    return float((np.abs(got - want) / denom).max())


# ------------------------------------------------------------------ config 1 of BASELINE.json

@pytest.mark.parametrize("name", list(oracle.FILTERS))
def test_c1_64x64_4spp_each_filter_exact(gpu, orc, name):
    """BASELINE.json configs[0]: each filter kernel on a 64x64 film, 4 spp."""
    film, of = run_pair(gpu, orc, name, (64, 64), [0, 0, 1, 1], (0, 0, 64, 64), 4, gpu.SPLAT_EXACT)
    assert np.array_equal(u32(film.read_pixels()), u32(of.pixels()))
    assert np.array_equal(u32(film.resolve_rgb(1.0)), u32(of.write_image_rgb(1.0)))


@pytest.mark.parametrize("name", list(oracle.FILTERS))
def test_c1_generic_gather_exact(gpu, orc, name):
    from pbrt_b200 import _lib

    _lib.lib.pbrt_b200_debug_force_generic_splat(1)
    try:
        film, of = run_pair(gpu, orc, name, (64, 64), [0, 0, 1, 1], (0, 0, 64, 64), 4, gpu.SPLAT_EXACT)
    finally:
        _lib.lib.pbrt_b200_debug_force_generic_splat(0)
    assert np.array_equal(u32(film.read_pixels()), u32(of.pixels()))


@pytest.mark.parametrize("name", list(oracle.FILTERS))
@pytest.mark.parametrize("mode", ["fma", "atomic"])
def test_c1_tolerance_modes(gpu, orc, name, mode):
    m = gpu.SPLAT_FMA if mode == "fma" else gpu.SPLAT_ATOMIC
    film, of = run_pair(gpu, orc, name, (64, 64), [0, 0, 1, 1], (0, 0, 64, 64), 4, m)
    got, want = film.read_pixels(), of.pixels()
    assert rel_err(got[:, :4], want[:, :4]) <= REL_TOL
    if mode == "fma":  # the weight sum involves no multiply: identical
        assert np.array_equal(u32(got[:, 3]), u32(want[:, 3]))


@pytest.mark.parametrize("name,res,sb,spp", [("gaussian", (64, 64), (0, 0, 64, 64), 4), ("lanczos", (90, 41), (-4, -4, 94, 45), 16),
                                             ("mitchell", (70, 30), (3, 2, 66, 29), 5), ("box", (40, 40), (0, 0, 40, 40), 1)])
def test_warp_aggregated_atomic_scatter_within_tolerance(gpu, orc, monkeypatch, name, res, sb, spp):
    """The shared-atomic scatter with one atomic per run of lanes that hold samples of the same pixel (kept for the ncu
    comparison, PBRT_B200_ATOMIC_AGG=1): a tree order of additions, inside the north-star tolerance."""
    monkeypatch.setenv("PBRT_B200_ATOMIC_AGG", "1")
    film, of = run_pair(gpu, orc, name, res, [0, 0, 1, 1], sb, spp, gpu.SPLAT_ATOMIC)
    got, want = film.read_pixels(), of.pixels()
    assert rel_err(got[:, :4], want[:, :4]) <= REL_TOL


# ------------------------------------------------------------------ shapes, clipping, ragged edges

CASES = [
    # name, res, crop, sample bounds, spp
    ("gaussian", (301, 77), [0, 0, 1, 1], "film", 16),          # sample bounds reach outside the film
    ("gaussian", (301, 77), [0.2, 0.1, 0.9, 0.8], "film", 4),   # cropped film
    ("mitchell", (130, 140), [0, 0, 1, 1], (5, 7, 120, 133), 9),  # tile strictly inside, spp = 3x3
    ("mitchell", (257, 40), [0, 0, 1, 1], (-2, -2, 259, 42), 1),  # 1 spp, strip width not a multiple of the CTA
    ("lanczos", (90, 70), [0, 0, 1, 1], "film", 4),             # h = 4
    ("triangle", (64, 33), [0, 0, 1, 1], (10, 3, 11, 30), 5),   # one pixel column of samples, spp not a square
    ("box", (50, 50), [0, 0, 1, 1], (0, 0, 50, 50), 7),         # h = 1, footprint 1x1 or 2x2
    ("gaussian", (40, 300), [0, 0, 1, 1], "film", 2),           # tall: many row segments
    ("gaussian", (64, 64), [0, 0, 1, 1], (70, 70, 80, 80), 4),  # tile misses the film entirely
    ("gaussian", (64, 64), [0, 0, 1, 1], (62, 62, 80, 80), 4),  # tile touches one corner
]


@pytest.mark.parametrize("name,res,crop,sb,spp", CASES)
def test_clipping_and_ragged_shapes_exact(gpu, orc, name, res, crop, sb, spp):
    if sb == "film":
        _, kind, rad, p0, p1 = make_filter(gpu, name)
        sb = OracleFilm(orc, res, crop, rad, np.ones(256, np.float32)).sample_bounds()
    film, of = run_pair(gpu, orc, name, res, crop, sb, spp, gpu.SPLAT_EXACT)
    assert np.array_equal(u32(film.read_pixels()), u32(of.pixels()))


@pytest.mark.parametrize("radius", [(1.0, 1.0), (1.5, 1.5), (2.5, 2.5), (3.0, 3.0), (3.4, 3.4), (4.4, 4.4)])
def test_radii_covering_every_window_size(gpu, orc, radius):
    """h = floor(r + .5) = 1, 2, 3, 3, 3, 4 — including radii where r + .5 is an integer."""
    film, of = run_pair(gpu, orc, "gaussian", (80, 50), [0, 0, 1, 1], (-4, -4, 84, 54), 4, gpu.SPLAT_EXACT, radius=radius)
    assert np.array_equal(u32(film.read_pixels()), u32(of.pixels()))


@pytest.mark.parametrize("spp", [25, 64, 100, 256])
def test_large_spp_uses_narrower_strips_exact(gpu, orc, spp):
    """One sample row of a strip must fit in shared memory: 64 spp -> 64-column strips, 256 spp -> 32."""
    film, of = run_pair(gpu, orc, "gaussian", (70, 12), [0, 0, 1, 1], (0, 0, 70, 12), spp, gpu.SPLAT_EXACT)
    assert np.array_equal(u32(film.read_pixels()), u32(of.pixels()))


@pytest.mark.parametrize("name", ["gaussian", "lanczos"])
@pytest.mark.parametrize("spp", [1, 3, 6, 12, 20, 24, 32, 48, 64])
def test_samples_per_pixel_the_strip_width_is_chosen_for(gpu, orc, monkeypatch, name, spp):
    """The class kernel picks 128 / 96 / 64 columns so that spp divides the strip (a thread keeps its sample index:
    uniform path) and otherwise runs every sample per lane: 1..32 spp at radius 2 and 4, a film wider than two strips;
    48 and 64 spp with 64-bit index masks on 96 / 64 columns (PBRT_B200_WIDE=1; off by default: the window kernel is
    faster there)."""
    monkeypatch.setenv("PBRT_B200_WIDE", "1")
    film, of = run_pair(gpu, orc, name, (300, 21), [0, 0, 1, 1], (-2, -2, 302, 23), spp, gpu.SPLAT_EXACT, seed=spp)
    assert np.array_equal(u32(film.read_pixels()), u32(of.pixels()))


@pytest.mark.parametrize("radius", [(8.0, 8.0), (2.0, 3.0), (0.3, 0.3), (5.0, 5.0)])
def test_radii_served_by_the_generic_gather(gpu, orc, radius):
    film, of = run_pair(gpu, orc, "triangle", (48, 40), [0, 0, 1, 1], (0, 0, 48, 40), 4, gpu.SPLAT_EXACT, radius=radius)
    assert np.array_equal(u32(film.read_pixels()), u32(of.pixels()))


def test_samples_on_exact_pixel_and_half_pixel_positions(gpu, orc):
    """Positions where ceil/floor of (p - .5 -+ r) and the table bins sit exactly on their boundaries."""
    def snap(xy, rgbw):
        xy = xy.copy()
        base = np.floor(xy)
        frac = np.tile(np.array([[0.0, 0.0], [0.5, 0.5], [0.0, 0.5], [0.25, 0.75], [0.5, 0.0], [0.125, 0.375],
                                 [0.75, 0.25], [0.9999999, 0.5], [0.5, 0.9999999]], dtype=np.float32), (len(xy) // 9 + 1, 1))[: len(xy)]
        return (base + frac).astype(np.float32), rgbw
    for name in ("gaussian", "mitchell", "lanczos", "box"):
        film, of = run_pair(gpu, orc, name, (40, 30), [0, 0, 1, 1], (-4, -4, 44, 34), 9, gpu.SPLAT_EXACT, jitter=snap)
        assert np.array_equal(u32(film.read_pixels()), u32(of.pixels())), name


@pytest.mark.parametrize("name", ["gaussian", "lanczos"])
def test_phases_one_ulp_below_a_pixel_centre_at_power_of_two_coordinates(gpu, orc, name):
    """pd + r crosses into a coarser float grid just below a power of two: a sample one ulp left of (above) a pixel
    centre then reaches the pixel r to its right (below) although its phase is negative.  Snap a third of the samples
    there, on both axes, around x = 1024 and y = 32."""
    def snap(xy, rgbw):
        xy = xy.copy()
        c = np.floor(xy) + np.float32(0.5)
        below = np.nextafter(c, np.float32(-np.inf)).astype(np.float32)
        above = np.nextafter(c, np.float32(np.inf)).astype(np.float32)
        xy[0::6, 0] = below[0::6, 0]
        xy[1::6, 1] = below[1::6, 1]
        xy[2::6] = below[2::6]
        xy[3::6, 0] = above[3::6, 0]
        xy[4::6] = c[4::6]
        return xy, rgbw
    film, of = run_pair(gpu, orc, name, (1100, 48), [0, 0, 1, 1], (1008, 20, 1040, 44), 16, gpu.SPLAT_EXACT, jitter=snap)
    assert np.array_equal(u32(film.read_pixels()), u32(of.pixels()))


def test_max_sample_luminance_and_weights(gpu, orc):
    def bright(xy, rgbw):
        rgbw = rgbw.copy()
        rgbw[::3, :3] *= 40.0                       # above the clamp
        rgbw[:, 3] = 0.25 + (np.arange(len(rgbw)) % 5) * 0.5   # sample weights != 1
        rgbw[::17, :3] *= -1.0                      # negative radiance stays unclamped
        return xy, rgbw
    film, of = run_pair(gpu, orc, "mitchell", (70, 45), [0, 0, 1, 1], (0, 0, 70, 45), 8, gpu.SPLAT_EXACT,
                        max_lum=2.5, jitter=bright)
    assert np.array_equal(u32(film.read_pixels()), u32(of.pixels()))


def test_accumulates_across_calls_like_successive_tile_merges(gpu, orc):
    """Several passes into one film == the oracle merging one tile per pass (order of passes kept)."""
    res, spp = (96, 64), 4
    filt, kind, rad, p0, p1 = make_filter(gpu, "mitchell")
    table = oracle.filter_table(orc, kind, rad, p0, p1)
    film = gpu.Film.new(res, [[0, 0], [1, 1]], filt, 35.0, "x.pfm", 1.0, float("inf"))
    of = OracleFilm(orc, res, [0, 0, 1, 1], rad, table)
    for seed, sb in ((1, (0, 0, 96, 64)), (2, (0, 0, 96, 64)), (3, (20, 10, 70, 50))):
        xy, rgbw = oracle.synth_samples(orc, sb, spp, seed)
        film.add_samples_tile([[sb[0], sb[1]], [sb[2], sb[3]]], spp, xy, rgbw, gpu.SPLAT_EXACT)
        of.add_samples_pass(sb, spp, xy, rgbw)
    film.check()
    assert np.array_equal(u32(film.read_pixels()), u32(of.pixels()))


def test_device_resident_streams_and_synth_generator(gpu, orc):
    """The device sample generator reproduces the oracle's stream bit for bit; device pointers are used in place."""
    from pbrt_b200 import synth

    b = (3, 2, 67, 45)
    xy_d, rgbw_d, n = synth.samples(b, 16, seed=1)
    xy, rgbw = oracle.synth_samples(orc, b, 16, 1)
    assert np.array_equal(u32(xy_d.to_numpy(np.float32, (n, 2))), u32(xy))
    assert np.array_equal(u32(rgbw_d.to_numpy(np.float32, (n, 4))), u32(rgbw))
    filt, kind, rad, p0, p1 = make_filter(gpu, "gaussian")
    film = gpu.Film.new((70, 50), [[0, 0], [1, 1]], filt, 35.0, "x.pfm", 1.0, float("inf"))
    of = OracleFilm(orc, (70, 50), [0, 0, 1, 1], rad, oracle.filter_table(orc, kind, rad, p0, p1))
    film.add_samples_tile([[3, 2], [67, 45]], 16, xy_d, rgbw_d, gpu.SPLAT_EXACT)
    film.check()
    of.add_samples_pass(b, 16, xy, rgbw)
    assert np.array_equal(u32(film.read_pixels()), u32(of.pixels()))
    # a shard's generator call reproduces its slice of the whole stream
    xy_s, rgbw_s, ns = synth.samples((3, 10, 67, 20), 16, seed=1, index_bounds=b)
    sl = slice((10 - 2) * 64 * 16, (20 - 2) * 64 * 16)
    assert np.array_equal(u32(xy_s.to_numpy(np.float32, (ns, 2))), u32(xy[sl]))


def test_not_pixel_major_is_reported(gpu, orc):
    filt, kind, rad, p0, p1 = make_filter(gpu, "gaussian")
    film = gpu.Film.new((32, 32), [[0, 0], [1, 1]], filt, 35.0, "x.pfm", 1.0, float("inf"))
    xy, rgbw = oracle.synth_samples(orc, (0, 0, 32, 32), 4)
    xy[100, 0] += 3.0  # leaves its nominal pixel
    film.add_samples_tile([[0, 0], [32, 32]], 4, xy, rgbw, gpu.SPLAT_EXACT)
    with pytest.raises(gpu.PbrtError) as e:
        film.check()
    assert e.value.code == 4
    film.check()  # the sticky error is cleared once reported


def test_arbitrary_order_add_samples_within_tolerance(gpu, orc):
    """pbrt_film_add_samples: shuffled samples, global atomics; and FilmTile.add_sample on the host mirror."""
    res, spp = (48, 36), 4
    filt, kind, rad, p0, p1 = make_filter(gpu, "gaussian")
    table = oracle.filter_table(orc, kind, rad, p0, p1)
    xy, rgbw = oracle.synth_samples(orc, (0, 0, *res), spp)
    of = OracleFilm(orc, res, [0, 0, 1, 1], rad, table)
    of.add_samples_pass((0, 0, *res), spp, xy, rgbw)
    perm = np.random.default_rng(1).permutation(len(xy))
    film = gpu.Film.new(res, [[0, 0], [1, 1]], filt, 35.0, "x.pfm", 1.0, float("inf"))
    film.add_samples([[0, 0], list(res)], xy[perm], rgbw[perm])
    assert rel_err(film.read_pixels()[:, :4], of.pixels()[:, :4]) <= REL_TOL
    film2 = gpu.Film.new(res, [[0, 0], [1, 1]], filt, 35.0, "x.pfm", 1.0, float("inf"))
    tile = film2.get_film_tile([[0, 0], list(res)])
    for p, l in zip(xy[:500], rgbw[:500]):
        tile.add_sample(p, l[:3], float(l[3]))
    film2.merge_film_tile(tile)
    of2 = OracleFilm(orc, res, [0, 0, 1, 1], rad, table)
    ot = of2.get_film_tile((0, 0, *res))
    orc.orc_ext_tile_add_samples(ot, 500, oracle.fp(np.ascontiguousarray(xy[:500])), oracle.fp(np.ascontiguousarray(rgbw[:500])))
    of2.merge(ot)
    assert rel_err(film2.read_pixels()[:, :4], of2.pixels()[:, :4]) <= REL_TOL


# ------------------------------------------------------------------ full-size properties (BASELINE configs[1])

def test_full_size_c2_properties(gpu, orc):
    """1920x1080, 16 spp, Gaussian r=2 (configs[1]): a cropped band against the oracle bit for bit, and
    size-independent properties on the whole frame: determinism, linearity in L, weight sums independent of L."""
    from pbrt_b200 import synth

    res, spp = (1920, 1080), 16
    filt, kind, rad, p0, p1 = make_filter(gpu, "gaussian")
    table = oracle.filter_table(orc, kind, rad, p0, p1)
    b = (0, 0, *res)
    xy_d, rgbw_d, n = synth.samples(b, spp, seed=1)
    films = []
    for _ in range(2):
        f = gpu.Film.new(res, [[0, 0], [1, 1]], filt, 35.0, "x.pfm", 1.0, float("inf"))
        f.add_samples_tile([[0, 0], list(res)], spp, xy_d, rgbw_d, gpu.SPLAT_EXACT)
        f.check()
        films.append(f.read_pixels())
    assert np.array_equal(u32(films[0]), u32(films[1]))                       # run-to-run deterministic
    # oracle on a 1920 x 40 band (rows 500..540): same samples (shard of the stream), bit-exact
    y0, y1, halo = 500, 540, 3
    band = (0, y0 - halo, 1920, y1 + halo)
    xy_b, rgbw_b, nb = synth.samples(band, spp, seed=1, index_bounds=b)
    of = OracleFilm(orc, res, [0, y0 / 1080, 1, y1 / 1080], rad, table)
    assert of.cropped() == (0, y0, 1920, y1)
    of.add_samples_pass(band, spp, xy_b.to_numpy(np.float32, (nb, 2)), rgbw_b.to_numpy(np.float32, (nb, 4)), threads=8)
    got = films[0].reshape(1080, 1920, 7)[y0:y1].reshape(-1, 7)
    assert np.array_equal(u32(got), u32(of.pixels()))
    # weights: every interior pixel received 16 spp x its 5x5 neighbourhood; positive and smooth
    w = films[0][:, 3].reshape(1080, 1920)
    assert w.min() > 0 and abs(w[10:-10, 10:-10].mean() / w[540, 960] - 1) < 0.05
    # linearity: doubling L (a power of two) doubles xyz exactly and leaves the weights untouched
    rgbw = rgbw_d.to_numpy(np.float32, (n, 4))
    rgbw[:, :3] *= 2.0
    rg2 = gpu.DeviceBuffer.from_numpy(rgbw)
    f2 = gpu.Film.new(res, [[0, 0], [1, 1]], filt, 35.0, "x.pfm", 1.0, float("inf"))
    f2.add_samples_tile([[0, 0], list(res)], spp, xy_d, rg2, gpu.SPLAT_EXACT)
    p2 = f2.read_pixels()
    assert np.array_equal(u32(p2[:, 3]), u32(films[0][:, 3]))
    assert np.array_equal(u32(p2[:, :3]), u32(films[0][:, :3] * np.float32(2.0)))
    # fma mode at full size stays inside the tolerance
    f3 = gpu.Film.new(res, [[0, 0], [1, 1]], filt, 35.0, "x.pfm", 1.0, float("inf"))
    f3.add_samples_tile([[0, 0], list(res)], spp, xy_d, rgbw_d, gpu.SPLAT_FMA)
    assert rel_err(f3.read_pixels()[:, :4], films[0][:, :4]) <= REL_TOL


def test_back_to_back_passes_overlap_without_changing_a_bit(gpu, orc, monkeypatch):
    """Passes over the same resident streams are launched as programmatic dependents (pass i+1 starts in the SM slots
    pass i's early CTAs leave and waits for it before touching the film).  Full-size configs[1], three passes: equal,
    bit for bit, to the same passes with the overlap switched off, and a band of it to the oracle's three passes."""
    from pbrt_b200 import synth

    res, spp = (1920, 1080), 16
    filt, kind, rad, p0, p1 = make_filter(gpu, "gaussian")
    table = oracle.filter_table(orc, kind, rad, p0, p1)
    b = (0, 0, *res)
    xy_d, rgbw_d, n = synth.samples(b, spp, seed=3)
    out = []
    for no_pdl in ("0", "1"):
        monkeypatch.setenv("PBRT_B200_NO_PDL", no_pdl)
        f = gpu.Film.new(res, [[0, 0], [1, 1]], filt, 35.0, "x.pfm", 1.0, float("inf"))
        for _ in range(3):
            f.add_samples_tile([[0, 0], list(res)], spp, xy_d, rgbw_d, gpu.SPLAT_EXACT)
        f.check()
        out.append(f.read_pixels())
    assert np.array_equal(u32(out[0]), u32(out[1]))
    y0, y1, halo = 1020, 1040, 3   # rows around 1024: the pre-pass's power-of-two check tier as well
    band = (0, y0 - halo, 1920, y1 + halo)
    xy_b, rgbw_b, nb = synth.samples(band, spp, seed=3, index_bounds=b)
    of = OracleFilm(orc, res, [0, y0 / 1080, 1, y1 / 1080], rad, table)
    assert of.cropped() == (0, y0, 1920, y1)
    for _ in range(3):
        of.add_samples_pass(band, spp, xy_b.to_numpy(np.float32, (nb, 2)), rgbw_b.to_numpy(np.float32, (nb, 4)), threads=8)
    got = out[0].reshape(1080, 1920, 7)[y0:y1].reshape(-1, 7)
    assert np.array_equal(u32(got), u32(of.pixels()))


def test_overlap_passes_over_alternating_buffers_changes_no_bit(gpu, orc):
    """pbrt_b200_overlap_passes(1): consecutive passes overlap although they read different sample buffers (both complete
    before the loop).  Four passes alternating between two buffers on the full-size configs[1] film: bit-identical to
    the same passes without the option."""
    from pbrt_b200 import synth

    res, spp = (1920, 1080), 16
    filt, *_ = make_filter(gpu, "gaussian")
    b = (0, 0, *res)
    bufs = [synth.samples(b, spp, seed=s)[:2] for s in (5, 6)]
    out = []
    for on in (False, True):
        assert gpu.overlap_passes(on) is False
        f = gpu.Film.new(res, [[0, 0], [1, 1]], filt, 35.0, "x.pfm", 1.0, float("inf"))
        for i in range(4):
            f.add_samples_tile([[0, 0], list(res)], spp, *bufs[i & 1], gpu.SPLAT_EXACT)
        f.check()
        out.append(f.read_pixels())
        gpu.overlap_passes(False)
    assert np.array_equal(u32(out[0]), u32(out[1]))


@pytest.mark.parametrize("name", list(oracle.FILTERS))
def test_c1_matches_committed_golden(gpu, name):
    """CUDA path against the committed fixture alone (no oracle call): device sample generator -> splat -> resolve."""
    import hashlib
    import json
    from pathlib import Path

    from pbrt_b200 import synth

    g = json.loads((Path(__file__).resolve().parent / "golden" / "ext_c1_golden.json").read_text())["filters"][name]
    filt, kind, rad, p0, p1 = make_filter(gpu, name)
    assert hashlib.sha256(gpu.filter_table(filt).tobytes()).hexdigest() == g["table_sha256"]
    film = gpu.Film.new((64, 64), [[0, 0], [1, 1]], filt, 35.0, "x.pfm", 1.0, float("inf"))
    xy, rgbw, n = synth.samples((0, 0, 64, 64), 4, seed=1)
    film.add_samples_tile((0, 0, 64, 64), 4, xy, rgbw, gpu.SPLAT_EXACT)
    film.check()
    px = film.read_pixels()
    assert [int(v) for v in px[32 * 64 + 32, :4].view(np.uint32)] == g["pixel_32_32_xyzw_bits"]
    assert hashlib.sha256(np.ascontiguousarray(px[:, :4]).tobytes()).hexdigest() == g["pixels_sha256"]
    assert hashlib.sha256(film.resolve_rgb(1.0).tobytes()).hexdigest() == g["rgb_sha256"]


@pytest.mark.parametrize("name,tile,spp,crop", [("gaussian", 16, 4, [0, 0, 1, 1]), ("mitchell", 16, 16, [0.1, 0.15, 0.9, 0.95]),
                                                  ("lanczos", 24, 4, [0, 0, 1, 1]), ("box", 8, 4, [0, 0, 1, 1]),
                                                  ("triangle", 16, 3, [0, 0, 1, 1])])
def test_batched_tiles_equal_sequential_tiles(gpu, orc, name, tile, spp, crop):
    """pbrt's workflow — many small tiles, each get_film_tile -> add_sample* -> merge_film_tile — in one call.
    The oracle does it tile by tile in order; pixels in overlapping tile borders must still match bit for bit.
    The triangle case uses radius 5 (outside the window kernel) and exercises the per-tile fallback."""
    res = (150, 70)
    radius = (5.0, 5.0) if name == "triangle" else None
    filt, kind, rad, p0, p1 = make_filter(gpu, name, radius)
    table = oracle.filter_table(orc, kind, rad, p0, p1)
    film = gpu.Film.new(res, [[crop[0], crop[1]], [crop[2], crop[3]]], filt, 35.0, "x.pfm", 1.0, float("inf"))
    of = OracleFilm(orc, res, crop, rad, table)
    sbx = of.sample_bounds()
    sbs = [(x, y, min(x + tile, sbx[2]), min(y + tile, sbx[3])) for y in range(sbx[1], sbx[3], tile) for x in range(sbx[0], sbx[2], tile)]
    sbs.append((10, 10, 10, 20))  # an empty tile in the middle of the batch
    rng = np.random.default_rng(2)
    rng.shuffle(sbs)              # order matters where borders overlap; any order must work
    xs, ls = [], []
    for i, sb in enumerate(sbs):
        xy, rgbw = oracle.synth_samples(orc, sb, spp, seed=i + 1)
        xs.append(xy)
        ls.append(rgbw)
        t = of.get_film_tile(sb)
        if len(xy):
            orc.orc_ext_tile_add_samples(t, len(xy), oracle.fp(xy), oracle.fp(rgbw))
        of.merge(t)
    xy_all, rgbw_all = np.concatenate(xs), np.concatenate(ls)
    for _ in range(2):  # twice: the second call accumulates on top and reuses the cached merge index
        film.add_samples_tiles(sbs, spp, xy_all, rgbw_all, mode=gpu.SPLAT_EXACT)
    film.check()
    for i, sb in enumerate(sbs):
        t = of.get_film_tile(sb)
        if len(xs[i]):
            orc.orc_ext_tile_add_samples(t, len(xs[i]), oracle.fp(xs[i]), oracle.fp(ls[i]))
        of.merge(t)
    assert np.array_equal(u32(film.read_pixels()), u32(of.pixels()))


@pytest.mark.parametrize("seed", range(12))
def test_random_configurations_exact(gpu, orc, seed):
    """Randomised sweep: film size, crop window, filter, radius, spp, sample bounds (inside, straddling, outside)."""
    rng = np.random.default_rng(1000 + seed)
    name = list(oracle.FILTERS)[rng.integers(0, 5)]
    r = float(rng.choice([0.5, 1.0, 1.25, 1.5, 2.0, 2.3, 2.5, 3.0, 3.7, 4.0, 4.49, 6.0]))
    res = (int(rng.integers(8, 200)), int(rng.integers(8, 120)))
    c = np.sort(rng.random(2) * 0.4)
    crop = [float(c[0]), float(rng.random() * 0.3), float(1 - c[1] * 0.5), float(1 - rng.random() * 0.3)]
    spp = int(rng.choice([1, 2, 3, 4, 8, 9, 16, 17]))
    filt, kind, rad, p0, p1 = make_filter(gpu, name, (r, r))
    table = oracle.filter_table(orc, kind, rad, p0, p1)
    film = gpu.Film.new(res, [[crop[0], crop[1]], [crop[2], crop[3]]], filt, 35.0, "x.pfm", 1.0, float("inf"))
    of = OracleFilm(orc, res, crop, rad, table)
    assert film.cropped_pixel_bounds.as4() == of.cropped()
    assert film.get_sample_bounds().as4() == of.sample_bounds()
    for k in range(3):
        x0, y0 = int(rng.integers(-10, res[0])), int(rng.integers(-10, res[1]))
        sb = (x0, y0, x0 + int(rng.integers(1, res[0] + 10)), y0 + int(rng.integers(1, res[1] + 10)))
        assert film._tile_bounds(sb)[0].as4() == of.tile_bounds(sb)
        xy, rgbw = oracle.synth_samples(orc, sb, spp, seed=seed * 10 + k)
        rgbw[:, 3] = rng.random(len(rgbw), dtype=np.float32) + 0.5
        film.add_samples_tile(sb, spp, xy, rgbw, gpu.SPLAT_EXACT)
        of.add_samples_pass(sb, spp, xy, rgbw, threads=4)
    film.check()
    assert np.array_equal(u32(film.read_pixels()), u32(of.pixels())), (name, r, res, crop, spp)
    assert np.array_equal(u32(film.resolve_rgb(0.5)), u32(of.write_image_rgb(0.5)))


@pytest.mark.parametrize("name,res,band", [("mitchell", (3840, 2160), (1000, 1012)), ("lanczos", (7680, 4320), (4312, 4320))])
def test_full_size_c3_c5_band_exact(gpu, orc, name, res, band):
    """BASELINE configs[2] and [4] at full film size (one 16-spp pass, as the stream is fed): a band of rows
    against the oracle bit for bit (for C5 the band touches the bottom edge of the film), run-to-run determinism,
    and the sharded run (8 row shards, as on 8 GPUs) equal to the single film."""
    from pbrt_b200 import dist as pdist
    from pbrt_b200 import synth

    spp = 16
    filt, kind, rad, p0, p1 = make_filter(gpu, name)
    table = oracle.filter_table(orc, kind, rad, p0, p1)
    full = (0, 0, *res)
    xy_d, rgbw_d, n = synth.samples(full, spp, seed=1)
    film = gpu.Film.new(res, [[0, 0], [1, 1]], filt, 35.0, "x.pfm", 1.0, float("inf"))
    film.add_samples_tile(full, spp, xy_d, rgbw_d, gpu.SPLAT_EXACT)
    film.check()
    ref = film.resolve_rgb(1.0).reshape(res[1], res[0], 3)
    px = film.read_pixels().reshape(res[1], res[0], 7)
    y0, y1 = band
    halo = pdist.halo_rows(rad[1]) + 1
    sb = (0, max(0, y0 - halo), res[0], min(res[1], y1 + halo))
    xy_b, rgbw_b, nb = synth.samples(sb, spp, seed=1, index_bounds=full)
    of = OracleFilm(orc, res, [0, y0 / res[1], 1, y1 / res[1]], rad, table)
    assert of.cropped() == (0, y0, res[0], y1)
    of.add_samples_pass(sb, spp, xy_b.to_numpy(np.float32, (nb, 2)), rgbw_b.to_numpy(np.float32, (nb, 4)), threads=8)
    assert np.array_equal(u32(px[y0:y1].reshape(-1, 7)), u32(of.pixels()))
    del px
    # the same frame from 8 row shards, each fed its rows + halo
    world = 8
    for rank in (0, 3, 7):
        f = gpu.Film.new(res, [[0, 0], [1, 1]], filt, 35.0, "x.pfm", 1.0, float("inf"), rank=rank, nranks=world)
        ob = f.owned_pixel_bounds
        ssb = pdist.shard_sample_bounds(film.cropped_pixel_bounds, (ob.p_min.y, ob.p_max.y), rad[1])
        sxy, srgbw, sn = synth.samples(ssb.as4(), spp, seed=1, index_bounds=full)
        f.add_samples_tile(ssb.as4(), spp, sxy, srgbw, gpu.SPLAT_EXACT)
        f.check()
        got = f.resolve_rgb(1.0).reshape(-1, res[0], 3)
        assert np.array_equal(u32(got), u32(ref[ob.p_min.y:ob.p_max.y])), rank
        f.close()
        del sxy, srgbw


def test_filmtile_add_sample_in_renderer_order_is_exact(gpu, orc):
    """FilmTile.add_sample on the host mirror: samples recorded in pixel-major order go through the exact kernel."""
    res, spp = (24, 18), 4
    filt, kind, rad, p0, p1 = make_filter(gpu, "mitchell")
    table = oracle.filter_table(orc, kind, rad, p0, p1)
    film = gpu.Film.new(res, [[0, 0], [1, 1]], filt, 35.0, "x.pfm", 1.0, float("inf"))
    of = OracleFilm(orc, res, [0, 0, 1, 1], rad, table)
    sb = (2, 3, 20, 15)
    xy, rgbw = oracle.synth_samples(orc, sb, spp)
    tile = film.get_film_tile([[sb[0], sb[1]], [sb[2], sb[3]]])
    ot = of.get_film_tile(sb)
    for p, l in zip(xy, rgbw):
        tile.add_sample(p, l[:3], float(l[3]))
        orc.orc_ext_tile_add_sample(ot, float(p[0]), float(p[1]), oracle.farr(l[:3]), float(l[3]))
    film.merge_film_tile(tile)
    of.merge(ot)
    film.check()
    assert np.array_equal(u32(film.read_pixels()), u32(of.pixels()))


def test_filmtile_add_sample_with_non_finite_radiance_fails_at_the_merge(gpu):
    """merge_film_tile flushes the samples FilmTile.add_sample recorded and checks the film: a contract violation
    surfaces at the call that caused it (the reference's debug_assert!s), not at a later check()."""
    filt, *_ = make_filter(gpu, "gaussian")
    film = gpu.Film.new((16, 16), [[0, 0], [1, 1]], filt, 35.0, "x.pfm", 1.0, float("inf"))
    tile = film.get_film_tile([[4, 4], [6, 6]])
    for y in range(4, 6):
        for x in range(4, 6):
            tile.add_sample((x + 0.5, y + 0.5), (float("nan"), 1.0, 1.0), 1.0)
    with pytest.raises(Exception):
        film.merge_film_tile(tile)


def test_pinned_async_pipeline_equals_synchronous_calls(gpu, orc):
    """PBRT_MEM_PINNED_ASYNC: uploads double-buffered on the copy stream, read-back enqueued; same film, same frames."""
    res, spp = (128, 96), 4
    filt, kind, rad, p0, p1 = make_filter(gpu, "gaussian")
    a = gpu.Film.new(res, [[0, 0], [1, 1]], filt, 35.0, "x.pfm", 1.0, float("inf"))
    b = gpu.Film.new(res, [[0, 0], [1, 1]], filt, 35.0, "x.pfm", 1.0, float("inf"))
    n = res[0] * res[1] * spp
    hx = [gpu.PinnedBuffer(np.float32, (n, 2)) for _ in range(3)]
    hl = [gpu.PinnedBuffer(np.float32, (n, 4)) for _ in range(3)]
    outs = [gpu.PinnedBuffer(np.float32, (res[0] * res[1], 3)) for _ in range(3)]
    want = []
    for i in range(3):
        xy, rgbw = oracle.synth_samples(orc, (0, 0, *res), spp, seed=i + 1)
        hx[i].array[:] = xy
        hl[i].array[:] = rgbw
        a.add_samples_tile((0, 0, *res), spp, xy, rgbw, gpu.SPLAT_EXACT)
        want.append(a.resolve_rgb(1.0).copy())
    for i in range(3):  # nothing waits in between: three uploads, kernels and read-backs in flight
        b.add_samples_tile((0, 0, *res), spp, hx[i].array, hl[i].array, gpu.SPLAT_EXACT, pinned_async=True)
        b.resolve_rgb(1.0, out=outs[i].array, pinned_async=True)
    gpu.synchronize()
    b.check()
    for i in range(3):
        assert np.array_equal(u32(outs[i].array), u32(want[i])), i
    assert np.array_equal(u32(a.read_pixels()), u32(b.read_pixels()))


@pytest.mark.parametrize("weights", ["ones", "stream"])
def test_separate_rgb_and_weight_streams_equal_interleaved(gpu, orc, weights):
    """pbrt_film_add_samples_tile_rgb (AddSample's argument shape: L and sampleWeight apart; NULL weights = 1)
    interleaves on the device and must give the film of the rgbw form bit for bit — from host memory, from
    device memory and through the pinned double-buffered pipeline, odd sample counts included."""
    res, spp = (67, 45), 3          # 67*45*3 samples: not a multiple of 4 -> the pack kernel's tail path
    filt, kind, rad, p0, p1 = make_filter(gpu, "gaussian")
    sb = (0, 0, *res)
    xy, rgbw = oracle.synth_samples(orc, sb, spp, seed=9)
    rng = np.random.default_rng(3)
    if weights == "stream":
        rgbw[:, 3] = rng.random(len(rgbw), dtype=np.float32) + 0.5
    rgb = np.ascontiguousarray(rgbw[:, :3])
    sw = np.ascontiguousarray(rgbw[:, 3]) if weights == "stream" else None
    table = oracle.filter_table(orc, kind, rad, p0, p1)
    of = OracleFilm(orc, res, [0, 0, 1, 1], rad, table)
    ot = of.get_film_tile(sb)
    orc.orc_ext_tile_add_samples(ot, len(xy), oracle.fp(xy), oracle.fp(rgbw))
    of.merge(ot)
    want = u32(of.pixels())

    def fresh():
        return gpu.Film.new(res, [[0, 0], [1, 1]], filt, 35.0, "x.pfm", 1.0, float("inf"))

    f = fresh()                                    # host buffers
    f.add_samples_tile_rgb(sb, spp, xy, rgb, sw, gpu.SPLAT_EXACT)
    f.check()
    assert np.array_equal(u32(f.read_pixels()), want)

    f = fresh()                                    # device buffers
    dxy, drgb = gpu.DeviceBuffer.from_numpy(xy), gpu.DeviceBuffer.from_numpy(rgb)
    dsw = gpu.DeviceBuffer.from_numpy(sw) if sw is not None else None
    f.add_samples_tile_rgb(sb, spp, dxy, drgb, dsw, gpu.SPLAT_EXACT)
    f.check()
    assert np.array_equal(u32(f.read_pixels()), want)

    f = fresh()                                    # pinned, enqueued; twice, alternating staging sets
    hxy, hrgb = gpu.PinnedBuffer(np.float32, xy.shape), gpu.PinnedBuffer(np.float32, rgb.shape)
    hxy.array[:], hrgb.array[:] = xy, rgb
    hsw = None
    if sw is not None:
        hsw = gpu.PinnedBuffer(np.float32, sw.shape)
        hsw.array[:] = sw
    g = fresh()
    for film in (f, g, f):
        film.add_samples_tile_rgb(sb, spp, hxy.array, hrgb.array, None if hsw is None else hsw.array,
                                  gpu.SPLAT_EXACT, pinned_async=True)
    gpu.synchronize()
    g.check()
    assert np.array_equal(u32(g.read_pixels()), want)
    of.merge(_again(orc, of, sb, xy, rgbw))
    assert np.array_equal(u32(f.read_pixels()), u32(of.pixels()))


def _again(orc, of, sb, xy, rgbw):
    ot = of.get_film_tile(sb)
    orc.orc_ext_tile_add_samples(ot, len(xy), oracle.fp(xy), oracle.fp(rgbw))
    return ot
