"""Independent pin of the Tier-2 extension: the C oracle against a second restatement (tests/ref_py, numpy, written from
SURVEY.md App. A rather than from the oracle) and against analytic properties of the four extra filters.

Parity of Tier 2 stays "unpinned" by definition — the reference has no add_sample and no such filters — but a mistake
shared by the oracle and the kernels (one author) would have to be made a third time, differently structured, to pass."""
import math

import numpy as np
import pytest

import oracle
from oracle import OracleFilm
from ref_py import film_np


def u32(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("name", list(oracle.FILTERS))
def test_oracle_equals_numpy_restatement_on_c1(orc, name):
    """BASELINE configs[0]: 64x64 film, 4 spp, every filter — oracle == numpy scatter, bit for bit (same table)."""
    kind, rad, p0, p1 = oracle.FILTERS[name]
    table = oracle.filter_table(orc, kind, rad, p0, p1)
    xy, rgbw = oracle.synth_samples(orc, (0, 0, 64, 64), 4)
    of = OracleFilm(orc, (64, 64), [0, 0, 1, 1], rad, table)
    of.add_samples_pass((0, 0, 64, 64), 4, xy, rgbw)
    fn = film_np.FilmNp((64, 64), [0, 0, 1, 1], rad, table)
    fn.add_samples_pass(xy, rgbw)
    assert np.array_equal(u32(of.pixels()[:, :4]), u32(fn.pixels_xyzw()))


def test_oracle_equals_numpy_restatement_cropped_clamped_weighted(orc):
    """crop window, samples outside the film, luminance clamp, sample weights != 1, snapped positions"""
    kind, rad, p0, p1 = oracle.FILTERS["mitchell"]
    table = oracle.filter_table(orc, kind, rad, p0, p1)
    res, crop, sb, spp = (48, 40), [0.1, 0.2, 0.9, 0.85], (2, 5, 46, 37), 4
    xy, rgbw = oracle.synth_samples(orc, sb, spp, seed=5)
    rgbw[::3, :3] *= 30.0
    rgbw[:, 3] = 0.25 + (np.arange(len(rgbw)) % 5) * 0.5
    xy[::7] = np.floor(xy[::7]) + np.float32(0.5)     # phase 0 on both axes
    xy[3::11, 0] = np.floor(xy[3::11, 0])             # left pixel edge
    of = OracleFilm(orc, res, crop, rad, table, max_lum=2.5)
    of.add_samples_pass(sb, spp, xy, rgbw)
    fn = film_np.FilmNp(res, crop, rad, table, max_lum=2.5)
    fn.add_samples_pass(xy, rgbw)
    assert np.array_equal(u32(of.pixels()[:, :4]), u32(fn.pixels_xyzw()))


@pytest.mark.parametrize("name", list(oracle.FILTERS))
def test_filter_tables_close_to_float64_closed_forms(orc, name):
    """the f32 tables (libm expf / sinf) against the float64 closed forms written from App. A.2"""
    kind, rad, p0, p1 = oracle.FILTERS[name]
    got = oracle.filter_table(orc, kind, rad, p0, p1).astype(np.float64)
    want = film_np.filter_table(name, rad, p0, p1)
    # f32 polynomial cancellation near the edge of the support: absolute, not relative, agreement there
    assert np.allclose(got, want, rtol=2e-5, atol=6e-7), float(np.abs(got - want).max())


def test_mitchell_is_a_partition_of_unity():
    """B + 2C = 1 (here 1/3, 1/3): the integer shifts of the 1-D kernel over its radius-2 support sum to a constant"""
    B = C = 1.0 / 3.0
    # mitchell_1d takes x / radius in [-1, 1] and works on u = |2x| in [0, 2]: the Mitchell-Netravali cubic k(u).
    # Its unit shifts in u sum to one (the closed form is only ever evaluated inside the support: window it here)
    def k(u):
        return np.where(np.abs(u) <= 2.0, film_np.mitchell_1d(u * 0.5, B, C), 0.0)
    u = np.linspace(-0.5, 0.5, 101)
    total = sum(k(u + s) for s in range(-3, 4))
    assert np.allclose(total, 1.0, atol=1e-12)
    assert abs(film_np.mitchell_1d(np.array([0.0]), B, C)[0] - (6 - 2 * B) / 6) < 1e-15
    assert abs(film_np.mitchell_1d(np.array([1.0]), B, C)[0]) < 1e-15          # vanishes at the edge of the support
    # C1 continuity at the joint |2x| = 1
    e = 1e-7
    assert abs(film_np.mitchell_1d(np.array([0.5 - e]), B, C)[0] - film_np.mitchell_1d(np.array([0.5 + e]), B, C)[0]) < 1e-6


def test_gaussian_values():
    r, a = (2.0, 2.0), 2.0
    g0 = 1.0 - math.exp(-a * 4.0)
    assert abs(film_np.gaussian(0.0, 0.0, r, a) - g0 * g0) < 1e-15
    assert film_np.gaussian(2.0, 0.0, r, a) == 0.0 and film_np.gaussian(0.3, 2.5, r, a) == 0.0   # zero at and beyond the radius
    xs = np.linspace(0, 2, 50)
    v = film_np.gaussian(xs, 0.0, r, a)
    assert np.all(np.diff(v) <= 0) and np.all(v >= 0)


def test_sinc_zero_crossings_and_window():
    tau, r = 3.0, (4.0, 4.0)
    for k in (1, 2, 3):
        assert abs(film_np.lanczos(float(k), 0.0, r, tau)) < 1e-15             # sin(pi k) = 0
    assert abs(film_np.lanczos(3.0, 0.0, r, tau)) < 1e-15                       # window sinc(x / tau) = 0 at x = tau
    assert film_np.lanczos(4.0001, 0.0, r, tau) == 0.0 and film_np.lanczos(0.0, 0.0, r, tau) == 1.0
    assert abs(film_np.lanczos(0.5, 0.0, r, tau) - (math.sin(math.pi * .5) / (math.pi * .5)) * (math.sin(math.pi / 6) / (math.pi / 6))) < 1e-15


def test_triangle_integral_and_box():
    r = (2.0, 2.0)
    xs = np.linspace(-2, 2, 4001)
    v = film_np.triangle(xs[None, :], xs[:, None], r)
    integral = v.sum() * (xs[1] - xs[0]) ** 2
    assert abs(integral - (r[0] ** 2) * (r[1] ** 2)) < 1e-2                      # (area of a 1-D tent = r^2) squared
    assert np.all(film_np.evaluate("box", xs, 0.0, (0.5, 0.5)) == 1.0)


@pytest.mark.parametrize("name", list(oracle.FILTERS))
def test_oracle_evaluate_matches_closed_forms_off_the_table(orc, name):
    """orc_filter_evaluate at random points inside the support, not only at the 256 table positions"""
    import ctypes as C
    kind, rad, p0, p1 = oracle.FILTERS[name]
    f = oracle.OFilter()
    orc.orc_filter_init(C.byref(f), kind, rad[0], rad[1], p0, p1)
    rng = np.random.default_rng(11)
    pts = rng.uniform(-1, 1, (400, 2)) * np.array(rad)
    got = np.array([orc.orc_filter_evaluate(C.byref(f), float(np.float32(x)), float(np.float32(y))) for x, y in pts])
    want = film_np.evaluate(name, pts[:, 0].astype(np.float32).astype(np.float64), pts[:, 1].astype(np.float32).astype(np.float64), rad, p0, p1)
    assert np.allclose(got, want, rtol=3e-5, atol=1e-6)
