// test_film.cpp — the reference's own film / filter / texture tests and doctests, restated in C++
// against include/pbrt_b200.hpp (which calls the C ABI).  Each check cites the reference assertion.
// Built by __graft_entry__.build(), run on the GPU box by tests/test_gpu_cpp.py.
#include <cstdio>
#include <cstring>

#include "../../include/pbrt_b200.hpp"

using namespace pbrt;

static int failures = 0;
#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } \
    } while (0)

static std::unique_ptr<Filter> box8() { return std::make_unique<BoxFilter>(Vector2f{8.f, 8.f}); }

// src/core/film.rs:151-164
static void get_sample_bounds_doctest() {
    Film film({1920, 1080}, Bounds2f::from({0.25f, 0.25f}, {0.75f, 0.75f}), box8(), 35.0f, "output.png", 1.f, 1.f);
    CHECK(film.get_sample_bounds() == Bounds2i::from({472, 262}, {1448, 818}));
    // src/core/film.rs:241-262
    CHECK(film.get_film_tile(Bounds2i::from({0, 0}, {1920, 1080})).get_pixel_bounds() ==
          Bounds2i::from({1920 / 4, 1080 / 4}, {3 * 1920 / 4, 3 * 1080 / 4}));
    CHECK(film.get_film_tile(Bounds2i::from({500, 500}, {600, 600})).get_pixel_bounds() ==
          Bounds2i::from({492, 492}, {608, 608}));
}

// src/core/film.rs:186-216
static void get_physical_extent_doctest() {
    const Bounds2f want = Bounds2f::from({-0.04f, -0.03f}, {0.04f, 0.03f});
    Film a({800, 600}, Bounds2f::from({0.f, 0.f}, {1.f, 1.f}), box8(), 100.f, "output.png", 1.f, 1.f);
    CHECK(a.get_physical_extent() == want);
    Film b({800, 600}, Bounds2f::from({0.25f, 0.25f}, {0.75f, 0.75f}), box8(), 100.f, "output.png", 1.f, 1.f);
    CHECK(b.get_physical_extent() == want);
}

// src/core/film.rs:296-311
static void merge_degenerate_tile_doctest() {
    Film film({20, 10}, Bounds2f::from({0.f, 0.f}, {1.f, 1.f}), box8(), 35.0f, "output.png", 1.f, 1.f);
    FilmTile left = film.get_film_tile(Bounds2i::from({0, 0}, {10, 10}));
    FilmTile right = film.get_film_tile(Bounds2i::from({10, 0}, {10, 10}));
    film.merge_film_tile(std::move(left));
    film.merge_film_tile(std::move(right));
    Float xyz[3];
    film.get_pixel_xyz({3, 3}, xyz);
    CHECK(xyz[0] == 0.f && xyz[1] == 0.f && xyz[2] == 0.f);
}

static void fill(FilmTile &t, const Spectrum &c) {
    t.get_pixel_bounds().for_each([&](Point2i pt) {
        FilmTilePixel &px = t.get_pixel_mut(pt);
        px.contrib_sum = c;
        px.filter_weight_sum = 1.f;
    });
}

// src/core/film.rs:503-535
static void merge_film_tile_test() {
    Film film({200, 10}, Bounds2f::from({0.f, 0.f}, {1.f, 1.f}), box8(), 35.0f, "merge_film_tile.png", 1.f, 1.f);
    FilmTile left = film.get_film_tile(Bounds2i::from({0, 0}, {100, 10}));
    FilmTile right = film.get_film_tile(Bounds2i::from({100, 0}, {200, 10}));
    const Spectrum green = Spectrum::from_rgb(0.f, 1.f, 0.f), red = Spectrum::from_rgb(1.f, 0.f, 0.f);
    fill(left, green);
    fill(right, red);
    film.merge_film_tile(std::move(left));
    film.merge_film_tile(std::move(right));
    std::vector<Float> rgb = film.write_image_rgb(1.f);
    Float got[3], want[3];
    film.get_pixel_xyz({4, 4}, got);
    green.to_xyz(want);
    CHECK(std::memcmp(got, want, sizeof got) == 0);  // film.rs:533, assert_eq!
    film.get_pixel_xyz({196, 4}, got);
    red.to_xyz(want);
    CHECK(std::memcmp(got, want, sizeof got) == 0);  // film.rs:534
    // the resolved buffer: pure green / pure red where one tile covers, an equal mix where both do
    const Float *g = &rgb[3 * (4 * 200 + 4)], *r = &rgb[3 * (4 * 200 + 196)], *m = &rgb[3 * (4 * 200 + 100)];
    CHECK(std::fabs(g[1] - 1.f) < 1e-5f && std::fabs(g[0]) < 1e-5f);
    CHECK(std::fabs(r[0] - 1.f) < 1e-5f && std::fabs(r[1]) < 1e-5f);
    CHECK(std::fabs(m[0] - 0.5f) < 1e-5f && std::fabs(m[1] - 0.5f) < 1e-5f);
}

// src/core/film.rs:537-571 (a smoke test in the reference)
static void merge_film_tile_rainbow_test() {
    const int64_t WIDTH = 200, HEIGHT = 100;
    Film film({WIDTH, HEIGHT}, Bounds2f::from({0.f, 0.f}, {1.f, 1.f}), box8(), 35.0f, "rainbow.png", 1.f, 1.f);
    auto fill_rainbow = [&](FilmTile &t) {
        t.get_pixel_bounds().for_each([&](Point2i pt) {
            FilmTilePixel &px = t.get_pixel_mut(pt);
            px.contrib_sum = Spectrum::from_rgb((Float)pt.x / (Float)WIDTH, (Float)pt.y / (Float)HEIGHT, (Float)(WIDTH - pt.x) / (Float)WIDTH);
            px.filter_weight_sum = 1.f;
        });
    };
    FilmTile left = film.get_film_tile(Bounds2i::from({0, 0}, {WIDTH / 2, HEIGHT}));
    FilmTile right = film.get_film_tile(Bounds2i::from({WIDTH / 2, 0}, {WIDTH, HEIGHT}));
    fill_rainbow(left);
    fill_rainbow(right);
    film.merge_film_tile(std::move(left));
    film.merge_film_tile(std::move(right));
    std::vector<Float> rgb = film.write_image_rgb(1.f);
    CHECK(rgb.size() == (size_t)(3 * WIDTH * HEIGHT));
    // outside the overlap a pixel resolves back to the colour it was filled with
    const Float *p = &rgb[3 * (50 * WIDTH + 20)];
    CHECK(std::fabs(p[0] - 20.f / 200.f) < 1e-5f && std::fabs(p[1] - 50.f / 100.f) < 1e-5f && std::fabs(p[2] - 180.f / 200.f) < 1e-5f);
}

// src/filters/box.rs:46-55, src/core/api.rs:1058-1064
static void box_filter_doctest() {
    const Float xwidth = 1.f;
    BoxFilter f = BoxFilter::create_box_filter(&xwidth, nullptr);
    CHECK(f.radius() == (Vector2f{1.f, 0.5f}));
    CHECK(f.inv_radius() == (Vector2f{1.f, 2.f}));
    CHECK(f.evaluate({0.3f, -0.2f}) == 1.f);
}

// src/textures/constant.rs:53-59, :84-95, :118-125
static void constant_texture_doctests() {
    SurfaceInteraction si;
    const Float ten = 10.f;
    CHECK(create_constant_float_texture(&ten).evaluate(si) == 10.f);
    CHECK(create_constant_float_texture().evaluate(si) == 1.f);
    const Spectrum red = Spectrum::from_rgb(1.f, 0.f, 0.f);
    CHECK(create_constant_spectrum_texture(&red).evaluate(si) == red);
    CHECK(create_constant_spectrum_texture().evaluate(si) == Spectrum::from(1.f));
    CHECK(ConstantTexture<Float>(10.f).evaluate(si) == 10.f);
    std::vector<Float> many = ConstantTexture<Float>(10.f).evaluate_batch(1000);
    bool all = many.size() == 1000;
    for (Float v : many) all = all && v == 10.f;
    CHECK(all);
    std::vector<Float> rgb = ConstantTexture<Spectrum>(red).evaluate_batch(333);
    all = rgb.size() == 999;
    for (size_t i = 0; i < rgb.size(); ++i) all = all && rgb[i] == (i % 3 == 0 ? 1.f : 0.f);
    CHECK(all);
}

// error behaviour: where the reference panics, the mirror throws
static void panics() {
    Film film({20, 10}, Bounds2f::from({0.f, 0.f}, {1.f, 1.f}), box8(), 35.0f, "output.png", 1.f, 1.f);
    bool threw = false;
    Float xyz[3];
    try { film.get_pixel_xyz({20, 0}, xyz); } catch (const Panic &) { threw = true; }  // film.rs:391-396
    CHECK(threw);
    threw = false;
    FilmTile t = film.get_film_tile(Bounds2i::from({0, 0}, {4, 4}));
    try { t.get_pixel({100, 100}); } catch (const Panic &) { threw = true; }  // film.rs:466-471
    CHECK(threw);
    threw = false;
    try { film.clear(); } catch (const Panic &) { threw = true; }  // film.rs:386-388 unimplemented!()
    CHECK(threw);
}

int main() {
    if (pbrt_b200_init(0) != PBRT_OK) {
        std::printf("no device: %s\n", pbrt_b200_last_error());
        return 2;
    }
    get_sample_bounds_doctest();
    get_physical_extent_doctest();
    merge_degenerate_tile_doctest();
    merge_film_tile_test();
    merge_film_tile_rainbow_test();
    box_filter_doctest();
    constant_texture_doctests();
    panics();
    std::printf(failures ? "%d check(s) failed\n" : "all reference tests passed (%d failures)\n", failures);
    return failures ? 1 : 0;
}
