// pbrt_b200.hpp — C++17 host-side mirror of the reference's film / filter / texture interfaces over
// the C ABI in pbrt_b200.h (header-only).
//
// The reference (wathiede/pbrt) is Rust and there is no rustc in the build image, so this is the
// compiled-language stand-in for the Rust shim in rust/: the same type and method names, argument
// meaning and error behaviour as src/core/film.rs, src/core/filter.rs, src/filters/box.rs,
// src/core/texture.rs and src/textures/constant.rs (cited per item), so that tests/cpp/test_film.cpp
// reads like the reference's own tests.  Where the reference panics (`unwrap`, `debug_assert!`,
// `unimplemented!`) these throw pbrt::Panic.
#pragma once

#include <cmath>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "pbrt_b200.h"

namespace pbrt {

using Float = float;  // src/lib.rs:24-44, default features

struct Panic : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline void check(int rc) {
    if (rc != PBRT_OK) throw Panic(std::string("pbrt_b200: ") + pbrt_b200_last_error());
}

// ---------------------------------------------------------------- geometry (2-D only)
struct Point2i {
    int64_t x = 0, y = 0;  // isize
    bool operator==(const Point2i &o) const { return x == o.x && y == o.y; }
};
struct Point2f {
    Float x = 0, y = 0;
    bool operator==(const Point2f &o) const { return x == o.x && y == o.y; }
    Point2f floor() const { return {std::floor(x), std::floor(y)}; }  // point.rs:293-295
    Point2f ceil() const { return {std::ceil(x), std::ceil(y)}; }     // point.rs:306-308
};
using Vector2f = Point2f;

struct Bounds2f {
    Point2f p_min, p_max;
    // Bounds2f::from([[ax,ay],[bx,by]]) sorts each axis (bounds.rs:119-130)
    static Bounds2f from(Point2f a, Point2f b) {
        return {{a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y}, {a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y}};
    }
    bool operator==(const Bounds2f &o) const { return p_min == o.p_min && p_max == o.p_max; }
};

struct Bounds2i {
    Point2i p_min, p_max;
    static Bounds2i from(Point2i a, Point2i b) {  // bounds.rs:119-130
        return {{a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y}, {a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y}};
    }
    bool operator==(const Bounds2i &o) const { return p_min == o.p_min && p_max == o.p_max; }
    int64_t area() const { return (p_max.x - p_min.x) * (p_max.y - p_min.y); }  // bounds.rs:195-198
    bool inside_exclusive(Point2i p) const {                                      // bounds.rs:210-212
        return p.x >= p_min.x && p.x < p_max.x && p.y >= p_min.y && p.y < p_max.y;
    }
    static Bounds2i intersect(const Bounds2i &a, const Bounds2i &b) {  // bounds.rs:244-252, not re-sorted
        return {{a.p_min.x > b.p_min.x ? a.p_min.x : b.p_min.x, a.p_min.y > b.p_min.y ? a.p_min.y : b.p_min.y},
                {a.p_max.x < b.p_max.x ? a.p_max.x : b.p_max.x, a.p_max.y < b.p_max.y ? a.p_max.y : b.p_max.y}};
    }
    // bounds.rs:284-288: row-major, y outer
    template <typename F>
    void for_each(F &&f) const {
        for (int64_t y = p_min.y; y < p_max.y; ++y)
            for (int64_t x = p_min.x; x < p_max.x; ++x) f(Point2i{x, y});
    }
};

namespace detail {
inline void to4(const Bounds2i &b, int32_t out[4]) {
    const int64_t v[4] = {b.p_min.x, b.p_min.y, b.p_max.x, b.p_max.y};
    for (int i = 0; i < 4; ++i) {
        if (v[i] < INT32_MIN || v[i] > INT32_MAX) throw Panic("bounds do not fit 32-bit device coordinates");
        out[i] = (int32_t)v[i];
    }
}
inline Bounds2i from4(const int32_t v[4]) { return {{v[0], v[1]}, {v[2], v[3]}}; }
}  // namespace detail

// ---------------------------------------------------------------- spectrum (RGB)
struct Spectrum {  // RGBSpectrum, spectrum.rs:149-189
    Float c[3] = {0, 0, 0};
    static Spectrum from_rgb(Float r, Float g, Float b) { return {{r, g, b}}; }
    static Spectrum from(Float v) { return {{v, v, v}}; }
    bool operator==(const Spectrum &o) const { return c[0] == o.c[0] && c[1] == o.c[1] && c[2] == o.c[2]; }
    // spectrum.rs:139-145, :166-168 — evaluated left to right, no contraction (build with -ffp-contract=off)
    void to_xyz(Float xyz[3]) const {
        xyz[0] = 0.412453f * c[0] + 0.357580f * c[1] + 0.180423f * c[2];
        xyz[1] = 0.212671f * c[0] + 0.715160f * c[1] + 0.072169f * c[2];
        xyz[2] = 0.019334f * c[0] + 0.119193f * c[1] + 0.950227f * c[2];
    }
};

// ---------------------------------------------------------------- filters
class Filter {  // filter.rs:22-29
  public:
    virtual ~Filter() = default;
    virtual Float evaluate(Point2f p) const = 0;
    virtual Vector2f radius() const = 0;
    virtual Vector2f inv_radius() const = 0;
};

class BoxFilter : public Filter {  // box.rs:30-77
  public:
    explicit BoxFilter(Vector2f radius) : radius_(radius), inv_radius_{1.f / radius.x, 1.f / radius.y} {}
    // create_box_filter(ParamSet): xwidth / ywidth default to 0.5 (box.rs:57-61)
    static BoxFilter create_box_filter(const Float *xwidth = nullptr, const Float *ywidth = nullptr) {
        return BoxFilter({xwidth ? *xwidth : 0.5f, ywidth ? *ywidth : 0.5f});
    }
    Float evaluate(Point2f) const override { return 1.f; }
    Vector2f radius() const override { return radius_; }
    Vector2f inv_radius() const override { return inv_radius_; }

  private:
    Vector2f radius_, inv_radius_;
};

// EXTENSION (no reference parity): the filters src/core/api.rs:954 names but does not implement; formulas
// live in the library (pbrt_filter_create).  kind: PBRT_FILTER_TRIANGLE / GAUSSIAN / MITCHELL / LANCZOS.
class NativeFilter : public Filter {
  public:
    NativeFilter(int kind, Vector2f radius, Float p0 = 0, Float p1 = 0) { check(pbrt_filter_create(kind, radius.x, radius.y, p0, p1, &h_)); }
    ~NativeFilter() override { pbrt_filter_destroy(h_); }
    NativeFilter(const NativeFilter &) = delete;
    Float evaluate(Point2f p) const override { return pbrt_filter_evaluate(h_, p.x, p.y); }
    Vector2f radius() const override { Float r[2]; pbrt_filter_radius(h_, r); return {r[0], r[1]}; }
    Vector2f inv_radius() const override { Float r[2]; pbrt_filter_inv_radius(h_, r); return {r[0], r[1]}; }

  private:
    PbrtFilter *h_ = nullptr;
};

// ---------------------------------------------------------------- film
constexpr int FILTER_TABLE_WIDTH = 16;  // film.rs:34

struct FilmTilePixel {  // film.rs:39-42; this layout IS the rgbw buffer of the C ABI
    Spectrum contrib_sum;
    Float filter_weight_sum = 0;
};
static_assert(sizeof(FilmTilePixel) == 16, "FilmTilePixel must be 4 packed floats");

class Film;

class FilmTile {  // film.rs:428-489
  public:
    Bounds2i get_pixel_bounds() const { return pixel_bounds_; }  // :461-463
    const FilmTilePixel &get_pixel(Point2i p) const { return pixels_[pixel_offset(p)]; }  // :479-482
    FilmTilePixel &get_pixel_mut(Point2i p) { return pixels_[pixel_offset(p)]; }          // :485-488

  private:
    friend class Film;
    size_t pixel_offset(Point2i p) const {  // :465-476 — the reference panics outside the tile
        if (!pixel_bounds_.inside_exclusive(p)) throw Panic("p outside tile pixel bounds");
        const int64_t width = pixel_bounds_.p_max.x - pixel_bounds_.p_min.x;
        return (size_t)((p.x - pixel_bounds_.p_min.x) + (p.y - pixel_bounds_.p_min.y) * width);
    }
    Bounds2i pixel_bounds_;
    std::vector<FilmTilePixel> pixels_;
};

class Film {  // film.rs:59-76
  public:
    Point2i full_resolution;
    std::unique_ptr<Filter> filter;
    Float diagonal_m;
    std::string filename;
    Bounds2i cropped_pixel_bounds;

    // Film::new, film.rs:82-137
    Film(Point2i resolution, Bounds2f crop_window, std::unique_ptr<Filter> filt, Float diagonal_mm, std::string name,
         Float scale, Float max_sample_luminance)
        : full_resolution(resolution), filter(std::move(filt)), diagonal_m(diagonal_mm * 0.001f), filename(std::move(name)) {
        // :113-123 — 256 calls through the trait, (x + .5) * r / 16 in that order
        const Float w = (Float)FILTER_TABLE_WIDTH;
        filter_table_.reserve(FILTER_TABLE_WIDTH * FILTER_TABLE_WIDTH);
        for (int y = 0; y < FILTER_TABLE_WIDTH; ++y)
            for (int x = 0; x < FILTER_TABLE_WIDTH; ++x)
                filter_table_.push_back(filter->evaluate({((Float)x + 0.5f) * filter->radius().x / w,
                                                          ((Float)y + 0.5f) * filter->radius().y / w}));
        const Float crop[4] = {crop_window.p_min.x, crop_window.p_min.y, crop_window.p_max.x, crop_window.p_max.y};
        const Float rad[2] = {filter->radius().x, filter->radius().y};
        check(pbrt_film_create((int32_t)resolution.x, (int32_t)resolution.y, crop, rad, filter_table_.data(), diagonal_mm,
                               scale, max_sample_luminance, &h_));
        int32_t b[4];
        check(pbrt_film_cropped_pixel_bounds(h_, b));
        cropped_pixel_bounds = detail::from4(b);
    }
    ~Film() { pbrt_film_destroy(h_); }
    Film(const Film &) = delete;

    Bounds2i get_sample_bounds() const {  // :166-175
        int32_t b[4];
        check(pbrt_film_get_sample_bounds(h_, b));
        return detail::from4(b);
    }
    Bounds2f get_physical_extent() const {  // :218-227
        Float e[4];
        check(pbrt_film_get_physical_extent(h_, e));
        return {{e[0], e[1]}, {e[2], e[3]}};
    }
    FilmTile get_film_tile(Bounds2i sample_bounds) const {  // :264-281
        int32_t sb[4], tb[4];
        int64_t n = 0;
        detail::to4(sample_bounds, sb);
        check(pbrt_film_tile_bounds(h_, sb, tb, &n));
        FilmTile t;
        t.pixel_bounds_ = detail::from4(tb);
        t.pixels_.assign((size_t)n, FilmTilePixel{});  // FilmTilePixel::default(), max(0, area) pixels (:446)
        return t;
    }
    void merge_film_tile(FilmTile tile) {  // :313-326 — takes the tile by value, as the reference does
        int32_t tb[4];
        detail::to4(tile.pixel_bounds_, tb);
        check(pbrt_film_merge_tile(h_, tb, reinterpret_cast<const Float *>(tile.pixels_.data()), 0));
    }
    void set_image(const std::vector<Spectrum> &) { throw Panic("not implemented"); }  // :329-331 unimplemented!()
    void add_splat(Point2f, Spectrum) { throw Panic("not implemented"); }               // :334-336
    void clear() { throw Panic("not implemented"); }                                    // :386-388
    // the rgb buffer write_image builds (:342-372); the container write is pbrt_b200/imageio.py's job
    std::vector<Float> write_image_rgb(Float splat_scale) const {
        std::vector<Float> rgb(3 * (size_t)(cropped_pixel_bounds.area() > 0 ? cropped_pixel_bounds.area() : 0));
        if (!rgb.empty()) check(pbrt_film_resolve_rgb(h_, splat_scale, rgb.data(), 0));
        return rgb;
    }
    void get_pixel_xyz(Point2i p, Float out[3]) const { check(pbrt_film_get_pixel_xyz(h_, (int32_t)p.x, (int32_t)p.y, out)); }  // :405-410
    PbrtFilm *handle() const { return h_; }

  private:
    PbrtFilm *h_ = nullptr;
    std::vector<Float> filter_table_;
};

// ---------------------------------------------------------------- textures
struct SurfaceInteraction {};  // interaction.rs:22-23 — carries nothing

template <typename T>
class Texture {  // texture.rs:24-30
  public:
    virtual ~Texture() = default;
    virtual T evaluate(const SurfaceInteraction &si) const = 0;
};

template <typename T>
class ConstantTexture : public Texture<T> {  // constant.rs:32-154
  public:
    explicit ConstantTexture(T value) : value_(value) {}
    T evaluate(const SurfaceInteraction &) const override { return value_; }  // :139-141, host side as in the reference
    // n lookups on the device (4 B / 12 B per lookup)
    std::vector<Float> evaluate_batch(uint64_t n) const {
        if constexpr (std::is_same<T, Float>::value) {
            std::vector<Float> out(n);
            if (n) check(pbrt_texture_constant_eval_f32(value_, n, out.data(), 0));
            return out;
        } else {
            std::vector<Float> out(3 * n);
            if (n) check(pbrt_texture_constant_eval_rgb(value_.c, n, out.data(), 0));
            return out;
        }
    }

  private:
    T value_;
};
// constant.rs:61-68 / :96-103 — `value` defaults to 1 / Spectrum::from(1.)
inline ConstantTexture<Float> create_constant_float_texture(const Float *value = nullptr) { return ConstantTexture<Float>(value ? *value : 1.f); }
inline ConstantTexture<Spectrum> create_constant_spectrum_texture(const Spectrum *value = nullptr) {
    return ConstantTexture<Spectrum>(value ? *value : Spectrum::from(1.f));
}

}  // namespace pbrt
