/*
 * pbrt_oracle.c — CPU restatement of the wathiede/pbrt film / filter / texture path.
 * TEST INFRASTRUCTURE ONLY — see pbrt_oracle.h for the rules and the parity status.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared (oracle/Makefile).
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 */
#include "pbrt_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

/* ------------------------------------------------------------------ prelude */

/* src/lib.rs:93-99 */
float orc_gamma_correct(float v) {
    if (v <= 0.0031308f) return 12.92f * v;
    return 1.055f * powf(v, 1.f / 2.4f) - 0.055f;
}

/* src/lib.rs:115-126 */
float orc_clamp_f(float v, float lo, float hi) {
    if (v < lo) return lo;
    if (v > hi) return hi;
    return v;
}
int64_t orc_clamp_i(int64_t v, int64_t lo, int64_t hi) {
    if (v < lo) return lo;
    if (v > hi) return hi;
    return v;
}

/* src/core/imageio.rs:66-68; `as u8` saturates and maps NaN to 0 */
uint8_t orc_to_byte(float v) {
    float c = orc_clamp_f(255.f * orc_gamma_correct(v) + 0.5f, 0.f, 255.f);
    if (!(c == c)) return 0;
    return (uint8_t)c;
}

/* ------------------------------------------------------------------ geometry */

/* src/core/geometry/point.rs:323-330 — Rust float->int `as` casts saturate, NaN -> 0 */
int64_t orc_f2i(float v) {
    if (!(v == v)) return 0;
    if (v >= 9223372036854775808.f) return INT64_MAX;
    if (v <= -9223372036854775808.f) return INT64_MIN;
    return (int64_t)v;
}

/* src/core/geometry/bounds.rs:119-130 — the From<[Point2;2]> ctor sorts each axis */
orc_bounds2i orc_bounds2i_from_points(int64_t ax, int64_t ay, int64_t bx, int64_t by) {
    orc_bounds2i b;
    b.x0 = ax < bx ? ax : bx;
    b.y0 = ay < by ? ay : by;
    b.x1 = ax > bx ? ax : bx;
    b.y1 = ay > by ? ay : by;
    return b;
}

/* src/core/geometry/bounds.rs:244-252 — result is NOT re-sorted, may be inverted */
orc_bounds2i orc_bounds2i_intersect(orc_bounds2i a, orc_bounds2i b) {
    orc_bounds2i r;
    r.x0 = a.x0 > b.x0 ? a.x0 : b.x0;
    r.y0 = a.y0 > b.y0 ? a.y0 : b.y0;
    r.x1 = a.x1 < b.x1 ? a.x1 : b.x1;
    r.y1 = a.y1 < b.y1 ? a.y1 : b.y1;
    return r;
}

/* src/core/geometry/bounds.rs:195-198 — plain product, positive for a doubly inverted box */
int64_t orc_bounds2i_area(orc_bounds2i b) { return (b.x1 - b.x0) * (b.y1 - b.y0); }

/* src/core/geometry/bounds.rs:210-212 */
int orc_bounds2i_inside_exclusive(orc_bounds2i b, int64_t x, int64_t y) {
    return x >= b.x0 && x < b.x1 && y >= b.y0 && y < b.y1;
}

/* src/core/geometry/bounds.rs:284-288 — y outer, x inner; empty ranges yield nothing */
int64_t orc_bounds2i_iter(orc_bounds2i b, int64_t *xy_out, int64_t cap) {
    int64_t n = 0;
    for (int64_t y = b.y0; y < b.y1; ++y)
        for (int64_t x = b.x0; x < b.x1; ++x) {
            if (n < cap) { xy_out[2 * n] = x; xy_out[2 * n + 1] = y; }
            ++n;
        }
    return n;
}

void orc_point2f_floor(const float in[2], float out[2]) { out[0] = floorf(in[0]); out[1] = floorf(in[1]); }
void orc_point2f_ceil(const float in[2], float out[2]) { out[0] = ceilf(in[0]); out[1] = ceilf(in[1]); }

/* ------------------------------------------------------------------ spectrum */

/* src/core/spectrum.rs:139-145 */
void orc_rgb_to_xyz(const float rgb[3], float xyz[3]) {
    float r = rgb[0], g = rgb[1], b = rgb[2];
    xyz[0] = 0.412453f * r + 0.357580f * g + 0.180423f * b;
    xyz[1] = 0.212671f * r + 0.715160f * g + 0.072169f * b;
    xyz[2] = 0.019334f * r + 0.119193f * g + 0.950227f * b;
}

/* src/core/spectrum.rs:129-135 */
void orc_xyz_to_rgb(const float xyz[3], float rgb[3]) {
    float x = xyz[0], y = xyz[1], z = xyz[2];
    rgb[0] = 3.240479f * x - 1.537150f * y - 0.498535f * z;
    rgb[1] = -0.969256f * x + 1.875991f * y + 0.041556f * z;
    rgb[2] = 0.055648f * x - 0.204043f * y + 1.057311f * z;
}

/* ------------------------------------------------------------------ filters */

/* src/filters/box.rs:37-42 for the radius / inv_radius pair; parameters are ext */
void orc_filter_init(orc_filter *f, int kind, float rx, float ry, float p0, float p1) {
    memset(f, 0, sizeof *f);
    f->kind = kind;
    f->radius[0] = rx; f->radius[1] = ry;
    f->inv_radius[0] = 1.f / rx; f->inv_radius[1] = 1.f / ry;
    f->p0 = p0; f->p1 = p1;
    if (kind == ORC_FILTER_GAUSSIAN) {
        f->exp_x = expf(-p0 * rx * rx);
        f->exp_y = expf(-p0 * ry * ry);
    }
}

/* src/filters/box.rs:57-61 */
void orc_box_filter_create(orc_filter *f, int has_xwidth, float xwidth, int has_ywidth, float ywidth) {
    orc_filter_init(f, ORC_FILTER_BOX, has_xwidth ? xwidth : 0.5f, has_ywidth ? ywidth : 0.5f, 0.f, 0.f);
}

static float ext_gaussian_1d(float alpha, float d, float expv) {
    float g = expf(-alpha * d * d) - expv;
    return g > 0.f ? g : 0.f;
}
static float ext_mitchell_1d(float B, float C, float x) {
    x = fabsf(2.f * x);
    if (x > 1.f)
        return ((-B - 6.f * C) * x * x * x + (6.f * B + 30.f * C) * x * x + (-12.f * B - 48.f * C) * x +
                (8.f * B + 24.f * C)) * (1.f / 6.f);
    return ((12.f - 9.f * B - 6.f * C) * x * x * x + (-18.f + 12.f * B + 6.f * C) * x * x + (6.f - 2.f * B)) *
           (1.f / 6.f);
}
static const float EXT_PI = 3.14159265358979323846f;
static float ext_sinc(float x) {
    x = fabsf(x);
    if (x < 1e-5f) return 1.f;
    return sinf(EXT_PI * x) / (EXT_PI * x);
}
static float ext_windowed_sinc(float x, float radius, float tau) {
    x = fabsf(x);
    if (x > radius) return 0.f;
    float lanczos = ext_sinc(x / tau);
    return ext_sinc(x) * lanczos;
}

/* box: src/filters/box.rs:66-68.  others: ext, pbrt-v3 7.8 */
float orc_filter_evaluate(const orc_filter *f, float px, float py) {
    switch (f->kind) {
    case ORC_FILTER_BOX:
        return 1.f;
    case ORC_FILTER_TRIANGLE: {
        float a = f->radius[0] - fabsf(px), b = f->radius[1] - fabsf(py);
        return (a > 0.f ? a : 0.f) * (b > 0.f ? b : 0.f);
    }
    case ORC_FILTER_GAUSSIAN:
        return ext_gaussian_1d(f->p0, px, f->exp_x) * ext_gaussian_1d(f->p0, py, f->exp_y);
    case ORC_FILTER_MITCHELL:
        return ext_mitchell_1d(f->p0, f->p1, px * f->inv_radius[0]) *
               ext_mitchell_1d(f->p0, f->p1, py * f->inv_radius[1]);
    case ORC_FILTER_LANCZOS:
        return ext_windowed_sinc(px, f->radius[0], f->p0) * ext_windowed_sinc(py, f->radius[1], f->p0);
    }
    return 0.f;
}

/* src/core/film.rs:113-123 — (x + .5) * r / 16, in that order */
void orc_filter_table(const orc_filter *f, float table[256]) {
    const float w = (float)ORC_FILTER_TABLE_WIDTH;
    int k = 0;
    for (int y = 0; y < ORC_FILTER_TABLE_WIDTH; ++y)
        for (int x = 0; x < ORC_FILTER_TABLE_WIDTH; ++x) {
            float fx = ((float)x + 0.5f) * f->radius[0] / w;
            float fy = ((float)y + 0.5f) * f->radius[1] / w;
            table[k++] = orc_filter_evaluate(f, fx, fy);
        }
}

/* ------------------------------------------------------------------ film */

struct orc_film {
    int64_t xres, yres;
    float crop[4];
    float radius[2];
    float diagonal_m;
    float scale;
    float max_sample_luminance;
    orc_bounds2i cropped;
    orc_pixel *pixels;
    int64_t npixels;
    float table[256];
};

struct orc_tile {
    orc_bounds2i pixel_bounds;
    float filter_radius[2], inv_filter_radius[2];
    const float *filter_table;
    int filter_table_size;
    float max_sample_luminance;
    orc_tile_pixel *pixels;
    int64_t npixels;
};

/* src/core/film.rs:82-137 */
orc_film *orc_film_new(int64_t xres, int64_t yres, const float crop[4], const float radius[2],
                       const float table[256], float diagonal_mm, float scale, float max_sample_luminance) {
    orc_film *f = (orc_film *)calloc(1, sizeof *f);
    f->xres = xres; f->yres = yres;
    memcpy(f->crop, crop, sizeof f->crop);
    f->radius[0] = radius[0]; f->radius[1] = radius[1];
    f->diagonal_m = diagonal_mm * 0.001f;
    f->scale = scale;
    f->max_sample_luminance = max_sample_luminance;
    f->cropped = orc_bounds2i_from_points(
        orc_f2i(ceilf((float)xres * crop[0])), orc_f2i(ceilf((float)yres * crop[1])),
        orc_f2i(ceilf((float)xres * crop[2])), orc_f2i(ceilf((float)yres * crop[3])));
    int64_t area = orc_bounds2i_area(f->cropped);
    f->npixels = area > 0 ? area : 0; /* (0..area) is empty for area <= 0 */
    f->pixels = (orc_pixel *)calloc((size_t)(f->npixels ? f->npixels : 1), sizeof(orc_pixel));
    memcpy(f->table, table, sizeof f->table);
    return f;
}

void orc_film_free(orc_film *f) {
    if (!f) return;
    free(f->pixels);
    free(f);
}

orc_bounds2i orc_film_cropped_pixel_bounds(const orc_film *f) { return f->cropped; }
orc_pixel *orc_film_pixels(orc_film *f) { return f->pixels; }
int64_t orc_film_pixel_count(const orc_film *f) { return f->npixels; }
const float *orc_film_table(const orc_film *f) { return f->table; }

/* src/core/film.rs:166-175 — Bounds2f::from sorts, then per-component `as isize` */
orc_bounds2i orc_film_get_sample_bounds(const orc_film *f) {
    float ax = floorf((float)f->cropped.x0 + 0.5f - f->radius[0]);
    float ay = floorf((float)f->cropped.y0 + 0.5f - f->radius[1]);
    float bx = ceilf((float)f->cropped.x1 - 0.5f + f->radius[0]);
    float by = ceilf((float)f->cropped.y1 - 0.5f + f->radius[1]);
    orc_bounds2i r;
    r.x0 = orc_f2i(ax < bx ? ax : bx);
    r.y0 = orc_f2i(ay < by ? ay : by);
    r.x1 = orc_f2i(ax > bx ? ax : bx);
    r.y1 = orc_f2i(ay > by ? ay : by);
    return r;
}

/* src/core/film.rs:218-227 */
orc_bounds2f orc_film_get_physical_extent(const orc_film *f) {
    float aspect = (float)f->yres / (float)f->xres;
    float x = sqrtf(f->diagonal_m * f->diagonal_m / (1.f + aspect * aspect));
    float y = aspect * x;
    float ax = -x / 2.f, ay = -y / 2.f, bx = x / 2.f, by = y / 2.f;
    orc_bounds2f r; /* [Point2f;2].into() sorts */
    r.x0 = ax < bx ? ax : bx; r.y0 = ay < by ? ay : by;
    r.x1 = ax > bx ? ax : bx; r.y1 = ay > by ? ay : by;
    return r;
}

/* src/core/film.rs:264-273 */
orc_bounds2i orc_film_tile_bounds(const orc_film *f, orc_bounds2i sb) {
    int64_t p0x = orc_f2i(ceilf((float)sb.x0 - 0.5f - f->radius[0]));
    int64_t p0y = orc_f2i(ceilf((float)sb.y0 - 0.5f - f->radius[1]));
    int64_t p1x = orc_f2i(floorf((float)sb.x1 - 0.5f + f->radius[0]) + 1.f);
    int64_t p1y = orc_f2i(floorf((float)sb.y1 - 0.5f + f->radius[1]) + 1.f);
    return orc_bounds2i_intersect(orc_bounds2i_from_points(p0x, p0y, p1x, p1y), f->cropped);
}

/* src/core/film.rs:439-456 */
static orc_tile *tile_new(orc_bounds2i pb, const float radius[2], const float *table, int table_size,
                          float max_sample_luminance) {
    orc_tile *t = (orc_tile *)calloc(1, sizeof *t);
    int64_t area = orc_bounds2i_area(pb);
    t->pixel_bounds = pb;
    t->npixels = area > 0 ? area : 0; /* 0.max(area): a doubly inverted box keeps a positive area */
    t->filter_radius[0] = radius[0]; t->filter_radius[1] = radius[1];
    t->inv_filter_radius[0] = 1.f / radius[0]; t->inv_filter_radius[1] = 1.f / radius[1];
    t->filter_table = table;
    t->filter_table_size = table_size;
    t->max_sample_luminance = max_sample_luminance;
    t->pixels = (orc_tile_pixel *)calloc((size_t)(t->npixels ? t->npixels : 1), sizeof(orc_tile_pixel));
    return t;
}

/* src/core/film.rs:264-281 */
orc_tile *orc_film_get_film_tile(const orc_film *f, orc_bounds2i sample_bounds) {
    return tile_new(orc_film_tile_bounds(f, sample_bounds), f->radius, f->table, ORC_FILTER_TABLE_WIDTH,
                    f->max_sample_luminance);
}

void orc_tile_free(orc_tile *t) {
    if (!t) return;
    free(t->pixels);
    free(t);
}
orc_bounds2i orc_tile_get_pixel_bounds(const orc_tile *t) { return t->pixel_bounds; }
int64_t orc_tile_pixel_count(const orc_tile *t) { return t->npixels; }
orc_tile_pixel *orc_tile_pixels(orc_tile *t) { return t->pixels; }

/* src/core/film.rs:465-476 */
static int64_t tile_pixel_offset(const orc_tile *t, int64_t x, int64_t y) {
    int64_t width = t->pixel_bounds.x1 - t->pixel_bounds.x0;
    return (x - t->pixel_bounds.x0) + (y - t->pixel_bounds.y0) * width;
}
/* src/core/film.rs:479-488 */
orc_tile_pixel *orc_tile_get_pixel(orc_tile *t, int64_t x, int64_t y) {
    if (!orc_bounds2i_inside_exclusive(t->pixel_bounds, x, y)) return NULL;
    return &t->pixels[tile_pixel_offset(t, x, y)];
}

/* src/core/film.rs:390-402 */
static int64_t film_pixel_offset(const orc_film *f, int64_t x, int64_t y) {
    int64_t width = f->cropped.x1 - f->cropped.x0;
    return (x - f->cropped.x0) + (y - f->cropped.y0) * width;
}

/* src/core/film.rs:313-326 */
void orc_film_merge_film_tile(orc_film *f, orc_tile *t) {
    orc_bounds2i pb = t->pixel_bounds;
    for (int64_t y = pb.y0; y < pb.y1; ++y)
        for (int64_t x = pb.x0; x < pb.x1; ++x) {
            const orc_tile_pixel *tp = &t->pixels[tile_pixel_offset(t, x, y)];
            orc_pixel *mp = &f->pixels[film_pixel_offset(f, x, y)];
            float xyz[3];
            orc_rgb_to_xyz(tp->contrib_sum, xyz);
            for (int i = 0; i < 3; ++i) mp->xyz[i] += xyz[i];
            mp->filter_weight_sum += tp->filter_weight_sum;
        }
    orc_tile_free(t);
}

/* Rust f32::max(x, 0.): NaN -> 0.  The sign of max(-0., 0.) is unspecified in Rust; +0 here. */
static float max0(float v) { return v > 0.f ? v : 0.f; }

/* src/core/film.rs:340-372 */
void orc_film_write_image_rgb(const orc_film *f, float splat_scale, float *rgb) {
    int64_t offset = 0;
    for (int64_t y = f->cropped.y0; y < f->cropped.y1; ++y)
        for (int64_t x = f->cropped.x0; x < f->cropped.x1; ++x, ++offset) {
            const orc_pixel *p = &f->pixels[film_pixel_offset(f, x, y)];
            float c[3];
            orc_xyz_to_rgb(p->xyz, c);
            float w = p->filter_weight_sum;
            if (w != 0.f) {
                float inv = 1.f / w;
                c[0] = max0(c[0] * inv);
                c[1] = max0(c[1] * inv);
                c[2] = max0(c[2] * inv);
            }
            float s[3];
            orc_xyz_to_rgb(p->splat_xyz, s);
            for (int i = 0; i < 3; ++i) {
                c[i] += splat_scale * s[i];
                c[i] *= f->scale;
                rgb[3 * offset + i] = c[i];
            }
        }
}

/* src/core/film.rs:405-410 */
void orc_film_get_pixel_xyz(const orc_film *f, int64_t x, int64_t y, float out[3]) {
    const orc_pixel *p = &f->pixels[film_pixel_offset(f, x, y)];
    out[0] = p->xyz[0]; out[1] = p->xyz[1]; out[2] = p->xyz[2];
}

/* ------------------------------------------------------------------ textures */

/* src/textures/constant.rs:61-68 (default 1.), :139-141 */
void orc_constant_texture_eval_f32(int has_value, float value, uint64_t n, float *out) {
    float v = has_value ? value : 1.f;
    for (uint64_t i = 0; i < n; ++i) out[i] = v;
}
/* src/textures/constant.rs:96-103 (default Spectrum::from(1.)), :139-141 */
void orc_constant_texture_eval_rgb(int has_value, const float value[3], uint64_t n, float *out) {
    float v[3] = {1.f, 1.f, 1.f};
    if (has_value) { v[0] = value[0]; v[1] = value[1]; v[2] = value[2]; }
    for (uint64_t i = 0; i < n; ++i) { out[3 * i] = v[0]; out[3 * i + 1] = v[1]; out[3 * i + 2] = v[2]; }
}

/* src/core/mipmap.rs:43-52 */
void orc_weight_lut(float out[128]) {
    const float alpha = 2.f;
    for (int i = 0; i < 128; ++i) {
        float r2 = (float)i / (float)(128 - 1);
        out[i] = expf(-alpha * r2) - expf(-alpha);
    }
}

/* ------------------------------------------------------------------ rng */

#define PCG32_DEFAULT_STATE 0x853c49e6748fea9bULL
#define PCG32_DEFAULT_STREAM 0xda3e39cb94b95bdbULL
#define PCG32_MULT 0x5851f42d4c957f2dULL

/* src/core/rng.rs:35-42 */
void orc_rng_default(orc_rng *r) { r->state = PCG32_DEFAULT_STATE; r->inc = PCG32_DEFAULT_STREAM; }

/* src/core/rng.rs:62-76 */
uint32_t orc_rng_uniform_u32(orc_rng *r) {
    uint64_t old = r->state;
    r->state = old * PCG32_MULT + r->inc;
    uint32_t xorshifted = (uint32_t)(((old >> 18) ^ old) >> 27);
    uint32_t rot = (uint32_t)(old >> 59);
    return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
}

/* src/core/rng.rs:53-59 */
void orc_rng_set_sequence(orc_rng *r, uint64_t sequence_index) {
    r->state = 0;
    r->inc = (sequence_index << 1) | 1;
    orc_rng_uniform_u32(r);
    r->state += PCG32_DEFAULT_STATE;
    orc_rng_uniform_u32(r);
}

/* src/core/rng.rs:79-87 */
uint32_t orc_rng_uniform_u32_threshold(orc_rng *r, uint32_t b) {
    uint32_t threshold = (~b + 1u) % b;
    for (;;) {
        uint32_t v = orc_rng_uniform_u32(r);
        if (v >= threshold) return v % b;
    }
}

/* src/core/rng.rs:91-93; ONE_MINUS_EPSILON = 1 - f32::EPSILON (:19) */
float orc_rng_uniform_float(orc_rng *r) {
    const float one_minus_eps = 1.f - 1.1920929e-07f;
    float v = (float)orc_rng_uniform_u32(r) * 2.3283064365386963e-10f;
    return one_minus_eps < v ? one_minus_eps : v;
}

/* ------------------------------------------------------------------ imageio */

/* src/core/imageio.rs:186-213 (little-endian host: scale -1) */
size_t orc_pfm_encode(const float *rgb, int64_t width, int64_t height, uint8_t *out, size_t cap) {
    char hdr[64];
    int h = snprintf(hdr, sizeof hdr, "PF\n%lld %lld\n-1\n", (long long)width, (long long)height);
    size_t need = (size_t)h + (size_t)(width * height * 3) * 4;
    if (!out || cap < need) return need;
    memcpy(out, hdr, (size_t)h);
    uint8_t *p = out + h;
    for (int64_t y = height - 1; y >= 0; --y) {
        memcpy(p, rgb + y * width * 3, (size_t)width * 12);
        p += width * 12;
    }
    return need;
}

/* =================================================================== Tier 2 — extension, parity unpinned */

/* pbrt-v3 RGBSpectrum::y(); the weights are row 2 of rgb_to_xyz (src/core/spectrum.rs:142) */
static float ext_lum(const float c[3]) { return 0.212671f * c[0] + 0.715160f * c[1] + 0.072169f * c[2]; }

/* pbrt-v3 FilmTile::AddSample over the fields at src/core/film.rs:428-436 */
void orc_ext_tile_add_sample(orc_tile *t, float px, float py, const float Lin[3], float sample_weight) {
    float L[3] = {Lin[0], Lin[1], Lin[2]};
    float ly = ext_lum(L);
    if (ly > t->max_sample_luminance) {
        float s = t->max_sample_luminance / ly;
        L[0] *= s; L[1] *= s; L[2] *= s;
    }
    float dx = px - 0.5f, dy = py - 0.5f;
    int64_t p0x = orc_f2i(ceilf(dx - t->filter_radius[0]));
    int64_t p0y = orc_f2i(ceilf(dy - t->filter_radius[1]));
    int64_t p1x = orc_f2i(floorf(dx + t->filter_radius[0])) + 1;
    int64_t p1y = orc_f2i(floorf(dy + t->filter_radius[1])) + 1;
    if (p0x < t->pixel_bounds.x0) p0x = t->pixel_bounds.x0;
    if (p0y < t->pixel_bounds.y0) p0y = t->pixel_bounds.y0;
    if (p1x > t->pixel_bounds.x1) p1x = t->pixel_bounds.x1;
    if (p1y > t->pixel_bounds.y1) p1y = t->pixel_bounds.y1;
    if (p1x <= p0x || p1y <= p0y) return;

    const int ts = t->filter_table_size;
    int ifx[64], ify[64];
    int nx = (int)(p1x - p0x), ny = (int)(p1y - p0y);
    if (nx > 64 || ny > 64) return; /* radius > 31 px is outside what this restatement covers */
    for (int64_t x = p0x; x < p1x; ++x) {
        float fx = fabsf(((float)x - dx) * t->inv_filter_radius[0] * (float)ts);
        int i = fx >= (float)ts ? ts - 1 : (int)floorf(fx);
        ifx[x - p0x] = i < ts - 1 ? i : ts - 1;
    }
    for (int64_t y = p0y; y < p1y; ++y) {
        float fy = fabsf(((float)y - dy) * t->inv_filter_radius[1] * (float)ts);
        int i = fy >= (float)ts ? ts - 1 : (int)floorf(fy);
        ify[y - p0y] = i < ts - 1 ? i : ts - 1;
    }
    for (int64_t y = p0y; y < p1y; ++y)
        for (int64_t x = p0x; x < p1x; ++x) {
            float w = t->filter_table[ify[y - p0y] * ts + ifx[x - p0x]];
            orc_tile_pixel *p = &t->pixels[tile_pixel_offset(t, x, y)];
            p->contrib_sum[0] += L[0] * sample_weight * w;
            p->contrib_sum[1] += L[1] * sample_weight * w;
            p->contrib_sum[2] += L[2] * sample_weight * w;
            p->filter_weight_sum += w;
        }
}

void orc_ext_tile_add_samples(orc_tile *t, uint64_t n, const float *xy, const float *rgbw) {
    for (uint64_t i = 0; i < n; ++i)
        orc_ext_tile_add_sample(t, xy[2 * i], xy[2 * i + 1], &rgbw[4 * i], rgbw[4 * i + 3]);
}

/* pbrt-v3 Film::AddSplat on Pixel::splat_xyz (src/core/film.rs:50-52, :334-336) */
void orc_ext_film_add_splat(orc_film *f, float px, float py, const float vin[3]) {
    float v[3] = {vin[0], vin[1], vin[2]};
    if (v[0] != v[0] || v[1] != v[1] || v[2] != v[2]) return;
    float ly = ext_lum(v);
    if (ly < 0.f) return;
    if (isinf(ly)) return;
    int64_t ix = orc_f2i(px), iy = orc_f2i(py);
    if (!orc_bounds2i_inside_exclusive(f->cropped, ix, iy)) return;
    if (ly > f->max_sample_luminance) {
        float s = f->max_sample_luminance / ly;
        v[0] *= s; v[1] *= s; v[2] *= s;
    }
    float xyz[3];
    orc_rgb_to_xyz(v, xyz);
    orc_pixel *p = &f->pixels[film_pixel_offset(f, ix, iy)];
    for (int i = 0; i < 3; ++i) p->splat_xyz[i] += xyz[i];
}

/* SURVEY.md App. C */
void orc_ext_synth_samples(orc_bounds2i b, int spp, uint64_t seed, float *xy, float *rgbw) {
    int n = 1;
    while (n * n < spp) ++n;
    const float fn = (float)n;
    int64_t W = b.x1 - b.x0;
    for (int64_t py = b.y0; py < b.y1; ++py)
        for (int64_t px = b.x0; px < b.x1; ++px) {
            uint64_t i = (uint64_t)((py - b.y0) * W + (px - b.x0));
            orc_rng rng;
            orc_rng_default(&rng);
            orc_rng_set_sequence(&rng, (seed << 32) + i);
            for (int s = 0; s < spp; ++s) {
                int sx = s % n, sy = s / n;
                float jx = orc_rng_uniform_float(&rng);
                float jy = orc_rng_uniform_float(&rng);
                float r = orc_rng_uniform_float(&rng);
                float g = orc_rng_uniform_float(&rng);
                float bl = orc_rng_uniform_float(&rng);
                uint64_t k = i * (uint64_t)spp + (uint64_t)s;
                xy[2 * k] = (float)px + ((float)sx + jx) / fn;
                xy[2 * k + 1] = (float)py + ((float)sy + jy) / fn;
                rgbw[4 * k] = r; rgbw[4 * k + 1] = g; rgbw[4 * k + 2] = bl; rgbw[4 * k + 3] = 1.f;
            }
        }
}

void orc_ext_synth_tile_fill(orc_tile *t, uint64_t seed, uint64_t tile_index) {
    for (int64_t p = 0; p < t->npixels; ++p) {
        orc_rng rng;
        orc_rng_default(&rng);
        orc_rng_set_sequence(&rng, (seed << 32) + (tile_index << 20) + (uint64_t)p);
        t->pixels[p].contrib_sum[0] = orc_rng_uniform_float(&rng);
        t->pixels[p].contrib_sum[1] = orc_rng_uniform_float(&rng);
        t->pixels[p].contrib_sum[2] = orc_rng_uniform_float(&rng);
        t->pixels[p].filter_weight_sum = 1.f;
    }
}

typedef struct {
    orc_tile *band;          /* sub-tile: same fields, pixel rows [y0,y1) of the full tile */
    orc_bounds2i sb;         /* sample bounds of the whole pass */
    int spp;
    const float *xy, *rgbw;
    int64_t sy0, sy1;        /* nominal sample rows whose footprint can reach the band */
} pass_job;

static void *pass_worker(void *arg) {
    pass_job *j = (pass_job *)arg;
    int64_t W = j->sb.x1 - j->sb.x0;
    for (int64_t y = j->sy0; y < j->sy1; ++y) {
        uint64_t first = (uint64_t)((y - j->sb.y0) * W) * (uint64_t)j->spp;
        orc_ext_tile_add_samples(j->band, (uint64_t)W * (uint64_t)j->spp, j->xy + 2 * first, j->rgbw + 4 * first);
    }
    return NULL;
}

void orc_ext_film_add_samples_pass(orc_film *f, orc_bounds2i sb, int spp, const float *xy, const float *rgbw,
                                   int threads) {
    orc_tile *tile = orc_film_get_film_tile(f, sb);
    orc_bounds2i pb = tile->pixel_bounds;
    int64_t H = pb.y1 - pb.y0, W = sb.x1 - sb.x0;
    if (tile->npixels == 0 || H <= 0 || W <= 0 || sb.y1 <= sb.y0) {
        orc_film_merge_film_tile(f, tile);
        return;
    }
    if (threads < 1) threads = 1;
    if (threads > H) threads = (int)H;
    if (threads == 1) {
        orc_ext_tile_add_samples(tile, (uint64_t)W * (uint64_t)(sb.y1 - sb.y0) * (uint64_t)spp, xy, rgbw);
        orc_film_merge_film_tile(f, tile);
        return;
    }
    int64_t halo = (int64_t)ceilf(f->radius[1] + 0.5f) + 1;
    pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof *th);
    pass_job *jobs = (pass_job *)calloc((size_t)threads, sizeof *jobs);
    orc_tile *bands = (orc_tile *)calloc((size_t)threads, sizeof *bands);
    int64_t tw = pb.x1 - pb.x0;
    for (int k = 0; k < threads; ++k) {
        int64_t y0 = pb.y0 + H * k / threads, y1 = pb.y0 + H * (k + 1) / threads;
        bands[k] = *tile;
        bands[k].pixel_bounds.y0 = y0;
        bands[k].pixel_bounds.y1 = y1;
        bands[k].pixels = tile->pixels + (y0 - pb.y0) * tw; /* rows are contiguous: alias, no copy */
        bands[k].npixels = (y1 - y0) * tw;
        jobs[k].band = &bands[k];
        jobs[k].sb = sb; jobs[k].spp = spp; jobs[k].xy = xy; jobs[k].rgbw = rgbw;
        jobs[k].sy0 = y0 - halo < sb.y0 ? sb.y0 : y0 - halo;
        jobs[k].sy1 = y1 + halo > sb.y1 ? sb.y1 : y1 + halo;
        if (jobs[k].sy1 < jobs[k].sy0) jobs[k].sy1 = jobs[k].sy0;
        pthread_create(&th[k], NULL, pass_worker, &jobs[k]);
    }
    for (int k = 0; k < threads; ++k) pthread_join(th[k], NULL);
    free(th); free(jobs); free(bands);
    orc_film_merge_film_tile(f, tile);
}
