/*
 * pbrt_oracle.h — CPU restatement of the film / filter / texture path of wathiede/pbrt.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and
 * only as the checker (or as the timed CPU arm), never as a fallback for the CUDA path.
 *
 * Parity status
 *   Tier 1 (reference-backed): pinned against every known answer the reference's own tests and
 *     doctests hold for this path (tests/test_oracle_kat.py).  The reference
 *     is Rust and there is no rustc in the build image, so it cannot be executed here; the
 *     pinning is against its asserted values, not against a live run.
 *   Tier 2 (ext_* symbols: add_sample, add_splat, triangle/gaussian/mitchell/lanczos filters):
 *     PARITY UNPINNED.  The reference declares but does not implement them
 *     (src/core/film.rs:428-436 unused fields, :334 unimplemented!, src/core/api.rs:954-956);
 *     they restate the published pbrt-v3 algorithm the reference is porting (README.md:12-13).
 *
 * All arithmetic is IEEE binary32, evaluated left to right with no contraction
 * (build with -ffp-contract=off), matching what rustc emits for the cited lines.
 */
#ifndef PBRT_ORACLE_H
#define PBRT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_FILTER_TABLE_WIDTH 16 /* src/core/film.rs:34 */

/* src/core/geometry/bounds.rs — Bounds2i as {x0,y0,x1,y1}; isize on the host is i64 */
typedef struct { int64_t x0, y0, x1, y1; } orc_bounds2i;
typedef struct { float x0, y0, x1, y1; } orc_bounds2f;

/* src/core/film.rs:47-55 — 7 floats, unpadded */
typedef struct { float xyz[3]; float filter_weight_sum; float splat_xyz[3]; } orc_pixel;
/* src/core/film.rs:39-42 */
typedef struct { float contrib_sum[3]; float filter_weight_sum; } orc_tile_pixel;

typedef struct orc_film orc_film;
typedef struct orc_tile orc_tile;

/* ---- numeric prelude: src/lib.rs:93-126 ---- */
float orc_gamma_correct(float v);
float orc_clamp_f(float v, float lo, float hi);
int64_t orc_clamp_i(int64_t v, int64_t lo, int64_t hi);
uint8_t orc_to_byte(float v); /* src/core/imageio.rs:66-68 */

/* ---- geometry ---- */
int64_t orc_f2i(float v);                       /* Rust `as isize`: saturating, NaN -> 0 */
orc_bounds2i orc_bounds2i_from_points(int64_t ax, int64_t ay, int64_t bx, int64_t by); /* bounds.rs:119-130 */
orc_bounds2i orc_bounds2i_intersect(orc_bounds2i a, orc_bounds2i b);                   /* bounds.rs:244-252 */
int64_t orc_bounds2i_area(orc_bounds2i b);                                             /* bounds.rs:195-198 */
int orc_bounds2i_inside_exclusive(orc_bounds2i b, int64_t x, int64_t y);               /* bounds.rs:210-212 */
/* bounds.rs:284-288: writes min(cap, n) points as x,y pairs in iteration order, returns n */
int64_t orc_bounds2i_iter(orc_bounds2i b, int64_t *xy_out, int64_t cap);
void orc_point2f_floor(const float in[2], float out[2]); /* point.rs:293-295 */
void orc_point2f_ceil(const float in[2], float out[2]);  /* point.rs:306-308 */

/* ---- spectrum: src/core/spectrum.rs:129-145 ---- */
void orc_rgb_to_xyz(const float rgb[3], float xyz[3]);
void orc_xyz_to_rgb(const float xyz[3], float rgb[3]);

/* ---- filters ---- */
enum { ORC_FILTER_BOX = 0, ORC_FILTER_TRIANGLE = 1, ORC_FILTER_GAUSSIAN = 2,
       ORC_FILTER_MITCHELL = 3, ORC_FILTER_LANCZOS = 4 };
typedef struct {
    int kind;
    float radius[2], inv_radius[2];
    float p0, p1;        /* gaussian: alpha,-; mitchell: B,C; lanczos: tau,- */
    float exp_x, exp_y;  /* gaussian only */
} orc_filter;
/* box.rs:37-42 (kind 0); kinds 1..4 are ext (pbrt-v3 ch. 7.8) */
void orc_filter_init(orc_filter *f, int kind, float rx, float ry, float p0, float p1);
/* box.rs:57-61: radius from xwidth/ywidth, defaults 0.5 when has_* is 0 */
void orc_box_filter_create(orc_filter *f, int has_xwidth, float xwidth, int has_ywidth, float ywidth);
float orc_filter_evaluate(const orc_filter *f, float px, float py);
/* film.rs:113-123 */
void orc_filter_table(const orc_filter *f, float table[256]);

/* ---- film: src/core/film.rs ---- */
orc_film *orc_film_new(int64_t xres, int64_t yres, const float crop[4], const float radius[2],
                       const float table[256], float diagonal_mm, float scale,
                       float max_sample_luminance);                         /* :82-137 */
void orc_film_free(orc_film *f);
orc_bounds2i orc_film_cropped_pixel_bounds(const orc_film *f);
orc_bounds2i orc_film_get_sample_bounds(const orc_film *f);                 /* :166-175 */
orc_bounds2f orc_film_get_physical_extent(const orc_film *f);               /* :218-227 */
orc_bounds2i orc_film_tile_bounds(const orc_film *f, orc_bounds2i sample_bounds); /* :264-273 */
orc_tile *orc_film_get_film_tile(const orc_film *f, orc_bounds2i sample_bounds);  /* :264-281 */
void orc_film_merge_film_tile(orc_film *f, orc_tile *t); /* :313-326; consumes (frees) t */
/* :340-372 — the rgb buffer handed to imageio::write_image; 3*area floats */
void orc_film_write_image_rgb(const orc_film *f, float splat_scale, float *rgb_out);
void orc_film_get_pixel_xyz(const orc_film *f, int64_t x, int64_t y, float out[3]); /* :405-410 */
orc_pixel *orc_film_pixels(orc_film *f);
int64_t orc_film_pixel_count(const orc_film *f);
const float *orc_film_table(const orc_film *f);

/* FilmTile: film.rs:428-489 */
void orc_tile_free(orc_tile *t);
orc_bounds2i orc_tile_get_pixel_bounds(const orc_tile *t);
int64_t orc_tile_pixel_count(const orc_tile *t);   /* max(0, area()) quirk, :446 */
orc_tile_pixel *orc_tile_pixels(orc_tile *t);
orc_tile_pixel *orc_tile_get_pixel(orc_tile *t, int64_t x, int64_t y); /* NULL if outside */

/* ---- textures: src/textures/constant.rs:61-68,96-103,139-141 ---- */
void orc_constant_texture_eval_f32(int has_value, float value, uint64_t n, float *out);
void orc_constant_texture_eval_rgb(int has_value, const float value[3], uint64_t n, float *out);

/* ---- mipmap.rs:43-52 ---- */
void orc_weight_lut(float out[128]);

/* ---- rng.rs:19-93 ---- */
typedef struct { uint64_t state, inc; } orc_rng;
void orc_rng_default(orc_rng *r);
void orc_rng_set_sequence(orc_rng *r, uint64_t sequence_index);
uint32_t orc_rng_uniform_u32(orc_rng *r);
uint32_t orc_rng_uniform_u32_threshold(orc_rng *r, uint32_t b);
float orc_rng_uniform_float(orc_rng *r);

/* ---- imageio.rs:186-213: PFM byte image (header + bottom-to-top rows); returns bytes written ---- */
size_t orc_pfm_encode(const float *rgb, int64_t width, int64_t height, uint8_t *out, size_t cap);

/* =================== Tier 2 — extension, parity unpinned =================== */
/* pbrt-v3 FilmTile::AddSample on the reference's FilmTile fields */
void orc_ext_tile_add_sample(orc_tile *t, float px, float py, const float L[3], float sample_weight);
/* n samples: xy[2n], rgbw[4n] = (L.r, L.g, L.b, sample_weight), stream order */
void orc_ext_tile_add_samples(orc_tile *t, uint64_t n, const float *xy, const float *rgbw);
/* pbrt-v3 Film::AddSplat on splat_xyz (film.rs:50-52,334-336) */
void orc_ext_film_add_splat(orc_film *f, float px, float py, const float v[3]);
/* Synthetic stratified stream, SURVEY.md App. C: pixel-major over `bounds`, spp = n*n per pixel */
void orc_ext_synth_samples(orc_bounds2i bounds, int spp, uint64_t seed, float *xy, float *rgbw);
/* App. C tile fill: pixel p of tile `tile_index` gets rgb from Rng(seed<<32 + tile_index<<20 + p), w=1 */
void orc_ext_synth_tile_fill(orc_tile *t, uint64_t seed, uint64_t tile_index);
/*
 * Whole-frame pass used as the timed CPU arm: get_film_tile(sample_bounds) -> add_samples ->
 * merge_film_tile, with the pixel rows of the tile split over `threads` workers.  Every pixel
 * still receives its samples in stream order, so the result is identical for any thread count.
 */
void orc_ext_film_add_samples_pass(orc_film *f, orc_bounds2i sample_bounds, int spp,
                                   const float *xy, const float *rgbw, int threads);

#ifdef __cplusplus
}
#endif
#endif
