/*
 * pbrt_b200.h — C ABI of the B200 film-reconstruction / texture-evaluation library.
 *
 * This is the drop-in boundary for the hot path of wathiede/pbrt (a Rust crate with no FFI of
 * its own): each entry point is what a Rust `extern "C"` block behind the reference's existing
 * `Film`, `FilmTile`, `Filter` and `Texture` signatures would bind.  The reference interface an
 * entry point replaces is cited as file:line relative to the reference root.  INTEGRATION.md
 * shows the Rust-side binding.
 *
 * Conventions
 *   - every function returns a PbrtStatus (0 = OK) unless stated; on failure
 *     pbrt_b200_last_error() holds a thread-local message.  Nothing unwinds across the ABI.
 *   - Float is f32, Spectrum is RGB (the reference's default features, src/lib.rs:24-44,
 *     src/core/spectrum.rs:151-153).  The f64 / sampled-spectrum builds are not supported.
 *   - bounds are int32 {x0, y0, x1, y1}, max exclusive (Bounds2i, src/core/geometry/bounds.rs).
 *     The reference uses isize; a host shim must range-check before narrowing.
 *   - `*_is_device` = 0 (PBRT_MEM_HOST): pointer is host memory, the call copies and, for outputs,
 *     returns when the data has arrived; 1 (PBRT_MEM_DEVICE): pointer is device memory on the film's
 *     device and is used in place; 2 (PBRT_MEM_PINNED_ASYNC, accepted by add_samples_tile[_rgb] and
 *     resolve_rgb / resolve_rgb8): pointer is page-locked host memory and the transfer is only
 *     enqueued — inputs are double-buffered on a copy stream so that the upload of one call overlaps
 *     the kernels of the previous one; the caller keeps the buffers untouched until
 *     pbrt_b200_synchronize().
 *   - work is ordered on one stream per process (pbrt_b200_set_stream); calls that return data
 *     to the host synchronise that stream, the others are asynchronous.
 *   - compute entry points may be called from several host threads at once (FilmTile is Send: workers
 *     merge their own tiles); each call takes one process-wide lock, as the reference takes its pixel
 *     Mutex (src/core/film.rs:73, :316), so concurrent calls are serialised in arrival order.
 *   - there is no CPU fallback: every compute entry point fails with PBRT_E_CUDA when no
 *     sm_100 device is usable.
 *
 * Tier marks: [T1] reference-backed behaviour; [T2] extension — the reference declares the
 * item but has no implementation (no reference parity possible); [UTIL] plumbing.
 */
#ifndef PBRT_B200_H
#define PBRT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    PBRT_OK = 0,
    PBRT_E_INVALID = 1,        /* bad argument */
    PBRT_E_CUDA = 2,           /* CUDA runtime failure / no device */
    PBRT_E_RANGE = 3,          /* point or bounds outside the film (reference: debug_assert / unwrap panic) */
    PBRT_E_NOT_PIXEL_MAJOR = 4,/* add_samples_tile: a sample lies outside its nominal pixel */
    PBRT_E_UNSUPPORTED = 5,
    PBRT_E_NOMEM = 6,
    PBRT_E_NONFINITE = 7       /* add_samples_tile: a sample's radiance * weight is not finite */
} PbrtStatus;

#define PBRT_FILTER_TABLE_WIDTH 16 /* src/core/film.rs:34 */

enum { PBRT_MEM_HOST = 0, PBRT_MEM_DEVICE = 1, PBRT_MEM_PINNED_ASYNC = 2 };

typedef struct PbrtFilm PbrtFilm;     /* src/core/film.rs:59-76 — device-resident */
typedef struct PbrtFilter PbrtFilter; /* src/core/filter.rs:22-29 — host object */

/* ------------------------------------------------------------------ runtime [UTIL] */
int pbrt_b200_version(void);
const char *pbrt_b200_last_error(void);
/* bind the process to CUDA device `device` and create the library stream */
int pbrt_b200_init(int device);
/* run all subsequent work on the caller's stream (a cudaStream_t; NULL = library stream,
 * (void*)0x1 = cudaStreamLegacy, the legacy default stream) */
int pbrt_b200_set_stream(void *cuda_stream);
int pbrt_b200_synchronize(void);
int pbrt_b200_device_info(int *device, int *sm_count, int *cc_major, int *cc_minor, uint64_t *hbm_bytes);
/* kernels launched by this library since init (bench.py's gpu_launches) */
uint64_t pbrt_b200_launch_count(void);
/* Consecutive pixel-major splat passes overlap (the next pass starts in the SM slots the one ahead leaves and waits for
 * it before it touches the film) when both read the same sample buffers.  on = 1 extends that to passes reading any
 * buffers: the caller then guarantees that the samples of a pass were complete before the PREVIOUS call on the stream
 * was issued (generated a pass ahead, or on another stream joined earlier) — a pass no longer waits for the kernel ahead
 * of it before its first load.  Returns the previous setting; default 0. */
int pbrt_b200_overlap_passes(int on);
/* raw device / pinned-host buffers for callers without a CUDA runtime binding */
int pbrt_b200_malloc(uint64_t bytes, void **dev_out);
int pbrt_b200_free(void *dev);
int pbrt_b200_host_alloc(uint64_t bytes, void **host_out); /* pinned */
int pbrt_b200_host_free(void *host);
int pbrt_b200_memcpy_h2d(void *dev, const void *host, uint64_t bytes);
int pbrt_b200_memcpy_d2h(void *host, const void *dev, uint64_t bytes);
int pbrt_b200_memset(void *dev, int byte, uint64_t bytes);
/* CUDA IPC for buffers from pbrt_b200_malloc: lets the ranks of one box map each other's frames */
int pbrt_b200_ipc_export(void *dev, uint8_t handle[64]);
int pbrt_b200_ipc_import(const uint8_t handle[64], void **dev_out);
int pbrt_b200_ipc_close(void *dev);

/* ------------------------------------------------------------------ filters (host) */
enum { PBRT_FILTER_BOX = 0,       /* [T1] src/filters/box.rs:30-77 */
       PBRT_FILTER_TRIANGLE = 1,  /* [T2] named at src/core/api.rs:954 */
       PBRT_FILTER_GAUSSIAN = 2,  /* [T2] p0 = alpha */
       PBRT_FILTER_MITCHELL = 3,  /* [T2] p0 = B, p1 = C */
       PBRT_FILTER_LANCZOS = 4 }; /* [T2] p0 = tau ("sinc" at api.rs:954) */
/* BoxFilter::new (box.rs:37-42) and the [T2] constructors */
int pbrt_filter_create(int kind, float radius_x, float radius_y, float p0, float p1, PbrtFilter **out);
/* BoxFilter::create_box_filter (box.rs:57-61): xwidth / ywidth default 0.5 when has_* = 0 */
int pbrt_box_filter_create_from_params(int has_xwidth, float xwidth, int has_ywidth, float ywidth, PbrtFilter **out);
void pbrt_filter_destroy(PbrtFilter *f);
float pbrt_filter_evaluate(const PbrtFilter *f, float x, float y);  /* Filter::evaluate, filter.rs:24 */
void pbrt_filter_radius(const PbrtFilter *f, float out[2]);        /* Filter::radius, filter.rs:26 */
void pbrt_filter_inv_radius(const PbrtFilter *f, float out[2]);    /* Filter::inv_radius, filter.rs:28 */
/* the 16x16 table Film::new precomputes (film.rs:113-123); row-major, y outer */
int pbrt_filter_table(const PbrtFilter *f, float table[256]);

/* ------------------------------------------------------------------ film */
/*
 * [T1] Film::new (film.rs:82-137).  The device contract for a filter is its radius and the
 * 16x16 table, so any host `Filter` impl works: the shim fills `table` by calling
 * filter.evaluate() exactly as film.rs:116-123 does (or pbrt_filter_table for built-ins).
 * `crop` = {min.x, min.y, max.x, max.y}.
 */
int pbrt_film_create(int32_t xres, int32_t yres, const float crop[4], const float radius[2],
                     const float table[256], float diagonal_mm, float scale, float max_sample_luminance,
                     PbrtFilm **out);
/*
 * [UTIL] Row-sharded film for one rank of `nranks`: identical to pbrt_film_create except that the
 * film stores (and clips to) rows [y0 + rank*H/nranks, y0 + (rank+1)*H/nranks) of the cropped
 * pixel bounds.  Clipping at the shard edge is the reference's own `∩ cropped_pixel_bounds`
 * (film.rs:272-273) applied to the row block.
 */
int pbrt_film_create_sharded(int32_t xres, int32_t yres, const float crop[4], const float radius[2],
                             const float table[256], float diagonal_mm, float scale,
                             float max_sample_luminance, int rank, int nranks, PbrtFilm **out);
int pbrt_film_destroy(PbrtFilm *film); /* Drop */
/* pub field cropped_pixel_bounds (film.rs:72) — of the whole film, also when sharded */
int pbrt_film_cropped_pixel_bounds(const PbrtFilm *film, int32_t out[4]);
/* rows/columns this handle stores; equals cropped_pixel_bounds unless sharded */
int pbrt_film_owned_pixel_bounds(const PbrtFilm *film, int32_t out[4]);
int pbrt_film_get_sample_bounds(const PbrtFilm *film, int32_t out[4]);   /* [T1] film.rs:166-175 */
int pbrt_film_get_physical_extent(const PbrtFilm *film, float out[4]);  /* [T1] film.rs:218-227 */
/*
 * [T1] the bounds arithmetic of Film::get_film_tile (film.rs:264-273).  `pixel_count` receives
 * max(0, area) as FilmTile::new computes it (film.rs:446), including its quirk for a doubly
 * inverted box.  For a sharded film the result is clipped to the owned rows.
 */
int pbrt_film_tile_bounds(const PbrtFilm *film, const int32_t sample_bounds[4], int32_t out[4],
                          int64_t *pixel_count);
/*
 * [T1] Film::new's crop bounds (film.rs:92-101), get_sample_bounds (:166-175), get_physical_extent (:218-227) and
 * get_film_tile's bounds (:264-273) as plain host functions: no film object, no device.  Any output pointer
 * of pbrt_film_geometry may be NULL.  `owned` is the row block of shard `rank` of `nranks` (== cropped for 1).
 * pbrt_film_geometry_tile_bounds clips against `clip` (the cropped bounds, or a shard's owned bounds).
 */
int pbrt_film_geometry(int32_t xres, int32_t yres, const float crop_window[4], const float filter_radius[2],
                       float diagonal_mm, int rank, int nranks, int32_t cropped[4], int32_t owned[4],
                       int32_t sample_bounds[4], float physical_extent[4]);
int pbrt_film_geometry_tile_bounds(const int32_t clip[4], const float filter_radius[2], const int32_t sample_bounds[4],
                                   int32_t out[4], int64_t *pixel_count);
/*
 * [UTIL] sample routing for a row-sharded film: a source that holds the nominal sample rows
 * [src_rows[0], src_rows[1]) of a pixel-major stream over `sample_bounds` owes shard g of `nranks` the rows
 * out_rows[2g] .. out_rows[2g+1] (its own row block plus floor(r.y + .5) halo rows either side, clipped; empty when
 * equal).  In a pixel-major stream that is one contiguous run of samples starting at
 * (out_rows[2g] - src_rows[0]) * width * spp; rows within the halo of a shard edge belong to two runs.  Host
 * arithmetic only; the exchange itself is the caller's (pbrt_b200/dist.py:route_samples uses NCCL send / recv).
 */
int pbrt_film_route_plan(const int32_t sample_bounds[4], const int32_t cropped[4], const float filter_radius[2],
                         int32_t nranks, const int32_t src_rows[2], int32_t out_rows[]);
/*
 * [T1] Film::merge_film_tile (film.rs:313-326).  `rgbw` is the tile's Vec<FilmTilePixel>
 * (film.rs:39-42): 4 floats per pixel {contrib_sum.rgb, filter_weight_sum}, row-major over
 * `tile_bounds`.  The tile is consumed by value in the reference, so the buffer is only read.
 * PBRT_E_RANGE if the tile is not inside the film (reference: debug_assert + unwrap, :390-402).
 */
int pbrt_film_merge_tile(PbrtFilm *film, const int32_t tile_bounds[4], const float *rgbw, int src_is_device);
/*
 * [T1] the same for `ntiles` tiles in ONE launch.  tile_bounds = ntiles x 4; tile i's pixels
 * start at rgbw + 4*offsets[i].  Equivalent to calling merge_film_tile for i = 0..ntiles-1 in
 * order: where tiles overlap, every film pixel receives its contributions in tile order.
 */
int pbrt_film_merge_tiles(PbrtFilm *film, int32_t ntiles, const int32_t *tile_bounds, const int64_t *offsets,
                          const float *rgbw, int64_t total_pixels, int src_is_device);
/*
 * [T2] get_film_tile(sample_bounds) -> FilmTile::add_sample for every sample -> merge_film_tile
 * (the method the unused FilmTile fields at film.rs:428-436 exist for; pbrt-v3 7.9.2), as one
 * kernel.  Samples are pixel-major over `sample_bounds` with `spp` per pixel: sample k of
 * pixel (px,py) sits at index ((py-y0)*W + (px-x0))*spp + k and must lie in
 * [px,px+1] x [py,py+1] (closed: px + jitter may round up onto the next pixel boundary).  xy = 2 floats, rgbw = {L.r, L.g, L.b, sample_weight} per sample.
 * Per pixel the samples are accumulated in stream order, then converted and added to the film
 * exactly as merge_film_tile does.  Radiance must be finite (pbrt zeroes non-finite samples before
 * AddSample); violations of either contract are reported by pbrt_film_check and leave the film
 * contents undefined.  Non-finite radiance is detected where it lands: in every mode a pixel sum that became
 * inf / NaN sets PBRT_E_NONFINITE (a sample whose footprint misses the film entirely goes unnoticed, and unharmed).
 * Radius 2 or 4 on both axes (triangle, gaussian, mitchell, lanczos defaults) with at most 32 samples per pixel
 * per call runs the phase-class kernel (splat_class.cu); other radii the window kernel; radii beyond 4 pixels
 * or unequal on the two axes a generic gather.  All three give identical results.
 */
enum { PBRT_SPLAT_EXACT = 0,   /* gather, mul then add: bit-identical to the CPU restatement */
       PBRT_SPLAT_FMA = 1,     /* gather, same order, fused multiply-add */
       PBRT_SPLAT_ATOMIC = 2 };/* scatter with shared-memory atomics; order not deterministic */
int pbrt_film_add_samples_tile(PbrtFilm *film, const int32_t sample_bounds[4], int32_t spp, const float *xy,
                               const float *rgbw, int src_is_device, int mode);
/*
 * [T2] the same with the radiance as separate streams, the shape of pbrt's AddSample(pFilm, L, sampleWeight)
 * arguments: rgb = 3 floats per sample, sample_weight = 1 float per sample or NULL for "every weight is 1"
 * (then nothing is transferred for it: 20 instead of 24 bytes per sample from a host buffer).  The streams
 * are interleaved on the device and the same kernels run; results are bit-identical to add_samples_tile
 * with rgbw = {rgb, sample_weight}.
 */
int pbrt_film_add_samples_tile_rgb(PbrtFilm *film, const int32_t sample_bounds[4], int32_t spp, const float *xy,
                                   const float *rgb, const float *sample_weight, int src_is_device, int mode);
/*
 * [T2] the same for `ntiles` tiles at once, the way a renderer works (pbrt renders 16x16-pixel tiles):
 * tile i has sample bounds sample_bounds[4i..4i+3] and its pixel-major samples start at sample
 * sample_offsets[i] of xy / rgbw.  Equivalent to add_samples_tile for i = 0 .. ntiles-1 in order: every
 * tile is accumulated on its own (FilmTile semantics), then the tiles are merged in tile order, so
 * pixels in overlapping tile borders receive their contributions exactly as sequential calls would give.
 * Two launches in total (splat into per-tile RGBW buffers, ordered merge) instead of 2 per tile.
 */
int pbrt_film_add_samples_tiles(PbrtFilm *film, int32_t ntiles, const int32_t *sample_bounds,
                                const int64_t *sample_offsets, int32_t spp, const float *xy, const float *rgbw,
                                int64_t total_samples, int src_is_device, int mode);
/*
 * [T2] the same for samples in arbitrary order and position (no pixel-major contract):
 * scatter with global atomics into a scratch tile, then merge.  Order not deterministic.
 */
int pbrt_film_add_samples(PbrtFilm *film, const int32_t sample_bounds[4], uint64_t n, const float *xy,
                          const float *rgbw, int src_is_device);
/* [T2] Film::add_splat (film.rs:334-336, unimplemented! in the reference) for n points */
int pbrt_film_add_splats(PbrtFilm *film, uint64_t n, const float *xy, const float *rgb, int src_is_device);
/* [T2] Film::set_image (film.rs:329-331) / Film::clear (film.rs:386-388), unimplemented! in the reference */
int pbrt_film_set_image(PbrtFilm *film, const float *rgb, int src_is_device);
int pbrt_film_clear(PbrtFilm *film);
/*
 * [T1] the pixel loop of Film::write_image (film.rs:340-372): fills the rgb buffer the reference
 * hands to imageio::write_image, 3 floats per owned pixel, row-major.
 */
int pbrt_film_resolve_rgb(const PbrtFilm *film, float splat_scale, float *out_rgb, int dst_is_device);
/* [T1] the same fused with imageio's to_byte (imageio.rs:66-68, lib.rs:93-99): 3 bytes per pixel */
int pbrt_film_resolve_rgb8(const PbrtFilm *film, float splat_scale, uint8_t *out_rgb8, int dst_is_device);
/* [T1] Film::get_pixel_xyz (film.rs:405-410) */
int pbrt_film_get_pixel_xyz(const PbrtFilm *film, int32_t x, int32_t y, float out[3]);
/* [UTIL] all owned pixels as the reference's Pixel (film.rs:47-55): 7 floats each */
int pbrt_film_read_pixels(const PbrtFilm *film, float *out7, int dst_is_device);
/*
 * [UTIL] device storage, for the multi-GPU plumbing (NCCL all-gather of the row blocks):
 * xyzw = float4 {xyz, filter_weight_sum} per owned pixel, splat = 3 floats per owned pixel.
 */
int pbrt_film_device_buffers(const PbrtFilm *film, void **xyzw, void **splat, int64_t *npixels);
/*
 * [T1]+[UTIL] the pixel loop of Film::write_image (film.rs:346-372) fused with the final assembly of a
 * row-sharded film: resolves this film's rows and stores them straight into `nframes` full-frame rgb
 * buffers — this rank's own and its peers' (mapped with pbrt_b200_ipc_import) — at the rows this film
 * owns.  frames[i] holds 3 floats per pixel of cropped_pixel_bounds.  Every rank calling this once,
 * followed by a barrier, is the all-gather: one kernel, peer stores over NVLink, no staging copy.
 */
int pbrt_film_resolve_rgb_to_frames(const PbrtFilm *film, float splat_scale, int32_t nframes, void *const *frames);
/* [UTIL] sticky asynchronous error of the film's kernels (e.g. PBRT_E_NOT_PIXEL_MAJOR); clears it */
int pbrt_film_check(PbrtFilm *film);

/* ------------------------------------------------------------------ textures */
/*
 * [T1] ConstantTexture<T>::evaluate (src/textures/constant.rs:139-141) for a batch of n lookups.
 * SurfaceInteraction is a zero-sized struct (src/core/interaction.rs:22-23), so a lookup has no
 * input.  The scalar trait method stays on the host; these are the throughput entry points.
 */
int pbrt_texture_constant_eval_f32(float value, uint64_t n, float *out, int dst_is_device);
int pbrt_texture_constant_eval_rgb(const float value[3], uint64_t n, float *out, int dst_is_device);

/* ------------------------------------------------------------------ misc */
/* [T1] mipmap.rs:43-52 — the only implemented piece of MIPMap; computed on the device */
int pbrt_mipmap_weight_lut(float out[128]);
/*
 * [UTIL] synthetic stratified sample stream generated in place in device
 * memory with the reference's PCG32 (src/core/rng.rs:53-93): xy_dev = 2*n floats,
 * rgbw_dev = 4*n floats, n = area(bounds)*spp.  `index_bounds` (may be NULL = bounds) is the
 * rectangle pixel indices are taken over, so a row shard reproduces its slice of the full stream.
 */
int pbrt_synth_samples(const int32_t bounds[4], const int32_t index_bounds[4], int32_t spp, uint64_t seed,
                       float *xy_dev, float *rgbw_dev);
/* [UTIL] App. C tile fill for the merge workload: device rgbw for ntiles tiles */
int pbrt_synth_tiles(int32_t ntiles, const int64_t *offsets, const int64_t *counts, uint64_t seed, float *rgbw_dev,
                     int64_t total_pixels);

#ifdef __cplusplus
}
#endif
#endif /* PBRT_B200_H */
