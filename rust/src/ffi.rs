//! ffi.rs — UNTESTED SOURCE (no rustc/cargo in the build image).
//!
//! `extern "C"` declarations for every symbol of include/pbrt_b200.h: first the ones the film / filter /
//! texture shims use, then the rest.  Everything returns a status code; nothing unwinds across the boundary.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_float, c_int, c_void};

#[repr(C)]
pub struct PbrtFilm {
    _private: [u8; 0],
}

pub const PBRT_OK: c_int = 0;
pub const PBRT_E_RANGE: c_int = 3;
pub const PBRT_SPLAT_EXACT: c_int = 0;
pub const PBRT_SPLAT_FMA: c_int = 1;
pub const PBRT_MEM_HOST: c_int = 0;
pub const PBRT_MEM_DEVICE: c_int = 1;
pub const PBRT_MEM_PINNED_ASYNC: c_int = 2;

extern "C" {
    pub fn pbrt_b200_init(device: c_int) -> c_int;
    pub fn pbrt_b200_last_error() -> *const c_char;
    pub fn pbrt_film_create(
        xres: i32,
        yres: i32,
        crop: *const c_float,   // [4] = {min.x, min.y, max.x, max.y}
        radius: *const c_float, // [2]
        table: *const c_float,  // [256], film.rs:113-123
        diagonal_mm: c_float,
        scale: c_float,
        max_sample_luminance: c_float,
        out: *mut *mut PbrtFilm,
    ) -> c_int;
    pub fn pbrt_film_destroy(film: *mut PbrtFilm) -> c_int;
    pub fn pbrt_film_cropped_pixel_bounds(film: *const PbrtFilm, out: *mut i32) -> c_int;
    pub fn pbrt_film_get_sample_bounds(film: *const PbrtFilm, out: *mut i32) -> c_int;
    pub fn pbrt_film_get_physical_extent(film: *const PbrtFilm, out: *mut c_float) -> c_int;
    pub fn pbrt_film_tile_bounds(film: *const PbrtFilm, sample_bounds: *const i32, out: *mut i32, pixel_count: *mut i64) -> c_int;
    pub fn pbrt_film_geometry(
        xres: i32,
        yres: i32,
        crop_window: *const c_float,
        filter_radius: *const c_float,
        diagonal_mm: c_float,
        rank: c_int,
        nranks: c_int,
        cropped: *mut i32,
        owned: *mut i32,
        sample_bounds: *mut i32,
        physical_extent: *mut c_float,
    ) -> c_int;
    pub fn pbrt_film_geometry_tile_bounds(
        clip: *const i32,
        filter_radius: *const c_float,
        sample_bounds: *const i32,
        out: *mut i32,
        pixel_count: *mut i64,
    ) -> c_int;
    pub fn pbrt_film_route_plan(
        sample_bounds: *const i32,
        cropped: *const i32,
        filter_radius: *const c_float,
        nranks: i32,
        src_rows: *const i32, // [2]
        out_rows: *mut i32,   // [2 * nranks]
    ) -> c_int;
    pub fn pbrt_film_merge_tile(film: *mut PbrtFilm, tile_bounds: *const i32, rgbw: *const c_float, src_is_device: c_int) -> c_int;
    pub fn pbrt_film_add_samples_tile(
        film: *mut PbrtFilm,
        sample_bounds: *const i32,
        spp: i32,
        xy: *const c_float,
        rgbw: *const c_float,
        src_is_device: c_int,
        mode: c_int,
    ) -> c_int;
    pub fn pbrt_film_add_samples_tile_rgb(
        film: *mut PbrtFilm,
        sample_bounds: *const i32,
        spp: i32,
        xy: *const c_float,
        rgb: *const c_float,
        sample_weight: *const c_float, // null = every weight is 1
        src_is_device: c_int,
        mode: c_int,
    ) -> c_int;
    pub fn pbrt_film_resolve_rgb(film: *const PbrtFilm, splat_scale: c_float, out_rgb: *mut c_float, dst_is_device: c_int) -> c_int;
    pub fn pbrt_film_resolve_rgb8(film: *const PbrtFilm, splat_scale: c_float, out_rgb8: *mut u8, dst_is_device: c_int) -> c_int;
    pub fn pbrt_film_get_pixel_xyz(film: *const PbrtFilm, x: i32, y: i32, out: *mut c_float) -> c_int;
    pub fn pbrt_film_check(film: *mut PbrtFilm) -> c_int;
    pub fn pbrt_texture_constant_eval_f32(value: c_float, n: u64, out: *mut c_float, dst_is_device: c_int) -> c_int;
    pub fn pbrt_texture_constant_eval_rgb(value: *const c_float, n: u64, out: *mut c_float, dst_is_device: c_int) -> c_int;
}

/// The rest of include/pbrt_b200.h (runtime plumbing, host filters, batched / sharded / extension entry points),
/// declared for completeness; the film / texture shims above do not need them.
#[repr(C)]
pub struct PbrtFilter {
    _private: [u8; 0],
}

extern "C" {
    pub fn pbrt_b200_version() -> c_int;
    pub fn pbrt_b200_set_stream(cuda_stream: *mut c_void) -> c_int;
    pub fn pbrt_b200_synchronize() -> c_int;
    pub fn pbrt_b200_device_info(device: *mut c_int, sm_count: *mut c_int, cc_major: *mut c_int, cc_minor: *mut c_int, hbm_bytes: *mut u64) -> c_int;
    pub fn pbrt_b200_launch_count() -> u64;
    pub fn pbrt_b200_overlap_passes(on: c_int) -> c_int;
    pub fn pbrt_b200_malloc(bytes: u64, dev_out: *mut *mut c_void) -> c_int;
    pub fn pbrt_b200_free(dev: *mut c_void) -> c_int;
    pub fn pbrt_b200_host_alloc(bytes: u64, host_out: *mut *mut c_void) -> c_int;
    pub fn pbrt_b200_host_free(host: *mut c_void) -> c_int;
    pub fn pbrt_b200_memcpy_h2d(dev: *mut c_void, host: *const c_void, bytes: u64) -> c_int;
    pub fn pbrt_b200_memcpy_d2h(host: *mut c_void, dev: *const c_void, bytes: u64) -> c_int;
    pub fn pbrt_b200_memset(dev: *mut c_void, byte: c_int, bytes: u64) -> c_int;
    pub fn pbrt_b200_ipc_export(dev: *mut c_void, handle: *mut u8) -> c_int;
    pub fn pbrt_b200_ipc_import(handle: *const u8, dev_out: *mut *mut c_void) -> c_int;
    pub fn pbrt_b200_ipc_close(dev: *mut c_void) -> c_int;
    pub fn pbrt_filter_create(kind: c_int, radius_x: c_float, radius_y: c_float, p0: c_float, p1: c_float, out: *mut *mut PbrtFilter) -> c_int;
    pub fn pbrt_box_filter_create_from_params(has_xwidth: c_int, xwidth: c_float, has_ywidth: c_int, ywidth: c_float, out: *mut *mut PbrtFilter) -> c_int;
    pub fn pbrt_filter_destroy(f: *mut PbrtFilter);
    pub fn pbrt_filter_evaluate(f: *const PbrtFilter, x: c_float, y: c_float) -> c_float;
    pub fn pbrt_filter_radius(f: *const PbrtFilter, out: *mut c_float);
    pub fn pbrt_filter_inv_radius(f: *const PbrtFilter, out: *mut c_float);
    pub fn pbrt_filter_table(f: *const PbrtFilter, table: *mut c_float) -> c_int;
    pub fn pbrt_film_create_sharded(xres: i32, yres: i32, crop: *const c_float, radius: *const c_float, table: *const c_float, diagonal_mm: c_float, scale: c_float, max_sample_luminance: c_float, rank: c_int, nranks: c_int, out: *mut *mut PbrtFilm) -> c_int;
    pub fn pbrt_film_owned_pixel_bounds(film: *const PbrtFilm, out: *mut i32) -> c_int;
    pub fn pbrt_film_merge_tiles(film: *mut PbrtFilm, ntiles: i32, tile_bounds: *const i32, offsets: *const i64, rgbw: *const c_float, total_pixels: i64, src_is_device: c_int) -> c_int;
    pub fn pbrt_film_add_samples_tiles(film: *mut PbrtFilm, ntiles: i32, sample_bounds: *const i32, sample_offsets: *const i64, spp: i32, xy: *const c_float, rgbw: *const c_float, total_samples: i64, src_is_device: c_int, mode: c_int) -> c_int;
    pub fn pbrt_film_add_samples(film: *mut PbrtFilm, sample_bounds: *const i32, n: u64, xy: *const c_float, rgbw: *const c_float, src_is_device: c_int) -> c_int;
    pub fn pbrt_film_add_splats(film: *mut PbrtFilm, n: u64, xy: *const c_float, rgb: *const c_float, src_is_device: c_int) -> c_int;
    pub fn pbrt_film_set_image(film: *mut PbrtFilm, rgb: *const c_float, src_is_device: c_int) -> c_int;
    pub fn pbrt_film_clear(film: *mut PbrtFilm) -> c_int;
    pub fn pbrt_film_read_pixels(film: *const PbrtFilm, out7: *mut c_float, dst_is_device: c_int) -> c_int;
    pub fn pbrt_film_device_buffers(film: *const PbrtFilm, xyzw: *mut *mut c_void, splat: *mut *mut c_void, npixels: *mut i64) -> c_int;
    pub fn pbrt_film_resolve_rgb_to_frames(film: *const PbrtFilm, splat_scale: c_float, nframes: i32, frames: *const *mut c_void) -> c_int;
    pub fn pbrt_mipmap_weight_lut(out: *mut c_float) -> c_int;
    pub fn pbrt_synth_samples(bounds: *const i32, index_bounds: *const i32, spp: i32, seed: u64, xy_dev: *mut c_float, rgbw_dev: *mut c_float) -> c_int;
    pub fn pbrt_synth_tiles(ntiles: i32, offsets: *const i64, counts: *const i64, seed: u64, rgbw_dev: *mut c_float, total_pixels: i64) -> c_int;
}

/// Map a status to the reference's error behaviour: programming errors panic (film.rs:391-402 are
/// `debug_assert!` + `unwrap()`), exactly where the reference would.
pub fn check(rc: c_int) {
    if rc != PBRT_OK {
        let msg = unsafe { std::ffi::CStr::from_ptr(pbrt_b200_last_error()) }.to_string_lossy().into_owned();
        panic!("pbrt_b200: {} (status {})", msg, rc);
    }
}

pub fn as_void<T>(p: *const T) -> *const c_void {
    p as *const c_void
}
