//! ffi.rs — UNTESTED SOURCE (no rustc/cargo in the build image).
//!
//! `extern "C"` declarations for include/pbrt_b200.h, one per symbol the film / filter / texture
//! shims use.  Everything returns a status code; nothing unwinds across the boundary.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_float, c_int, c_void};

#[repr(C)]
pub struct PbrtFilm {
    _private: [u8; 0],
}

pub const PBRT_OK: c_int = 0;
pub const PBRT_E_RANGE: c_int = 3;
pub const PBRT_SPLAT_EXACT: c_int = 0;

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
